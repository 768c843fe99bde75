// Convolution / Dense family as implicit GEMMs over channels-last tensors (sm_100a).
//
// One geometry description (GemmPlan) drives every variant:
//   pixel mode : D[m][n] = sum_{tap,c} Src[pix(m,tap)][c] * W(tap,c,n)      forward and dgrad
//   wgrad mode : D[(tap,c)][n] = sum_m Src[pix(m,tap)][c] * G[m][n]
// pix(m,tap): u_d = e_d*mstride + off_d(tap), valid iff 0 <= u_d < U_d, source coord = u_d >> ushift.
// That covers TF-SAME asymmetric padding, stride 2, the fused nearest x2 upsample (ushift=1 in
// forward; extra (delta,tap) generalized taps in dgrad) and the parity phases of the stride-2 dgrad.
//
// Plans beyond the plain ones: PHASED plans (several GEMMs that differ only in tap list and output offset in one
// launch: the parity phases of the stride-2 dgrad, the sub-pixel phases of the folded forward) and FOLDED plans
// (nearest x2 upsample + conv evaluated on the low-resolution tensor with pre-summed taps).
//
// Kernels:
//   igemm_tc_pixel_kernel : tcgen05.mma kind::tf32 on big/small tf32 splits of the fp32 operands (3xTF32: 3 products
//                       per k-step, fp32-grade accuracy on the tensor pipe), fp32 accumulators in TMEM with
//                       chunked promotion, A gathered straight into tensor memory, B streamed as pre-packed
//                       128B-swizzled stage images by bulk async copies, mbarrier full/empty pipeline, persistent
//                       CTAs over (M tile, n-tile/phase) items, tcgen05.ld epilogue with fused bias+activation.
//                       <B_MN, WG, PROF>: forward / dgrad / weight gradient, role timers on or off.
//   igemm_ffma_kernel, dense_small_kernel, pixel_smalln_kernel, skinny wgrad kernels, colsum*: fp32 CUDA-core
//                       kernels for tiny-M Dense layers and shapes the tensor-core kernel does not take
//                       (skinny.cu has the dedicated kernels of the 3-channel layers).
//
// Reference call sites: keras Conv2D/Conv3D/Dense in confignet/dnn_models/*.py (see include/confignet_b200.h).
#include "common.cuh"
#include <cuda.h>
#include <map>
#include <string>
#include <vector>
#include <mutex>
#include <algorithm>
#include <string.h>
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------
// Plans (host)
// ------------------------------------------------------------------------------------------------
enum { KIND_FWD = 0, KIND_DGRAD = 1 };

struct HostPlan {
  GemmPlan g;
  bool valid;
};

static std::mutex g_plan_mutex;
static std::map<std::string, HostPlan> g_plans;

static inline int ilog2(int v) { int r = 0; while ((1 << r) < v) ++r; return r; }

static void same_geometry(const cn_conv_desc* d, int U[3], int O[3], int pb[3]) {
  for (int i = 0; i < 3; ++i) {
    if (i < d->nd) {
      U[i] = d->in_dims[i] * d->upsample;
      if (d->pad >= 0) {               // ZeroPadding(pad) + VALID
        O[i] = (U[i] + 2 * d->pad - d->ksize[i]) / d->stride + 1;
        pb[i] = d->pad;
        continue;
      }
      O[i] = (U[i] + d->stride - 1) / d->stride;
      int tot = (O[i] - 1) * d->stride + d->ksize[i] - U[i];
      if (tot < 0) tot = 0;
      pb[i] = tot / 2;                 // TF SAME: the smaller half goes in front
    } else {
      U[i] = 1; O[i] = 1; pb[i] = 0;
    }
  }
}

static int validate_desc(const cn_conv_desc* d) {
  CN_REQUIRE(d != nullptr, CN_ERR_BAD_SHAPE, "null conv descriptor");
  CN_REQUIRE(d->nd == 0 || d->nd == 2 || d->nd == 3, CN_ERR_BAD_SHAPE, "nd must be 0, 2 or 3 (got %d)", d->nd);
  CN_REQUIRE(d->batch > 0 && d->cin > 0 && d->cout > 0, CN_ERR_BAD_SHAPE, "batch/cin/cout must be positive");
  CN_REQUIRE(d->stride == 1 || d->stride == 2, CN_ERR_UNSUPPORTED, "stride must be 1 or 2");
  CN_REQUIRE(d->upsample == 1 || d->upsample == 2, CN_ERR_UNSUPPORTED, "upsample must be 1 or 2");
  CN_REQUIRE(!(d->stride == 2 && d->upsample == 2), CN_ERR_UNSUPPORTED, "stride 2 with fused upsample is unsupported");
  CN_REQUIRE(d->pad >= -1 && d->pad <= 7, CN_ERR_BAD_SHAPE, "pad must be -1 (SAME) or 0..7");
  for (int i = 0; i < d->nd; ++i)
    CN_REQUIRE(d->pad < 0 || d->in_dims[i] * d->upsample + 2 * d->pad >= d->ksize[i], CN_ERR_BAD_SHAPE, "kernel larger than the padded input");
  for (int i = 0; i < 3; ++i) {
    CN_REQUIRE(d->in_dims[i] >= 1 && d->ksize[i] >= 1 && d->ksize[i] <= 7, CN_ERR_BAD_SHAPE, "bad dims/ksize");
    if (i >= d->nd) CN_REQUIRE(d->in_dims[i] == 1 && d->ksize[i] == 1, CN_ERR_BAD_SHAPE, "unused dims must be 1");
    CN_REQUIRE(d->in_dims[i] * d->upsample <= 1000, CN_ERR_UNSUPPORTED, "spatial extent too large for packed coordinates");
  }
  return CN_OK;
}

static int pack_off(const int off[3]) {
  return (off[0] + 8) | ((off[1] + 8) << 10) | ((off[2] + 8) << 20);
}

// Host-side construction of the plan for (desc, kind, phase); phase is only used by the stride-2 dgrad.
static int build_plan(const cn_conv_desc* d, int kind, int phase, GemmPlan* out, std::vector<int2>& taps) {
  int U[3], O[3], pb[3];
  same_geometry(d, U, O, pb);
  GemmPlan g;
  memset(&g, 0, sizeof(g));
  g.n_img = d->batch;
  taps.clear();
  const int kvol = d->ksize[0] * d->ksize[1] * d->ksize[2];
  const int cc = d->cin * d->cout;
  if (kind == KIND_FWD) {
    for (int i = 0; i < 3; ++i) {
      g.E[i] = O[i]; g.U[i] = U[i]; g.S[i] = d->in_dims[i]; g.Q[i] = O[i]; g.ooff[i] = 0;
    }
    g.mstride = d->stride; g.ushift = ilog2(d->upsample); g.ostride = 1;
    g.Csrc = d->cin; g.Cn = d->cout; g.wsc = d->cout; g.wsn = 1;
    for (int t = 0; t < kvol; ++t) {
      int t2 = t % d->ksize[2], t1 = (t / d->ksize[2]) % d->ksize[1], t0 = t / (d->ksize[2] * d->ksize[1]);
      int off[3] = {t0 - pb[0], t1 - pb[1], t2 - pb[2]};
      taps.push_back(make_int2(pack_off(off), t * cc));
    }
  } else {
    g.Csrc = d->cout; g.Cn = d->cin; g.wsc = 1; g.wsn = d->cout;
    int ph[3] = {0, 0, 0};
    if (d->stride == 2) {
      // phase bits: bit i = parity of output coordinate along spatial dim i
      for (int i = 0; i < d->nd; ++i) ph[i] = (phase >> i) & 1;
    }
    for (int i = 0; i < 3; ++i) {
      g.S[i] = O[i]; g.Q[i] = d->in_dims[i];
      if (d->stride == 2 && i < d->nd) {
        g.E[i] = (d->in_dims[i] - ph[i] + 1) / 2; g.U[i] = 2 * O[i]; g.ooff[i] = ph[i];
      } else {
        g.E[i] = d->in_dims[i]; g.U[i] = O[i]; g.ooff[i] = 0;
      }
    }
    if (d->stride == 2) { g.mstride = 2; g.ushift = 1; g.ostride = 2; }
    else if (d->upsample == 2) { g.mstride = 2; g.ushift = 0; g.ostride = 1; }
    else { g.mstride = 1; g.ushift = 0; g.ostride = 1; }
    const int ndelta = (d->upsample == 2) ? (1 << d->nd) : 1;
    for (int dl = 0; dl < ndelta; ++dl) {
      int dd[3] = {0, 0, 0};
      for (int i = 0; i < d->nd; ++i) dd[i] = (dl >> i) & 1;
      for (int t = 0; t < kvol; ++t) {
        int tt[3] = {t / (d->ksize[2] * d->ksize[1]), (t / d->ksize[2]) % d->ksize[1], t % d->ksize[2]};
        int off[3]; bool ok = true;
        for (int i = 0; i < 3; ++i) {
          off[i] = dd[i] + ph[i] + pb[i] - tt[i];
          if (d->stride == 2 && i < d->nd && (off[i] & 1)) ok = false;
        }
        if (!ok) continue;
        taps.push_back(make_int2(pack_off(off), t * cc));
      }
    }
  }
  for (size_t i = 0; i < taps.size(); ++i) {
    int px = taps[i].x;
    int o0 = (px & 1023), o1 = (px >> 10) & 1023, o2 = (px >> 20) & 1023;
    CN_REQUIRE(o0 >= 0 && o0 < 24 && o1 < 24 && o2 < 24, CN_ERR_UNSUPPORTED, "tap offset out of packed range");
  }
  CN_REQUIRE(taps.size() <= 256, CN_ERR_UNSUPPORTED, "too many generalized taps (%d)", (int)taps.size());
  g.ntaps = (int)taps.size();
  g.Ktot = g.ntaps * g.Csrc;
  long long M = (long long)g.n_img * g.E[0] * g.E[1] * g.E[2];
  CN_REQUIRE(M < (1ll << 31), CN_ERR_UNSUPPORTED, "too many output rows");
  g.M = (int)M;
  long long srcpix = (long long)g.n_img * g.S[0] * g.S[1] * g.S[2];
  CN_REQUIRE(srcpix * g.Csrc < (1ll << 32) && (long long)g.n_img * g.Q[0] * g.Q[1] * g.Q[2] * g.Cn < (1ll << 32),
             CN_ERR_UNSUPPORTED, "tensor too large for 32-bit element offsets");
  g.taps = nullptr;
  g.taps_total = g.ntaps;
  *out = g;
  return CN_OK;
}

// Cached device plan (tap table uploaded once; call before CUDA-graph capture).
static int get_plan(const cn_conv_desc* d, int kind, int phase, GemmPlan* out) {
  std::string key((const char*)d, sizeof(*d));
  key.push_back((char)kind);
  key.push_back((char)phase);
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) { *out = it->second.g; return CN_OK; }
  GemmPlan g;
  std::vector<int2> taps;
  int rc = build_plan(d, kind, phase, &g, taps);
  if (rc) return rc;
  int2* dtaps = nullptr;
  if (!taps.empty()) {
    CN_CHECK_CUDA(cudaMalloc(&dtaps, taps.size() * sizeof(int2)));
    CN_CHECK_CUDA(cudaMemcpy(dtaps, taps.data(), taps.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
  g.taps = dtaps;
  HostPlan hp; hp.g = g; hp.valid = true;
  g_plans[key] = hp;
  *out = g;
  return CN_OK;
}

// ------------------------------------------------------------------------------------------------
// Folded plans: nearest x2 upsample followed by a k-tap SAME conv, evaluated on the LOW-resolution tensor
// (hard part 5).  Per axis, output o = 2R+d reads upsampled u = o + t - pb, i.e. source
// R + floor((d + t - pb)/2): the taps t that land on the same source pixel are pre-summed.
//   forward : 2^nd sub-pixel phases (d per axis), each a conv with 2 (k3) or 3|2 (k4) taps per axis on the
//             low-resolution input, written at stride 2 - ONE phased launch.  27 -> 8 taps (3-D k3), 16 -> 6.25 (2-D k4).
//   dgrad   : gx[r] = sum_l gy[2r + l] . Wd[l],  l = pb-(k-1) .. pb+1 per axis, Wd[l] = sum of the taps
//             t in [pb-l, pb-l+1]: one pixel-mode GEMM with (k+1)^nd taps instead of 2^nd * k^nd.
//   wgrad   : gWd[l] = sum_r gy[2r + l] (x) x[r] with the dgrad plan (gy gathered, x dense), then
//             gw[t] = sum of the gWd[l] whose tap set contains t.
// A fold spec packs, per folded tap, the [lo, hi] tap range of each axis (3 bits each).
// ------------------------------------------------------------------------------------------------
enum { KIND_FWD_FOLD = 2, KIND_DGRAD_FOLD = 3, KIND_DGRAD_S2ALL = 4 };

struct FoldInfo {
  std::vector<int> spec;          // per folded tap
  std::vector<int> unfold;        // [kvol][8] folded-tap indices containing original tap t (-1 = none)
  int* d_spec = nullptr;
  int* d_unfold = nullptr;
  int nfold = 0;
};
static std::map<std::string, FoldInfo> g_folds;

__host__ __device__ __forceinline__ void fold_decode(int sp, int lo[3], int hi[3]) {
  lo[0] = sp & 7; hi[0] = (sp >> 3) & 7; lo[1] = (sp >> 6) & 7; hi[1] = (sp >> 9) & 7; lo[2] = (sp >> 12) & 7; hi[2] = (sp >> 15) & 7;
}
static int fold_encode(const int lo[3], const int hi[3]) {
  return lo[0] | (hi[0] << 3) | (lo[1] << 6) | (hi[1] << 9) | (lo[2] << 12) | (hi[2] << 15);
}
static inline int floordiv2(int v) { return (v >= 0) ? v / 2 : -((-v + 1) / 2); }

static bool fold_ok(const cn_conv_desc* d) {
  if (d->upsample != 2 || d->stride != 1 || d->pad >= 0 || (d->nd != 2 && d->nd != 3)) return false;
  for (int i = 0; i < d->nd; ++i) if (d->ksize[i] < 2 || d->ksize[i] > 5) return false;
  return true;
}

struct AxisOpt { int off, lo, hi; };   // source offset (forward) or gradient offset l (dgrad), tap range

static int build_fold_plan(const cn_conv_desc* d, int kind, GemmPlan* out, std::vector<int2>& taps, FoldInfo* fi) {
  int U[3], O[3], pb[3];
  same_geometry(d, U, O, pb);
  GemmPlan g;
  memset(&g, 0, sizeof(g));
  g.n_img = d->batch;
  taps.clear();
  fi->spec.clear();
  const int cc = d->cin * d->cout;
  if (kind == KIND_FWD_FOLD) {
    for (int i = 0; i < 3; ++i) {
      g.E[i] = d->in_dims[i]; g.U[i] = d->in_dims[i]; g.S[i] = d->in_dims[i]; g.Q[i] = O[i]; g.ooff[i] = 0;
    }
    g.mstride = 1; g.ushift = 0; g.ostride = 2;
    g.Csrc = d->cin; g.Cn = d->cout; g.wsc = d->cout; g.wsn = 1;
    const int nph = 1 << d->nd;
    struct Ph { int bits; std::vector<int2> taps; std::vector<int> spec; };
    std::vector<Ph> phases;
    for (int ph = 0; ph < nph; ++ph) {
      std::vector<AxisOpt> opt[3];
      for (int a = 0; a < 3; ++a) {
        if (a >= d->nd) { opt[a].push_back({0, 0, 0}); continue; }
        const int dd = (ph >> a) & 1;
        for (int t = 0; t < d->ksize[a]; ++t) {
          const int so = floordiv2(dd + t - pb[a]);
          if (!opt[a].empty() && opt[a].back().off == so) opt[a].back().hi = t;
          else opt[a].push_back({so, t, t});
        }
      }
      Ph P; P.bits = ph;
      for (auto& a0 : opt[0]) for (auto& a1 : opt[1]) for (auto& a2 : opt[2]) {
        int off[3] = {a0.off, a1.off, a2.off}, lo[3] = {a0.lo, a1.lo, a2.lo}, hi[3] = {a0.hi, a1.hi, a2.hi};
        P.taps.push_back(make_int2(pack_off(off), 0));
        P.spec.push_back(fold_encode(lo, hi));
      }
      phases.push_back(P);
    }
    // longest phases first: CTAs are dispatched in blockIdx order
    std::stable_sort(phases.begin(), phases.end(), [](const Ph& a, const Ph& b) { return a.taps.size() > b.taps.size(); });
    g.nphase = nph;
    int maxt = 0;
    for (int ph = 0; ph < nph; ++ph) {
      g.ph_tap0[ph] = (int)taps.size(); g.ph_ntaps[ph] = (int)phases[ph].taps.size(); g.ph_ooff[ph] = phases[ph].bits;
      if (g.ph_ntaps[ph] > maxt) maxt = g.ph_ntaps[ph];
      for (size_t i = 0; i < phases[ph].taps.size(); ++i) {
        int2 t = phases[ph].taps[i];
        t.y = (int)fi->spec.size() * cc;
        taps.push_back(t);
        fi->spec.push_back(phases[ph].spec[i]);
      }
    }
    g.ntaps = maxt;
  } else {   // KIND_DGRAD_FOLD: rows = low-resolution x pixels, source = gy at 2r + l
    for (int i = 0; i < 3; ++i) {
      g.E[i] = d->in_dims[i]; g.U[i] = O[i]; g.S[i] = O[i]; g.Q[i] = d->in_dims[i]; g.ooff[i] = 0;
    }
    g.mstride = 2; g.ushift = 0; g.ostride = 1;
    g.Csrc = d->cout; g.Cn = d->cin; g.wsc = 1; g.wsn = d->cout;
    std::vector<AxisOpt> opt[3];
    for (int a = 0; a < 3; ++a) {
      if (a >= d->nd) { opt[a].push_back({0, 0, 0}); continue; }
      for (int l = pb[a] - (d->ksize[a] - 1); l <= pb[a] + 1; ++l) {
        int lo = pb[a] - l, hi = pb[a] - l + 1;
        if (lo < 0) lo = 0;
        if (hi > d->ksize[a] - 1) hi = d->ksize[a] - 1;
        if (lo <= hi) opt[a].push_back({l, lo, hi});
      }
    }
    for (auto& a0 : opt[0]) for (auto& a1 : opt[1]) for (auto& a2 : opt[2]) {
      int off[3] = {a0.off, a1.off, a2.off}, lo[3] = {a0.lo, a1.lo, a2.lo}, hi[3] = {a0.hi, a1.hi, a2.hi};
      taps.push_back(make_int2(pack_off(off), (int)fi->spec.size() * cc));
      fi->spec.push_back(fold_encode(lo, hi));
    }
    g.ntaps = (int)taps.size();
    g.nphase = 1;
    // unfold lists: original tap t -> the folded taps whose ranges contain it
    const int kvol = d->ksize[0] * d->ksize[1] * d->ksize[2];
    fi->unfold.assign((size_t)kvol * 8, -1);
    for (int t = 0; t < kvol; ++t) {
      int tt[3] = {t / (d->ksize[2] * d->ksize[1]), (t / d->ksize[2]) % d->ksize[1], t % d->ksize[2]};
      int cnt = 0;
      for (size_t f = 0; f < fi->spec.size(); ++f) {
        int lo[3], hi[3]; fold_decode(fi->spec[f], lo, hi);
        bool in = true;
        for (int a = 0; a < 3; ++a) in = in && tt[a] >= lo[a] && tt[a] <= hi[a];
        if (in) { CN_REQUIRE(cnt < 8, CN_ERR_UNSUPPORTED, "unfold list overflow"); fi->unfold[(size_t)t * 8 + cnt++] = (int)f; }
      }
    }
  }
  fi->nfold = (int)fi->spec.size();
  CN_REQUIRE(taps.size() <= 256, CN_ERR_UNSUPPORTED, "too many folded taps (%d)", (int)taps.size());
  for (size_t i = 0; i < taps.size(); ++i) {
    int px = taps[i].x;
    int o0 = (px & 1023), o1 = (px >> 10) & 1023, o2 = (px >> 20) & 1023;
    CN_REQUIRE(o0 < 24 && o1 < 24 && o2 < 24, CN_ERR_UNSUPPORTED, "tap offset out of packed range");
  }
  g.Ktot = g.ntaps * g.Csrc;
  g.kb_stride = (g.Ktot + 31) / 32;
  g.taps_total = (int)taps.size();
  long long M = (long long)g.n_img * g.E[0] * g.E[1] * g.E[2];
  CN_REQUIRE(M < (1ll << 31), CN_ERR_UNSUPPORTED, "too many output rows");
  g.M = (int)M;
  CN_REQUIRE((long long)g.n_img * g.S[0] * g.S[1] * g.S[2] * g.Csrc < (1ll << 32) &&
             (long long)g.n_img * g.Q[0] * g.Q[1] * g.Q[2] * g.Cn < (1ll << 32), CN_ERR_UNSUPPORTED, "tensor too large for 32-bit element offsets");
  g.taps = nullptr;
  *out = g;
  return CN_OK;
}

// All parity phases of a stride-2 input gradient as one phased plan (even input dims, every phase has taps).
static int build_s2all_plan(const cn_conv_desc* d, GemmPlan* out, std::vector<int2>& taps) {
  const int nph = 1 << d->nd;
  struct Ph { int bits; std::vector<int2> taps; };
  std::vector<Ph> phases;
  GemmPlan g0;
  for (int ph = 0; ph < nph; ++ph) {
    GemmPlan g; std::vector<int2> t;
    int rc = build_plan(d, KIND_DGRAD, ph, &g, t); if (rc) return rc;
    if (ph == 0) g0 = g;
    CN_REQUIRE(!t.empty() && g.M == g0.M && g.E[0] == g0.E[0] && g.E[1] == g0.E[1] && g.E[2] == g0.E[2], CN_ERR_UNSUPPORTED,
               "stride-2 phases are not uniform");
    Ph P; P.bits = ph; P.taps = t;
    phases.push_back(P);
  }
  std::stable_sort(phases.begin(), phases.end(), [](const Ph& a, const Ph& b) { return a.taps.size() > b.taps.size(); });
  taps.clear();
  GemmPlan g = g0;
  g.nphase = nph;
  int maxt = 0;
  for (int ph = 0; ph < nph; ++ph) {
    g.ph_tap0[ph] = (int)taps.size(); g.ph_ntaps[ph] = (int)phases[ph].taps.size(); g.ph_ooff[ph] = phases[ph].bits;
    if (g.ph_ntaps[ph] > maxt) maxt = g.ph_ntaps[ph];
    taps.insert(taps.end(), phases[ph].taps.begin(), phases[ph].taps.end());
  }
  CN_REQUIRE(taps.size() <= 256, CN_ERR_UNSUPPORTED, "too many taps");
  g.ntaps = maxt; g.Ktot = maxt * g.Csrc; g.kb_stride = (g.Ktot + 31) / 32;
  g.taps_total = (int)taps.size();
  g.taps = nullptr;
  *out = g;
  return CN_OK;
}

static int get_special_plan(const cn_conv_desc* d, int kind, GemmPlan* out, FoldInfo** fold) {
  std::string key((const char*)d, sizeof(*d));
  key.push_back((char)kind);
  key.push_back((char)0);
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    if (!it->second.valid) return CN_ERR_UNSUPPORTED;
    *out = it->second.g;
    if (fold) *fold = &g_folds[key];
    return CN_OK;
  }
  GemmPlan g;
  std::vector<int2> taps;
  FoldInfo fi;
  int rc = (kind == KIND_DGRAD_S2ALL) ? build_s2all_plan(d, &g, taps) : build_fold_plan(d, kind, &g, taps, &fi);
  if (rc) { HostPlan hp; memset(&hp.g, 0, sizeof(hp.g)); hp.valid = false; g_plans[key] = hp; return rc; }
  int2* dtaps = nullptr;
  CN_CHECK_CUDA(cudaMalloc(&dtaps, taps.size() * sizeof(int2)));
  CN_CHECK_CUDA(cudaMemcpy(dtaps, taps.data(), taps.size() * sizeof(int2), cudaMemcpyHostToDevice));
  g.taps = dtaps;
  if (fi.nfold > 0) {
    CN_CHECK_CUDA(cudaMalloc(&fi.d_spec, fi.spec.size() * sizeof(int)));
    CN_CHECK_CUDA(cudaMemcpy(fi.d_spec, fi.spec.data(), fi.spec.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!fi.unfold.empty()) {
      CN_CHECK_CUDA(cudaMalloc(&fi.d_unfold, fi.unfold.size() * sizeof(int)));
      CN_CHECK_CUDA(cudaMemcpy(fi.d_unfold, fi.unfold.data(), fi.unfold.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
  }
  g_folds[key] = fi;
  HostPlan hp; hp.g = g; hp.valid = true;
  g_plans[key] = hp;
  *out = g;
  if (fold) *fold = &g_folds[key];
  return CN_OK;
}

// Wf[f][ci][co] = sum of w[t][ci][co] over the tap box of folded tap f.  grid (ceil(cc/4/256), nfold)
__global__ void __launch_bounds__(256)
fold_weights_kernel(const float* __restrict__ w, const int* __restrict__ spec, int k1, int k2, int cc4, float* __restrict__ out) {
  const int f = blockIdx.y;
  int lo[3], hi[3];
  fold_decode(spec[f], lo, hi);
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= cc4) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t0 = lo[0]; t0 <= hi[0]; ++t0)
    for (int t1 = lo[1]; t1 <= hi[1]; ++t1)
      for (int t2 = lo[2]; t2 <= hi[2]; ++t2) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(w) + (size_t)((t0 * k1 + t1) * k2 + t2) * cc4 + i);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
  reinterpret_cast<float4*>(out)[(size_t)f * cc4 + i] = a;
}

// gw[t][ci][co] = sum_{f in unfold[t]} Dp[(f*cout + co)*cin + ci]   (32x32 transpose tiles; grid (cin/32, cout/32, kvol), block (32, 8))
__global__ void unfold_wgrad_kernel(const float* __restrict__ Dp, const int* __restrict__ unfold, int cin, int cout, float* __restrict__ gw) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z, ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  int fl[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) fl[j] = unfold[t * 8 + j];
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    float a = 0.f;
    if (co < cout && ci < cin) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (fl[j] >= 0) a += Dp[((size_t)fl[j] * cout + co) * cin + ci];
    }
    tile[r][threadIdx.x] = a;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    if (ci < cin && co < cout) gw[((size_t)t * cin + ci) * cout + co] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------------------------------
// Device helpers shared by both kernel families
// ------------------------------------------------------------------------------------------------
struct RowInfo {
  uint32_t base;   // n_img * S0*S1*S2 (source pixel index of the sample), 0xffffffff = row out of range
  uint32_t pk;     // packed e_d*mstride
};

__host__ __device__ __forceinline__ RowInfo decode_row(const GemmPlan& p, int m) {
  RowInfo r;
  if (m >= p.M) { r.base = 0xffffffffu; r.pk = 0; return r; }
  int e2 = m % p.E[2]; m /= p.E[2];
  int e1 = m % p.E[1]; m /= p.E[1];
  int e0 = m % p.E[0]; m /= p.E[0];
  r.base = (uint32_t)m * (uint32_t)(p.S[0] * p.S[1] * p.S[2]);
  r.pk = (uint32_t)(e0 * p.mstride) | ((uint32_t)(e1 * p.mstride) << 10) | ((uint32_t)(e2 * p.mstride) << 20);
  return r;
}

// destination pixel index of row m
__host__ __device__ __forceinline__ uint32_t dest_pixel(const GemmPlan& p, int m) {
  int e2 = m % p.E[2]; m /= p.E[2];
  int e1 = m % p.E[1]; m /= p.E[1];
  int e0 = m % p.E[0]; m /= p.E[0];
  uint32_t q0 = e0 * p.ostride + p.ooff[0], q1 = e1 * p.ostride + p.ooff[1], q2 = e2 * p.ostride + p.ooff[2];
  return (((uint32_t)m * p.Q[0] + q0) * p.Q[1] + q1) * p.Q[2] + q2;
}

// source pixel index for (row, tap) or 0xffffffff when the tap falls into the padding
__host__ __device__ __forceinline__ uint32_t src_pixel(const GemmPlan& p, RowInfo r, int tap_pk) {
  if (r.base == 0xffffffffu) return 0xffffffffu;
  uint32_t s = r.pk + (uint32_t)tap_pk;
  int u0 = (int)(s & 1023u) - 8, u1 = (int)((s >> 10) & 1023u) - 8, u2 = (int)(s >> 20) - 8;
  bool ok = (unsigned)u0 < (unsigned)p.U[0] && (unsigned)u1 < (unsigned)p.U[1] && (unsigned)u2 < (unsigned)p.U[2];
  if (!ok) return 0xffffffffu;
  return r.base + (uint32_t)(((u0 >> p.ushift) * p.S[1] + (u1 >> p.ushift)) * p.S[2] + (u2 >> p.ushift));
}

// ------------------------------------------------------------------------------------------------
// CUDA-core implicit GEMM (fp32 FFMA)
// ------------------------------------------------------------------------------------------------
enum { MODE_PIXEL = 0, MODE_WGRAD = 1 };

template <int MODE, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
igemm_ffma_kernel(GemmPlan p, const float* __restrict__ A, const float* __restrict__ B,
                  const float* __restrict__ bias, float* __restrict__ D, int act, float alpha,
                  int kchunk, long long part_stride) {
  // part_stride != 0: split-K - split blockIdx.z writes its raw partial sums into slab blockIdx.z of D (a workspace of
  // gridDim.z slabs of part_stride floats); splitk_reduce_kernel adds the slabs in split order, then bias + activation.
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int BK = 16;
  static_assert(NT % BM == 0 || BM % NT == 0, "row mapping");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int2 s_taps[256];
  const int tid = threadIdx.x;
  for (int i = tid; i < p.ntaps; i += NT) s_taps[i] = p.taps[i];
  __syncthreads();
  const int Mg = (MODE == MODE_PIXEL) ? p.M : p.Ktot;
  const int Ng = p.Cn;
  const int Kg = (MODE == MODE_PIXEL) ? p.Ktot : p.M;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(Kg, kbeg + kchunk);

  // per-thread fixed A row
  const int arow = tid % BM;
  RowInfo rinfo; rinfo.base = 0xffffffffu; rinfo.pk = 0;
  int a_c = 0, a_tap_pk = 0; bool a_row_ok = false;
  if (MODE == MODE_PIXEL) {
    rinfo = decode_row(p, m0 + arow);
  } else {
    int r = m0 + arow;
    a_row_ok = r < Mg;
    if (a_row_ok) { int kt = r / p.Csrc; a_c = r - kt * p.Csrc; a_tap_pk = s_taps[kt].x; }
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int ty = tid / (BN / TN), tx = tid % (BN / TN);

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile
    for (int e = tid; e < BM * BK; e += NT) {
      int kk = e / BM;
      int k = k0 + kk;
      float v = 0.f;
      if (k < kend) {
        if (MODE == MODE_PIXEL) {
          int kt = k / p.Csrc; int c = k - kt * p.Csrc;
          uint32_t sp = src_pixel(p, rinfo, s_taps[kt].x);
          if (sp != 0xffffffffu) v = __ldg(A + (size_t)sp * p.Csrc + c);
        } else if (a_row_ok) {
          RowInfo ri = decode_row(p, k);
          uint32_t sp = src_pixel(p, ri, a_tap_pk);
          if (sp != 0xffffffffu) v = __ldg(A + (size_t)sp * p.Csrc + a_c);
        }
      }
      As[kk][arow] = v;
    }
    // ---- B tile
    for (int e = tid; e < BN * BK; e += NT) {
      int col = e % BN, kk = e / BN;
      int k = k0 + kk, n = n0 + col;
      float v = 0.f;
      if (k < kend && n < Ng) {
        if (MODE == MODE_PIXEL) {
          int kt = k / p.Csrc; int c = k - kt * p.Csrc;
          v = __ldg(B + (size_t)s_taps[kt].y + (size_t)c * p.wsc + (size_t)n * p.wsn);
        } else {
          v = __ldg(B + (size_t)k * p.Cn + n);
        }
      }
      Bs[kk][col] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= Mg) continue;
    size_t rowoff = (MODE == MODE_PIXEL) ? (size_t)dest_pixel(p, m) * p.Cn : (size_t)m * p.Cn;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= Ng) continue;
      float v = acc[i][j];
      if (part_stride) {
        D[(size_t)blockIdx.z * (size_t)part_stride + rowoff + n] = v;
      } else {
        if (bias != nullptr) v += bias[n];
        D[rowoff + n] = cn_apply_act(v, act, alpha);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 helpers (inline PTX)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// waits for two barriers at once: both polls are in flight together (one round trip instead of two when both are complete)
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
  uint32_t da, db;
  do {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "selp.u32 %1, 1, 0, q;\n\t}"
        : "=r"(da), "=r"(db) : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b) : "memory");
  } while (!(da & db));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, async proxy), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 operands, fp32 accumulate, issued by ONE thread for the CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 3xTF32 split: x = big + small, both exactly representable in tf32 (built by masking the 13 low mantissa bits, see
// sts_split4; f2tf32 is the cvt.rna form the first version used, kept for experiments).  Residual ~2^-21 |x|.
// small*big + big*small + big*big on the tf32 tensor pipe reproduces the fp32 product to ~5e-7, which the
// ill-conditioned gradients of this network need: measured on B200, single tf32 is 3.6e-2 off on the
// generator output and a bf16 hi/lo split (~1e-5 per product) 1e-2..5e-2 off on some parameter gradients.
__device__ __forceinline__ uint32_t f2tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ void sts_split4(uint32_t addr_big, uint32_t addr_small, float4 v) {
  // big = x with the 13 low mantissa bits cleared (exactly a tf32), small = tf32-truncated (x - big): the
  // subtraction is exact, so x = big + small up to 2^-21 |x|; 3 ALU ops per element instead of 2 cvt + 1 sub.
  const float x[4] = {v.x, v.y, v.z, v.w};
  uint32_t bg[4], sm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bg[i] = __float_as_uint(x[i]) & 0xffffe000u;
    sm[i] = __float_as_uint(x[i] - __uint_as_float(bg[i])) & 0xffffe000u;
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr_big), "r"(bg[0]), "r"(bg[1]), "r"(bg[2]), "r"(bg[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr_small), "r"(sm[0]), "r"(sm[1]), "r"(sm[2]), "r"(sm[3]) : "memory");
}

// UMMA shared-memory descriptor, version 1 (sm_100).  addr/lbo/sbo in bytes.
//   layout 2 = SWIZZLE_128B, K-major tiles: rows of 128 B (32 tf32 of K), 8-row groups SBO = 1024 B apart.
//   layout 1 = SWIZZLE_128B_BASE32B, the only MN-major layout for tf32: atoms of 4 k-rows x 128 B (32 tf32
//              along M/N), 32-byte units XOR-ed with the k-row index; LBO = stride between atoms along M/N,
//              SBO = stride between groups of 4 k-rows.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor: tf32 x tf32 -> f32, M=128
__host__ __device__ inline uint32_t umma_idesc_tf32(int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format f32
  d |= 2u << 7;                      // A format tf32
  d |= 2u << 10;                     // B format tf32
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}
// byte offset of the 16-byte chunk (k-row r, chunk jn = 4 fp32 along M/N) inside an MN-major tf32 tile whose
// k-groups (4 rows) are `sbo` bytes apart and whose 32-column atoms (512 B) are contiguous
__device__ __forceinline__ uint32_t mn_chunk_off(int r, int jn, uint32_t sbo) {
  return (uint32_t)(r >> 2) * sbo + (uint32_t)(jn >> 3) * 512u + (uint32_t)(r & 3) * 128u +
         ((uint32_t)(((jn & 7) >> 1) ^ (r & 3)) << 5) + ((uint32_t)(jn & 1) << 4);
}

constexpr int TC_BM = 128;        // GEMM rows per CTA (UMMA M)
constexpr int TC_BK = 32;         // K elements per stage = one 128-byte swizzle row of tf32
constexpr int TC_CHUNK_KB = 8;    // k-blocks (256 K elements) accumulated in the tensor core before promotion
// The running total of the chunked promotion lives in shared memory as ceil(bn / 32) chunks of [128 rows][32 columns]
// fp32 with the 16-byte units of a row XOR-ed with (row & 7): exactly the SWIZZLE_128B box layout of a TMA tensor store
// (the final tile leaves through cp.async.bulk.tensor straight from here), conflict-free for the promotion's per-row
// 16-byte accesses (a quarter warp = 8 consecutive rows = 8 different units) and for the channel-major pass (one row =
// one 128-byte line).
constexpr int TC_TOT_CHUNK = 128 * 128;   // bytes per 32-column chunk
__device__ __forceinline__ uint32_t tot_unit_off(int row, int unit) {      // byte offset of 16-byte unit `unit` (0..7) of `row` inside a chunk
  return (uint32_t)row * 128u + ((uint32_t)(unit ^ (row & 7)) << 4);
}

// The tensor core adds into its fp32 accumulator with truncation: measured on B200 the result drifts by
// ~1.1e-8 * K relative (1.2e-4 at K = 18432), a bias that plain fp32 FMAs do not have.  So the MMA warp
// accumulates at most TC_CHUNK_KB k-blocks into one of two ping-pong TMEM accumulators and the promotion
// warps add each finished chunk into a running fp32 total (round-to-nearest FADD on the CUDA cores) while the
// tensor core works on the next chunk.  store(cb, v) receives the final 32-column groups of this thread's
// accumulator row.  The running total lives in shared memory ([column][128 rows] fp32: lane = row, so the
// accesses are conflict-free), which leaves the tensor memory to the two accumulators and the A stages.
#ifdef CN_TC_COMPACT_ACT
// Activation of the tcgen05 epilogue in three instructions per element: none / LeakyReLU / ReLU are one select on a
// per-launch negative-side slope (1, alpha, 0; the fused add of +0 turns ReLU's -0 into +0, as v > 0 ? v : 0 gives);
// tanh (no tensor-core layer of the training step uses it) stays out of line.  The epilogue unrolls this 64 times.
__device__ __noinline__ float tc_tanh_out_of_line(float v) { return tanhf(v); }
__device__ __forceinline__ float tc_act(float v, float slope, bool is_tanh) {
  if (is_tanh) return tc_tanh_out_of_line(v);
  return v >= 0.f ? v : fmaf(v, slope, 0.f);
}
#endif
template <class StoreFn, class KeepFn>
__device__ __forceinline__ void tc_promote_smem_and_store(uint32_t tmem_base, uint8_t* tot, int pw, int bn, int bn_r, int c0, int nchunks,
                                                          uint32_t bar_accfull, uint32_t bar_accempty, bool keep_last, StoreFn store, KeepFn keep) {
  // c0 = accumulation chunks this CTA has consumed before this item (persistent CTAs): the ping-pong accumulator
  // and the barrier phases follow the GLOBAL chunk index; every chunk, the last of an item included, releases its
  // accumulator so that a later item can reuse it.
  // keep_last: the final sums go back to `tot` after keep(cb, v) has finished them (bias, activation) - the caller then
  // writes the tile from shared memory (TMA tensor store, or lanes along the channels); else store(cb, v) writes them.
  const uint32_t lanebits = (uint32_t)(pw * 32) << 16;
  const int row = pw * 32 + (threadIdx.x & 31);
  for (int c = 0; c < nchunks; ++c) {
    const int gc = c0 + c;
    const int b = gc & 1;
    mbar_wait(bar_accfull + 8 * b, (gc >> 1) & 1);
    tc_fence_after();
    const bool last = c == nchunks - 1;
    for (int cb = 0; cb < bn; cb += 32) {
      uint32_t v[32];
      tc_ld32(tmem_base + lanebits + b * bn_r + cb, v);
      uint8_t* chunk = tot + (cb >> 5) * TC_TOT_CHUNK;
      if (c > 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 t = *reinterpret_cast<const float4*>(chunk + tot_unit_off(row, u));
          v[4 * u] = __float_as_uint(__uint_as_float(v[4 * u]) + t.x); v[4 * u + 1] = __float_as_uint(__uint_as_float(v[4 * u + 1]) + t.y);
          v[4 * u + 2] = __float_as_uint(__uint_as_float(v[4 * u + 2]) + t.z); v[4 * u + 3] = __float_as_uint(__uint_as_float(v[4 * u + 3]) + t.w);
        }
      }
      if (last && !keep_last) { store(cb, v); continue; }
      if (last) keep(cb, v);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<float4*>(chunk + tot_unit_off(row, u)) =
            make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
    }
    tc_fence_before(); __syncwarp(); if ((threadIdx.x & 31) == 0) mbar_arrive(bar_accempty + 8 * b);
  }
}

// TMA tensor store of one 32-column chunk of the finished tile: box {32 channels, 128 rows} at (column c0, row r0) of the
// layer's output seen as a (rows, channels) matrix; rows / columns beyond the tensor are clipped by the unit.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int r0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(smem_src) : "memory");
}
// The same chunk of a stride-2 PHASE tile (parity phases of the stride-2 input gradient, sub-pixel phases of the folded 2-D
// upsample + conv): the output seen as (channels, x parity, x / 2, y parity, sample * H / 2 + y / 2); the 128 GEMM rows are
// box {32, 1, bw, 1, 128 / bw} of it, in exactly the row order of the running-total buffer.
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t smem_src, int c0, int px, int x0, int py, int r0) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(px), "r"(x0), "r"(py), "r"(r0), "r"(smem_src) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }     // the 4 promotion / epilogue warps

// Weight pre-pack: writes, for every (n-tile, k-block), the exact shared-memory image of the B stage
// (big tile then small tile, swizzled) so that the conv kernel fetches a whole B stage with ONE bulk copy
// issued by one thread.  The split into tf32 big/small parts happens here once per layer call instead of
// once per M-tile.  grid = (k-blocks, n-tiles), 256 threads.
template <int B_MN>
__global__ void __launch_bounds__(256)
pack_weights_kernel(const GemmPlan pin, const float* __restrict__ W, float* __restrict__ out, int bn, int bn_smem, int total_kb) {
  __shared__ int2 s_taps[256];
  GemmPlan p = pin;
  const int ntile = cn_select_phase(p, blockIdx.y, gridDim.y);     // total_kb = slot stride of the buffer
  if ((int)blockIdx.x * TC_BK >= p.Ktot) return;                    // beyond this phase's K: the slot is never read
  for (int i = threadIdx.x; i < p.ntaps; i += 256) s_taps[i] = p.taps[i];
  __syncthreads();
  const int kb = blockIdx.x, nt = blockIdx.y;
  const uint32_t b_bytes = (uint32_t)bn_smem * TC_BK * 4;
  uint8_t* tile = reinterpret_cast<uint8_t*>(out) + ((size_t)nt * total_kb + kb) * 2 * b_bytes;
  const int n0 = ntile * bn, k0 = kb * TC_BK;
  const int nchunks = 8 * bn_smem;
  const uint32_t sbo = (uint32_t)(bn_smem >> 5) * 512u;
  for (int q = threadIdx.x; q < nchunks; q += 256) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t off;
    if (B_MN) {
      const int cpr = bn_smem >> 2;
      const int r = q / cpr, jc = q - r * cpr;
      const int k = k0 + r, n = n0 + 4 * jc;
      if (k < p.Ktot && n < p.Cn && 4 * jc < bn) {
        const int kt = k / p.Csrc, c = k - kt * p.Csrc;
        v = ldg128(W + (size_t)s_taps[kt].y + (size_t)c * p.wsc + n);
      }
      off = mn_chunk_off(r, jc, sbo);
    } else {
      const int r = q >> 3, j = q & 7;
      const int k = k0 + 4 * j, n = n0 + r;
      if (k < p.Ktot && r < bn && n < p.Cn) {
        const int kt = k / p.Csrc, c = k - kt * p.Csrc;
        v = ldg128(W + (size_t)s_taps[kt].y + (size_t)n * p.wsn + c);
      }
      off = (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4));
    }
    const float x[4] = {v.x, v.y, v.z, v.w};
    uint4 bg, sm;
    uint32_t* pb = reinterpret_cast<uint32_t*>(&bg);
    uint32_t* ps = reinterpret_cast<uint32_t*>(&sm);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pb[i] = __float_as_uint(x[i]) & 0xffffe000u;
      ps[i] = __float_as_uint(x[i] - __uint_as_float(pb[i])) & 0xffffe000u;
    }
    *reinterpret_cast<uint4*>(tile + off) = bg;
    *reinterpret_cast<uint4*>(tile + b_bytes + off) = sm;
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 pixel-mode kernel (forward: B MN-major = Keras kernel as is; dgrad: B K-major)
//
// Measured on B200 (profiles/r01_layer_time_dbg.txt): with one CTA per SM every k-block pulls 16 KB of
// activations and 32 KB of pre-split weights through L2, and at ~6.4 TB/s of L2->SM bandwidth that, not the
// tensor pipe, bounds the kernel.  So the CTAs of a thread-block cluster (consecutive M tiles of the same
// N tile) share the weight stream: each CTA fetches 1/C of every B stage and multicasts it to all C CTAs.
// Warp roles (448 threads): 0-7 gather A, 8-11 promote + epilogue, 12 MMA issue, 13 B producer.
// ------------------------------------------------------------------------------------------------
// true in exactly one (converged) lane of the warp; ptxas recognises the pattern and issues the guarded
// tcgen05 instructions without a per-operand uniformity loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

constexpr int TCP_THREADS = 448;
constexpr int TCP_MMA_WARP = 12;
constexpr int TCP_B_WARP = 13;
constexpr int TCP_A_COLS = 64;       // TMEM columns of one A stage: 32 big + 32 small

// D[tmem] (+)= A[tmem] * B[smem]: A is the 128 x 8 tf32 block at TMEM address `tmem_a` (lane = GEMM row,
// column = k), K-major by construction
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_st16_nowait(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

constexpr int TCP_MAX_A = 6;         // A stages in tensor memory (as many as fit beside the two accumulators)
// Lanes per GEMM row in the pixel-mode gather (1, 2, 4 or 8).  LPR lanes read adjacent 16-byte chunks of one row, so a load
// instruction touches 32 / LPR lines of 128 B instead of 32 (the L1 data pipe pays per line touched and was 90 % busy with
// LPR = 1, profiles/r01_ncu_full_tc_final_summary.txt); an LPR x LPR register <-> lane transpose (log2 LPR butterfly stages
// of shuffles) then gives every thread the eight chunks of ITS row.  More lanes per row = fewer L1 wavefronts but more
// shuffle / select instructions in the gather warps; measured per value in profiles/r02_gather_lpr_ab.txt.
#ifndef CN_LPR
#define CN_LPR 2      // measured best on every fwd / dgrad layer of the bench step but two (1 / 2 / 4 / 8: 33.3 / 30.9 / 32.2 / 40.0 ms per step)
#endif
constexpr int TCP_LPR = CN_LPR;
constexpr int TCP_HPR = 8 / TCP_LPR;             // load instructions per row group (chunk groups of LPR chunks)
struct TcpLayout {
  // dynamic smem, 1024-byte aligned: nb B stages [B big][B small]; running total [bn_r][128] fp32; barriers, tmem ptr, taps
  uint32_t stage_bytes, b_bytes, tot_off, bar_off, tmem_off, taps_off, total;
};
__host__ __device__ inline TcpLayout tcp_layout(int nb, int bn_smem) {
  TcpLayout l;
  l.b_bytes = bn_smem * TC_BK * 4;
  l.stage_bytes = 2 * l.b_bytes;
  l.tot_off = nb * l.stage_bytes;
  l.bar_off = l.tot_off + ((bn_smem + 31) / 32) * TC_TOT_CHUNK;
  l.tmem_off = l.bar_off + (2 * nb + 2 * TCP_MAX_A + 4) * 8;      // full_b[], empty_b[], full_a[], empty_a[], acc_full[2], acc_empty[2]
  l.taps_off = (l.tmem_off + 4 + 7) & ~7u;
  l.total = l.taps_off + 256 * 8;
  return l;
}

// Gradient pre-pack for the wgrad GEMM: the B operand tile of (n-tile, k-block) is 32 consecutive output pixels
// x bn_smem channels of G, MN-major (channels contiguous, as in HBM).  Same stage image format as above.
// colpart != nullptr: the block also writes the column sums of its 32 x bn piece of G to colpart[kb][n] - the bias
// gradient's partial sums come out of the pass that reads G anyway (sum_slabs_kernel adds the k-blocks in order).
__global__ void __launch_bounds__(256)
pack_grad_kernel(const float* __restrict__ G, int M, int Cn, float* __restrict__ out, int bn, int bn_smem, int total_kb,
                 float* __restrict__ colpart) {
  const int kb = blockIdx.x, nt = blockIdx.y;
  const uint32_t b_bytes = (uint32_t)bn_smem * TC_BK * 4;
  uint8_t* tile = reinterpret_cast<uint8_t*>(out) + ((size_t)nt * total_kb + kb) * 2 * b_bytes;
  const int n0 = nt * bn, m0 = kb * TC_BK;
  const int cpr = bn_smem >> 2, nchunks = 8 * bn_smem;
  const uint32_t sbo = (uint32_t)(bn_smem >> 5) * 512u;
  for (int q = threadIdx.x; q < nchunks; q += 256) {
    const int r = q / cpr, jc = q - r * cpr;
    const int m = m0 + r, n = n0 + 4 * jc;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M && n < Cn && 4 * jc < bn) v = ldg128(G + (size_t)m * Cn + n);
    const uint32_t off = mn_chunk_off(r, jc, sbo);
    const float x[4] = {v.x, v.y, v.z, v.w};
    uint4 bg, sm;
    uint32_t* pb = reinterpret_cast<uint32_t*>(&bg);
    uint32_t* ps = reinterpret_cast<uint32_t*>(&sm);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pb[i] = __float_as_uint(x[i]) & 0xffffe000u;
      ps[i] = __float_as_uint(x[i] - __uint_as_float(pb[i])) & 0xffffe000u;
    }
    *reinterpret_cast<uint4*>(tile + off) = bg;
    *reinterpret_cast<uint4*>(tile + b_bytes + off) = sm;
  }
  if (colpart != nullptr) {
    // the 32 rows were just read (they sit in L1 / L2): thread t < bn adds column n0 + t over them in row order
    const int t = threadIdx.x, n = n0 + t;
    if (t < bn && n < Cn) {
      float a = 0.f;
      const int rows = min(TC_BK, M - m0);
      for (int r = 0; r < rows; ++r) a += __ldg(G + (size_t)(m0 + r) * Cn + n);
      colpart[(size_t)kb * Cn + n] = a;
    }
  }
}

// wgrad A gather of one k-block (32 consecutive output pixels starting at (xn, x0, x1, x2)) for one (tap, channel)
// row.  SEG = pixels per image-row segment inside the k-block: 32 when the row length E2 is a multiple of 32
// (the k-block stays inside one image row), else SEG = E2 in {4, 8, 16} (32/SEG whole rows).  Straight-line
// code per variant: the fully general per-pixel carry version is ~30 KB of SASS per call site and thrashes the
// instruction cache (measured: 340 clk per pixel).
struct WgGeom {                // plan geometry with the image row in position 2 (2-D plans carry it in position 1)
  int E[3], U[3], S[3], n_img, mstride, ushift, Csrc;
};
__device__ __forceinline__ WgGeom wg_geom(const GemmPlan& p) {
  WgGeom g;
  const bool shift = p.E[2] == 1 && p.U[2] == 1 && p.S[2] == 1;
  g.E[0] = shift ? 1 : p.E[0]; g.E[1] = shift ? p.E[0] : p.E[1]; g.E[2] = shift ? p.E[1] : p.E[2];
  g.U[0] = shift ? 1 : p.U[0]; g.U[1] = shift ? p.U[0] : p.U[1]; g.U[2] = shift ? p.U[1] : p.U[2];
  g.S[0] = shift ? 1 : p.S[0]; g.S[1] = shift ? p.S[0] : p.S[1]; g.S[2] = shift ? p.S[1] : p.S[2];
  g.n_img = p.n_img; g.mstride = p.mstride; g.ushift = p.ushift; g.Csrc = p.Csrc;
  return g;
}
template <int SEG>
__device__ __forceinline__ void wg_load(const WgGeom& p, const float* __restrict__ Ac, bool wrok, int woff0, int woff1, int woff2,
                                        int xn, int x0, int x1, int x2, uint32_t* v) {
#pragma unroll
  for (int s = 0; s < TC_BK / SEG; ++s) {
    const int u0 = x0 * p.mstride + woff0, u1 = x1 * p.mstride + woff1;
    const bool rowok = wrok && xn < p.n_img && (unsigned)u0 < (unsigned)p.U[0] && (unsigned)u1 < (unsigned)p.U[1];
    const uint32_t rowbase = (((uint32_t)xn * p.S[0] + (uint32_t)(u0 >> p.ushift)) * p.S[1] + (uint32_t)(u1 >> p.ushift)) * p.S[2];
    int u2 = x2 * p.mstride + woff2;
#pragma unroll
    for (int jj = 0; jj < SEG; ++jj) {
      const bool ok = rowok && (unsigned)u2 < (unsigned)p.U[2];
      v[s * SEG + jj] = ok ? __float_as_uint(__ldg(Ac + (size_t)(rowbase + (uint32_t)(u2 >> p.ushift)) * p.Csrc)) : 0u;
      u2 += p.mstride;
    }
    if (SEG == 1) {             // any geometry: one pixel per segment, cursor advanced with uniform carries
      if (++x2 == p.E[2]) { x2 = 0; if (++x1 == p.E[1]) { x1 = 0; if (++x0 == p.E[0]) { x0 = 0; ++xn; } } }
    } else if (SEG < TC_BK) {   // SEG == E2: the next segment is the next image row
      if (++x1 == p.E[1]) { x1 = 0; if (++x0 == p.E[0]) { x0 = 0; ++xn; } }
    }
  }
}

// WG = 1 is the weight-gradient GEMM D[(tap,c)][n] = sum_m Src[pix(m,tap)][c] * G[m][n] (K = output pixels):
// the same pipeline with a different A gather - thread = GEMM row (tap, c), one k-block = that row's values
// at 32 consecutive output pixels, which adjacent lanes (adjacent channels) read as coalesced lines; the pixel
// cursor advances incrementally with warp-uniform carries - and B = the gradient, pre-split into stage images
// by pack_grad_kernel (MN-major, as it lies in HBM).
// PROF = 1 compiles the (PROF ? clock64() : 0ll) role timers in (scripts/gpu_role_prof.py); the production instantiation has
// none - they cost a dozen registers in a kernel that sits at the 128-register limit.
template <int B_MN, int WG, int PROF>
__global__ void __launch_bounds__(TCP_THREADS)
igemm_tc_pixel_kernel(const GemmPlan pin, const float* __restrict__ A, const float* __restrict__ Wp,
                      const float* __restrict__ bias, float* __restrict__ D, int act, float alpha,
                      int bn, int bn_smem, int nb, int tmem_cols, int kb_per_split, long long part_stride,
                      int csize, int ny, int dbg, long long* prof, const __grid_constant__ CUtensorMap omap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // Work items.  ny == 0 (wgrad, clusters): one item per CTA, (blockIdx.x, blockIdx.y) = (M tile, n-tile/phase).
  // ny > 0: PERSISTENT - items w = (M tile, y) with y = w % ny (n-tile and phase) fastest; CTA i takes w = i, i + G, ...
  // (G = gridDim.x, coprime with ny so that every CTA sees all phases).  All pipeline state - stage indices,
  // barrier phases, the accumulator ping-pong - runs on across items, so the epilogue stores of item j (measured:
  // 15-33 % of a one-item CTA, profiles/r01_role_prof_v7_dbg.txt) overlap the main loop of item j+1 and the
  // prologue (barrier init, TMEM allocation, pipeline fill) is paid once per SM.
  const int tiles_m = ((WG ? pin.Ktot : pin.M) + TC_BM - 1) / TC_BM;
  const int n_items = ny > 0 ? tiles_m * ny : 1;
  const int w_first = ny > 0 ? (int)blockIdx.x : 0, w_step = ny > 0 ? (int)gridDim.x : 1;
  const bool do_prof = PROF && prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  long long pt[6] = {0, 0, 0, 0, 0, 0};
  const long long t_begin = (PROF ? clock64() : 0ll);
  const TcpLayout L = tcp_layout(nb, bn_smem);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_fullb = sbase + L.bar_off, bar_emptyb = bar_fullb + 8 * nb;   // empty_b: free in EVERY CTA of the cluster
  const uint32_t bar_fulla = bar_emptyb + 8 * nb, bar_emptya = bar_fulla + 8 * TCP_MAX_A;
  const uint32_t bar_accfull = bar_emptya + 8 * TCP_MAX_A, bar_accempty = bar_accfull + 16;
  const int bn_r = (bn_smem + 31) / 32 * 32;
  const uint32_t a_col0 = 2 * bn_r;       // TMEM: [acc ping][acc pong][A stage 0] .. [A stage na-1]
  const int na = min(TCP_MAX_A, (512 - 2 * bn_r) / TCP_A_COLS) & ~1;   // even: a stage keeps the parity of its k-blocks
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_off);
  int2* s_taps = reinterpret_cast<int2*>(smem + L.taps_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb_stride = pin.nphase > 1 ? pin.kb_stride : ((WG ? pin.M : pin.Ktot) + TC_BK - 1) / TC_BK;   // slot stride of the packed B buffer
  const int kb_beg = blockIdx.z * kb_per_split;             // split-K over gridDim.z (atomic epilogue)
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  // item w -> plan of its phase, tile origin, index of its B slot row, offset of its taps in s_taps, its k-blocks
  auto item = [&](int w, GemmPlan& p, int& m0, int& n0, int& y, int& tap0, int& num_kb) {
    p = pin;
    int tile;
    if (ny > 0) { y = w % ny; tile = w / ny; } else { y = blockIdx.y; tile = blockIdx.x; }
    const int ntile = cn_select_phase(p, y, ny > 0 ? ny : (int)gridDim.y);
    tap0 = (int)(p.taps - pin.taps);
    m0 = tile * TC_BM; n0 = ntile * bn;
    const int total_kb = ((WG ? p.M : p.Ktot) + TC_BK - 1) / TC_BK;
    num_kb = min(total_kb, kb_beg + kb_per_split) - kb_beg;   // host guarantees >= 1
  };

  for (int i = tid; i < pin.taps_total; i += TCP_THREADS) s_taps[i] = pin.taps[i];     // the taps of ALL phases
  if (tid == 0) {
    for (int s = 0; s < nb; ++s) { mbar_init(bar_fullb + 8 * s, 1); mbar_init(bar_emptyb + 8 * s, csize); }
    for (int s = 0; s < TCP_MAX_A; ++s) { mbar_init(bar_fulla + 8 * s, 4); mbar_init(bar_emptya + 8 * s, 1); }   // one arrival per gather warp
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accempty + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TCP_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + L.tmem_off), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();      // every CTA's barriers are initialised before a peer multicasts into them
  tc_fence_after();
  // The kernel allocates all 512 columns, so the allocation starts at column 0 / lane 0.  Using the constant
  // keeps every TMEM address in uniform registers (a base read back from shared memory costs an
  // ELECT + R2UR round trip per tcgen05.mma: ~60 clk each, measured).
  if (*tmem_ptr != 0u) { asm volatile("trap;"); }
  constexpr uint32_t tmem_base = 0u;

  if (warp < 8) {
    // ===== A gather straight into tensor memory: thread = GEMM row (TMEM lane), one k-block = the 128
    //       contiguous bytes (32 channels) of that row's source pixel.  Warps w and w+4 share a lane quarter
    //       and alternate k-blocks (A stage = k-block parity); the loads of a thread's next k-block are in
    //       flight while the current one is split into its tf32 big/small parts and written with tcgen05.st. =====
    const int q4 = warp & 3, par = warp >> 2;
    const uint32_t a_t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + a_col0;
    int s_sa = par, s_u = 0;                 // A stage and use count of this warp's next k-block to store (run on across items)
    int gk_base = 0;                         // k-blocks of this CTA's earlier items: the two warp sets alternate GLOBAL k-blocks
   for (int w = w_first; w < n_items; w += w_step) {
    GemmPlan p; int m0, n0, ysel, tap0, num_kb;
    item(w, p, m0, n0, ysel, tap0, num_kb);
    const int kb_first = (par - gk_base) & 1;
    gk_base += num_kb;
    // ---- pixel mode state.  Gather map (TCP_LPR lanes per row): load instruction i = g * TCP_HPR + h, lane (r, c) =
    // (lane / LPR, lane % LPR) reads the 16-byte chunk h * LPR + c of row r * LPR + g of the warp's 32-row group - the
    // LPR lanes of a group read adjacent chunks of one row.  After the transpose (put_a) thread (r, c) holds the eight
    // chunks of row r * LPR + c = ITS OWN lane index, splits them and writes them to its tensor-memory lane.
    const RowInfo row = decode_row(p, WG ? p.M : m0 + q4 * 32 + lane);
    auto src_off = [&](int kt) -> uint32_t {
      if (kt >= p.ntaps) return 0xffffffffu;
      const uint32_t sp = src_pixel(p, row, s_taps[tap0 + kt].x);
      return sp != 0xffffffffu ? sp * (uint32_t)p.Csrc : 0xffffffffu;
    };
    // position of this warp's next k-block to load: (tap, channel) and whether k is still inside Ktot;
    // advanced by two k-blocks per load without divisions
    int l_kt, l_c;
    { const int k = (kb_beg + kb_first) * TC_BK; l_kt = k / p.Csrc; l_c = k - l_kt * p.Csrc; }
    uint32_t l_so = 0xffffffffu;
    int l_so_kt = -1;
    auto load_a = [&](float4* v) {
      const long long t0 = (PROF ? clock64() : 0ll);
      const int c = lane & (TCP_LPR - 1), r = lane / TCP_LPR;
      const bool skip = (dbg & 1) != 0;
      if (l_c + TC_BK <= p.Csrc) {
        // common case: the 32 channels of this k-block lie inside one tap - the row owners keep their source offset
        // for the tap, the loading lanes fetch it with one shuffle per row group (none with one lane per row)
        if (l_kt != l_so_kt) { l_so = src_off(l_kt); l_so_kt = l_kt; }
#pragma unroll
        for (int g = 0; g < TCP_LPR; ++g) {
          const uint32_t so = TCP_LPR == 1 ? l_so : __shfl_sync(0xffffffffu, l_so, r * TCP_LPR + g);    // owner of row r * LPR + g
#pragma unroll
          for (int h = 0; h < TCP_HPR; ++h)
            v[g * TCP_HPR + h] = (so != 0xffffffffu && !skip) ? ldg128(A + (size_t)so + l_c + 4 * (h * TCP_LPR + c))
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        // the k-block straddles taps (Csrc not a multiple of 32) or the end of K: every chunk finds its own (tap, channel)
        // and the source pixel of its row from the owner's row decode
#pragma unroll
        for (int g = 0; g < TCP_LPR; ++g) {
          RowInfo ri = row;
          if (TCP_LPR > 1) {
            ri.base = __shfl_sync(0xffffffffu, row.base, r * TCP_LPR + g);
            ri.pk = __shfl_sync(0xffffffffu, row.pk, r * TCP_LPR + g);
          }
          int kt = l_kt, cc = l_c + 4 * c;
          while (cc >= p.Csrc) { cc -= p.Csrc; ++kt; }
#pragma unroll
          for (int h = 0; h < TCP_HPR; ++h) {
            const uint32_t sp = kt < p.ntaps ? src_pixel(p, ri, s_taps[tap0 + kt].x) : 0xffffffffu;
            v[g * TCP_HPR + h] = (sp != 0xffffffffu && !skip) ? ldg128(A + (size_t)sp * p.Csrc + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
            cc += 4 * TCP_LPR;
            while (cc >= p.Csrc) { cc -= p.Csrc; ++kt; }
          }
        }
      }
      l_c += 2 * TC_BK;
      while (l_c >= p.Csrc) { l_c -= p.Csrc; ++l_kt; }
      pt[2] += (PROF ? clock64() : 0ll) - t0;
    };
    // LPR x LPR transpose of the register index g against the lane's chunk slot c (per h): afterwards register
    // g' * TCP_HPR + h of lane (r, c) holds chunk h * LPR + g' of row r * LPR + c (= lane)
    auto transpose_rows = [&](float4* v) {
#pragma unroll
      for (int m = 1; m < TCP_LPR; m <<= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int g = 0; g < TCP_LPR; ++g) {
          if (g & m) continue;
#pragma unroll
          for (int h = 0; h < TCP_HPR; ++h) {
            float4& lo = v[g * TCP_HPR + h];
            float4& hi = v[(g | m) * TCP_HPR + h];
            const float4 send = up ? lo : hi;
            float4 recv;
            recv.x = __shfl_xor_sync(0xffffffffu, send.x, m); recv.y = __shfl_xor_sync(0xffffffffu, send.y, m);
            recv.z = __shfl_xor_sync(0xffffffffu, send.z, m); recv.w = __shfl_xor_sync(0xffffffffu, send.w, m);
            if (up) lo = recv; else hi = recv;
          }
        }
      }
    };
    // ---- wgrad mode state: this thread's (tap, channel) row and the warp-uniform pixel cursor (n, e0, e1, e2)
    const int wr = m0 + q4 * 32 + lane;
    const bool wrok = WG && wr < p.Ktot;
    int woff0 = 0, woff1 = 0, woff2 = 0, wc = 0;
    int e2 = 0, e1 = 0, e0 = 0, en = 0;
    const WgGeom wg = wg_geom(p);
    if (WG) {
      if (wrok) {
        const int kt = wr / p.Csrc; wc = wr - kt * p.Csrc;
        const int pk = s_taps[tap0 + kt].x;
        woff0 = (pk & 1023) - 8; woff1 = ((pk >> 10) & 1023) - 8; woff2 = (pk >> 20) - 8;
        if (wg.E[2] != p.E[2] || wg.E[1] != p.E[1]) { woff2 = woff1; woff1 = woff0; woff0 = 0; }    // 2-D plan: row in position 1
      }
      int m = (kb_beg + kb_first) * TC_BK;
      e2 = m % wg.E[2]; m /= wg.E[2]; e1 = m % wg.E[1]; m /= wg.E[1]; e0 = m % wg.E[0]; en = m / wg.E[0];
    }
    const float* Ac = A + wc;
    const float* Ac_ld = (dbg & 1) ? nullptr : Ac;
    auto load_w = [&](float4* vv) {
      uint32_t* v = reinterpret_cast<uint32_t*>(vv);
      const bool ok = wrok && Ac_ld != nullptr;
      const int seg = wg.E[2];
      if (seg % TC_BK == 0) wg_load<32>(wg, Ac, ok, woff0, woff1, woff2, en, e0, e1, e2, v);
#ifndef CN_WG_SEG32_ONLY      // code-size experiment (profiles/r02_m4_wgrad_code_size_ab.txt): only the 32-pixel and the general variant
      else if (seg == 16) wg_load<16>(wg, Ac, ok, woff0, woff1, woff2, en, e0, e1, e2, v);
      else if (seg == 8) wg_load<8>(wg, Ac, ok, woff0, woff1, woff2, en, e0, e1, e2, v);
      else if (seg == 4) wg_load<4>(wg, Ac, ok, woff0, woff1, woff2, en, e0, e1, e2, v);
#endif
      else wg_load<1>(wg, Ac, ok, woff0, woff1, woff2, en, e0, e1, e2, v);
      e2 += 2 * TC_BK;                                           // the sibling warp takes the next k-block
      while (e2 >= wg.E[2]) {
        e2 -= wg.E[2];
        if (++e1 == wg.E[1]) { e1 = 0; if (++e0 == wg.E[0]) { e0 = 0; ++en; } }
      }
    };
    auto load_any = [&](float4* v) { if (WG) load_w(v); else load_a(v); };
    auto store_a = [&]() -> uint32_t {
      const int sa = s_sa;
      { const long long t0 = (PROF ? clock64() : 0ll); if (s_u >= 1) mbar_wait(bar_emptya + 8 * sa, (s_u - 1) & 1); pt[0] += (PROF ? clock64() : 0ll) - t0; }
      tc_fence_after();
      s_sa += 2;
      if (s_sa >= na) { s_sa -= na; ++s_u; }
      return (uint32_t)sa;
    };
    auto put_a = [&](uint32_t sa, float4* v) {
      const long long t0 = (PROF ? clock64() : 0ll);
      const uint32_t a_t = a_t0 + sa * TCP_A_COLS;
      if (!WG && TCP_LPR > 1) transpose_rows(v);
      if (!(dbg & 8)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t bg[16], sm[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // chunk k = 4h + q of this thread's row: register (k % LPR) * HPR + k / LPR in pixel mode, register k in wgrad mode
            const int kch = 4 * h + q;
            const float4 vv = WG ? v[kch] : v[(kch % TCP_LPR) * TCP_HPR + kch / TCP_LPR];
            const float x[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // kind::tf32 reads the upper 19 bits of an operand word and ignores the 13 low mantissa bits (measured:
              // profiles/r02_probes_first_run.txt, probe 1 - truncation, exact to 1.6e-7): the big half is the RAW value,
              // the small half the exact residual x - trunc(x), itself truncated by the tensor core.  2 ALU ops per
              // element instead of 3, bit-identical products.
              bg[4 * q + e] = __float_as_uint(x[e]);
              sm[4 * q + e] = __float_as_uint(x[e] - __uint_as_float(__float_as_uint(x[e]) & 0xffffe000u));
            }
          }
          tc_st16_nowait(a_t + 16 * h, bg);
          tc_st16_nowait(a_t + 32 + 16 * h, sm);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();                      // one arrival per warp: 128 per-thread arrivals on one barrier serialise
      if (lane == 0) mbar_arrive(bar_fulla + 8 * sa);
      pt[1] += (PROF ? clock64() : 0ll) - t0;
    };
    float4 va[8], vb[8];
    int kb = kb_first;
    if (kb < num_kb) load_any(va);
    for (; kb < num_kb; kb += 4) {
      if (kb + 2 < num_kb) load_any(vb);
      put_a(store_a(), va);
      if (kb + 2 < num_kb) {
        if (kb + 4 < num_kb) load_any(va);
        put_a(store_a(), vb);
      }
    }
   }
    if (do_prof && warp == 0 && lane == 0) { prof[0] = pt[0]; prof[1] = pt[1]; prof[2] = pt[2]; prof[3] = (PROF ? clock64() : 0ll) - t_begin; }
  } else if (warp < 12) {
    // ===== promotion + epilogue: TMEM -> registers -> global (warp pw owns TMEM lanes 32pw..32pw+31) =====
    const int pw = warp - 8;
#ifdef CN_TC_COMPACT_ACT
    const float act_slope = act == CN_ACT_LRELU ? alpha : (act == CN_ACT_RELU ? 0.f : 1.f);
    const bool act_tanh = act == CN_ACT_TANH;
#endif
    const int chunk_kb = (dbg >> 8) ? (dbg >> 8) : TC_CHUNK_KB;
    int c0 = 0;                              // accumulation chunks of this CTA's earlier items
    bool tma_pending = false;                // a TMA tensor store of this CTA may still be reading the running-total buffer
   for (int w = w_first; w < n_items; w += w_step) {
    GemmPlan p; int m0, n0, ysel, tap0, num_kb;
    item(w, p, m0, n0, ysel, tap0, num_kb);
    const int m = m0 + pw * 32 + lane;
    const bool mok = m < (WG ? p.Ktot : p.M);
    const size_t rowoff = mok ? (WG ? (size_t)m : (size_t)dest_pixel(p, m)) * p.Cn : 0;
    const int nchunks = (num_kb + chunk_kb - 1) / chunk_kb;
    // how the finished tile leaves: dbg bit 6 = TMA tensor store from the running-total buffer, bit 5 = a pass with the
    // lanes along the channels, else every thread stores its own row
    const bool tma_out = (dbg & 64) != 0, coal = (dbg & 32) != 0;
    uint8_t* tot = smem + L.tot_off;
    if (tma_out && tma_pending) {          // the previous item's stores must have read the buffer before it is written again
      if (pw == 0 && lane == 0) tma_store_wait_read();
      epi_bar_sync();
      tma_pending = false;
    }
    tc_promote_smem_and_store(tmem_base, tot, pw, bn, bn_r, c0, nchunks, bar_accfull, bar_accempty, coal || tma_out,
                              [&](int cb, const uint32_t* v) {
      if (mok && !(dbg & 16)) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          int n = n0 + cb + q;
          if (cb + q < bn && n < p.Cn) {
            float4 o;
            o.x = __uint_as_float(v[q]); o.y = __uint_as_float(v[q + 1]);
            o.z = __uint_as_float(v[q + 2]); o.w = __uint_as_float(v[q + 3]);
            if (part_stride) {       // split-K: raw partial sums into this split's slab (splitk_reduce_kernel finishes)
              *reinterpret_cast<float4*>(D + (size_t)blockIdx.z * (size_t)part_stride + rowoff + n) = o;
              continue;
            }
            if (bias != nullptr) { o.x += bias[n]; o.y += bias[n + 1]; o.z += bias[n + 2]; o.w += bias[n + 3]; }
#ifdef CN_TC_COMPACT_ACT
            o.x = tc_act(o.x, act_slope, act_tanh); o.y = tc_act(o.y, act_slope, act_tanh);
            o.z = tc_act(o.z, act_slope, act_tanh); o.w = tc_act(o.w, act_slope, act_tanh);
#else
            o.x = cn_apply_act(o.x, act, alpha); o.y = cn_apply_act(o.y, act, alpha);
            o.z = cn_apply_act(o.z, act, alpha); o.w = cn_apply_act(o.w, act, alpha);
#endif
            *reinterpret_cast<float4*>(D + rowoff + n) = o;
          }
        }
      }
    },
                              [&](int cb, uint32_t* v) {
      // finish the sums in registers (bias, activation) before they go back to the buffer
      if (part_stride) return;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int n = n0 + cb + q;
        float o = __uint_as_float(v[q]);
        if (bias != nullptr && cb + q < bn && n < p.Cn) o += bias[n];
#ifdef CN_TC_COMPACT_ACT
        v[q] = __float_as_uint(tc_act(o, act_slope, act_tanh));
#else
        v[q] = __float_as_uint(cn_apply_act(o, act, alpha));
#endif
      }
    });
    if (tma_out) {
      // generic-proxy writes -> visible to the TMA unit, all four warps done, then ONE thread stores the chunks: no
      // store instruction of this warp touches global memory, partial tiles are clipped by the tensor map
      fence_proxy_async();
      epi_bar_sync();
      if (pw == 0 && lane == 0 && !(dbg & 16)) {
        if (!WG && p.nphase > 1) {
          const int x0 = m0 % p.E[1], r0 = m0 / p.E[1];
          for (int cb = 0; cb < bn; cb += 32)
            tma_store_5d(&omap, smem_u32(tot + (cb >> 5) * TC_TOT_CHUNK), n0 + cb, p.ooff[1], x0, p.ooff[0], r0);
        } else {
          for (int cb = 0; cb < bn; cb += 32) tma_store_2d(&omap, smem_u32(tot + (cb >> 5) * TC_TOT_CHUNK), n0 + cb, m0);
        }
        tma_store_commit();
      }
      tma_pending = true;
    } else if (coal) {
      // The thread-per-row stores put 32 different lines under one instruction.  Here a warp walks the 32 rows it
      // promoted itself (no cross-warp dependency) with its lanes along the channels: one 128-byte line per instruction.
      __syncwarp();
      float* Dz = D + (part_stride ? (size_t)blockIdx.z * (size_t)part_stride : (size_t)0);
      const uint32_t myoff = mok ? (uint32_t)rowoff : 0xffffffffu;      // element offsets fit 32 bits (checked when the plan is built)
      if (!(dbg & 16)) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const uint32_t ro = __shfl_sync(0xffffffffu, myoff, r);
          if (ro == 0xffffffffu) continue;
          const int rr = pw * 32 + r;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = lane + 32 * i, n = n0 + c;
            if (c < bn && n < p.Cn)
              Dz[ro + n] = *reinterpret_cast<const float*>(tot + i * TC_TOT_CHUNK + tot_unit_off(rr, lane >> 2) + (lane & 3) * 4);
          }
        }
      }
      __syncwarp();                       // the rows may be overwritten by the next item's first chunk
    }
    c0 += nchunks;
   }
    if (tma_pending && pw == 0 && lane == 0) tma_store_wait_all();      // the last tile's stores complete before the CTA leaves
  } else if (warp == TCP_B_WARP) {
    // ===== B stage fetch: the pre-packed smem image of (n-tile, k-block) is one contiguous block; this CTA
    //       fetches slice `rank` of it and multicasts it to the same offset in all CTAs of the cluster =====
    if (lane == 0) {
      const uint32_t bytes = L.stage_bytes;
      const uint32_t slice = bytes / (uint32_t)csize;
      const uint32_t rank = csize > 1 ? cluster_ctarank() : 0u;
      int s = 0, ph = 1, git = 0;            // stage, phase and fetch count run on across items
     for (int w = w_first; w < n_items; w += w_step) {
      GemmPlan p; int m0, n0, ysel, tap0, num_kb;
      item(w, p, m0, n0, ysel, tap0, num_kb);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(Wp) + ((size_t)ysel * kb_stride + kb_beg) * bytes + rank * slice;
      for (int it = 0; it < num_kb; ++it, ++git) {
        { const long long t0 = (PROF ? clock64() : 0ll); if (git >= nb) mbar_wait(bar_emptyb + 8 * s, ph); pt[0] += (PROF ? clock64() : 0ll) - t0; }
        mbar_arrive_expect_tx(bar_fullb + 8 * s, bytes);
        const uint32_t dst = sbase + s * L.stage_bytes + rank * slice;
        if (csize > 1) bulk_g2s_mc(dst, src + (size_t)it * bytes, slice, bar_fullb + 8 * s, cmask);
        else bulk_g2s(dst, src + (size_t)it * bytes, bytes, bar_fullb + 8 * s);
        if (++s == nb) { s = 0; ph ^= 1; }
      }
     }
      if (do_prof) { prof[10] = pt[0]; prof[11] = (PROF ? clock64() : 0ll) - t_begin; }
    }
  } else {
    // ===== MMA issue (one lane): per k-block 4 k-steps x 3 split products, A from TMEM, B from smem.
    //       Stage indices / phases advance incrementally and the smem descriptors are a constant plus an
    //       address field: the single issuing thread must stay far below one k-block of MMA time. =====
    const uint32_t idesc = umma_idesc_tf32(bn_smem, 0, B_MN);
    const uint32_t sbo_b = B_MN ? (uint32_t)(bn_smem >> 5) * 512u : 1024u;
    const uint32_t lbo_b = B_MN ? 512u : 16u, lb = B_MN ? 1u : 2u;
    const uint32_t desc_hi = ((sbo_b >> 4) & 0x3fffu) | (1u << 14) | (lb << 29);
    const uint32_t desc_lo0 = ((lbo_b >> 4) & 0x3fffu) << 16;
    const uint32_t stage16 = L.stage_bytes >> 4, plane16 = L.b_bytes >> 4, kk16 = (B_MN ? 2 * sbo_b : 32u) >> 4;
    int sa = 0, pa = 0, sb = 0, pb = 0, b = 0, inchunk = 0, c = 0;
    const int chunk_kb = (dbg >> 8) ? (dbg >> 8) : TC_CHUNK_KB;
    uint32_t b16 = sbase >> 4;
    int last_num_kb = 0;          // role timers: k-blocks of ALL items of this CTA (the per-k-block averages divide by it)
    // ONE elected thread runs the whole role (waits, MMAs, commits): no election / reconvergence per k-block, and the
    // two operand barriers are polled together.  Role timers (r02_role_prof_lpr2.txt) had this warp busy all the time:
    // ~890 clk per k-block inside the issue section (the tensor pipe's back-pressure) plus ~450 clk of waits, election
    // and bookkeeping during which the pipe ran dry.
   if (elect_one()) {
   for (int w = w_first; w < n_items; w += w_step) {
    GemmPlan p; int m0, n0, ysel, tap0, num_kb;
    item(w, p, m0, n0, ysel, tap0, num_kb);
    last_num_kb += num_kb;
    for (int kb = 0; kb < num_kb; ++kb) {
      const bool chunk_first = inchunk == 0;
      const bool chunk_last = inchunk == chunk_kb - 1 || kb == num_kb - 1;
      long long t0 = (PROF ? clock64() : 0ll);
      if (chunk_first && c >= 2) { mbar_wait(bar_accempty + 8 * b, ((c >> 1) - 1) & 1); }
      long long t1 = (PROF ? clock64() : 0ll); pt[0] += t1 - t0;
      mbar_wait2(bar_fulla + 8 * sa, pa, bar_fullb + 8 * sb, pb);
      t0 = (PROF ? clock64() : 0ll); pt[1] += t0 - t1;
      t1 = t0;
      tc_fence_after();
      if (!(dbg & 4)) {
        const uint32_t d_t = tmem_base + b * bn_r;
        const uint32_t a_big = tmem_base + a_col0 + sa * TCP_A_COLS, a_small = a_big + 32;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk) {
          const uint32_t lo_big = desc_lo0 | ((b16 + kk * kk16) & 0x3fffu), lo_small = desc_lo0 | ((b16 + plane16 + kk * kk16) & 0x3fffu);
          const uint64_t bb = ((uint64_t)desc_hi << 32) | lo_big, bs = ((uint64_t)desc_hi << 32) | lo_small;
          tc_mma_tf32_ts(d_t, a_small + kk * 8, bb, idesc, !(chunk_first && kk == 0));
          tc_mma_tf32_ts(d_t, a_big + kk * 8, bs, idesc, 1);
          tc_mma_tf32_ts(d_t, a_big + kk * 8, bb, idesc, 1);
        }
      }
      tc_commit(bar_emptya + 8 * sa);
      if (csize > 1) tc_commit_mc(bar_emptyb + 8 * sb, cmask); else tc_commit(bar_emptyb + 8 * sb);
      if (chunk_last) tc_commit(bar_accfull + 8 * b);
      if (++sa == na) { sa = 0; pa ^= 1; }
      b16 += stage16;
      if (++sb == nb) { sb = 0; pb ^= 1; b16 = sbase >> 4; }
      if (++inchunk == chunk_kb) { inchunk = 0; ++c; b ^= 1; }
      pt[3] += (PROF ? clock64() : 0ll) - t1;
    }
    if (inchunk != 0) { inchunk = 0; ++c; b ^= 1; }      // the item's last chunk was committed short: the next item starts a new one
   }
    if (do_prof) { prof[4] = pt[0]; prof[5] = pt[1]; prof[6] = 0; prof[7] = pt[3]; prof[8] = (PROF ? clock64() : 0ll) - t_begin; prof[9] = last_num_kb; }
   }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();      // no CTA leaves while a peer may still signal its barriers
  if (do_prof && tid == 0) prof[12] = (PROF ? clock64() : 0ll) - t_begin;
  if (warp == TCP_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Skinny layers (HBM-bound): dedicated CUDA-core kernels
// ------------------------------------------------------------------------------------------------
// Pixel mode with Cn <= 4 outputs per pixel (map_final 32->3, dgrad into 3-channel images): one thread per
// output pixel, weights in shared memory (broadcast reads), float4 walks over the source channels.
__global__ void __launch_bounds__(256)
pixel_smalln_kernel(GemmPlan p, const float* __restrict__ A, const float* __restrict__ W,
                    const float* __restrict__ bias, float* __restrict__ D, int act, float alpha) {
  extern __shared__ float s_w[];                 // [Ktot][4] (n padded to 4)
  __shared__ int2 s_taps[256];
  const int tid = threadIdx.x;
  for (int i = tid; i < p.ntaps; i += 256) s_taps[i] = p.taps[i];
  __syncthreads();
  for (int i = tid; i < p.Ktot * 4; i += 256) {
    int k = i >> 2, n = i & 3;
    int kt = k / p.Csrc, c = k - kt * p.Csrc;
    s_w[i] = (n < p.Cn) ? W[(size_t)s_taps[kt].y + (size_t)c * p.wsc + (size_t)n * p.wsn] : 0.f;
  }
  __syncthreads();
  const int m = blockIdx.x * 256 + tid;
  if (m >= p.M) return;
  const RowInfo ri = decode_row(p, m);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = (p.Csrc & 3) == 0;
  for (int kt = 0; kt < p.ntaps; ++kt) {
    const uint32_t sp = src_pixel(p, ri, s_taps[kt].x);
    if (sp == 0xffffffffu) continue;
    const float* src = A + (size_t)sp * p.Csrc;
    const float4* wv = reinterpret_cast<const float4*>(s_w) + (size_t)kt * p.Csrc;
    if (vec) {
      for (int c = 0; c < p.Csrc; c += 4) {
        const float4 x = ldg128(src + c);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = wv[c + q];
          acc[0] = fmaf(xs[q], w4.x, acc[0]); acc[1] = fmaf(xs[q], w4.y, acc[1]);
          acc[2] = fmaf(xs[q], w4.z, acc[2]); acc[3] = fmaf(xs[q], w4.w, acc[3]);
        }
      }
    } else {
      for (int c = 0; c < p.Csrc; ++c) {
        const float x = __ldg(src + c);
        const float4 w4 = wv[c];
        acc[0] = fmaf(x, w4.x, acc[0]); acc[1] = fmaf(x, w4.y, acc[1]);
        acc[2] = fmaf(x, w4.z, acc[2]); acc[3] = fmaf(x, w4.w, acc[3]);
      }
    }
  }
  float* out = D + (size_t)dest_pixel(p, m) * p.Cn;
  for (int n = 0; n < p.Cn; ++n) {
    float v = acc[n] + (bias ? bias[n] : 0.f);
    out[n] = cn_apply_act(v, act, alpha);
  }
}

// Wgrad with few outputs (Ktot*Cn <= 2304: Cin = 3 or Cout = 3 layers): each block stages P pixels
// (gathered source patch + gradient row) in shared memory, each thread owns up to 9 (k,n) outputs.
constexpr int SKW_MAXOUT = 9;
__global__ void __launch_bounds__(256)
wgrad_skinny_kernel(GemmPlan p, const float* __restrict__ A, const float* __restrict__ G,
                    float* __restrict__ D, int P, int pix_per_block) {
  extern __shared__ float sm[];                  // xs[P][Ktot], gs[P][Cn]
  __shared__ int2 s_taps[256];
  __shared__ RowInfo s_rows[64];
  float* xs = sm;
  float* gs = sm + (size_t)P * p.Ktot;
  const int tid = threadIdx.x;
  for (int i = tid; i < p.ntaps; i += 256) s_taps[i] = p.taps[i];
  const int O = p.Ktot * p.Cn;
  int ok[SKW_MAXOUT], on[SKW_MAXOUT];
  float acc[SKW_MAXOUT];
#pragma unroll
  for (int j = 0; j < SKW_MAXOUT; ++j) {
    int o = tid + 256 * j;
    ok[j] = (o < O) ? o / p.Cn : -1;
    on[j] = (o < O) ? o % p.Cn : 0;
    acc[j] = 0.f;
  }
  const int mbeg = blockIdx.x * pix_per_block, mend = min(p.M, mbeg + pix_per_block);
  for (int m0 = mbeg; m0 < mend; m0 += P) {
    __syncthreads();
    if (tid < P) s_rows[tid] = decode_row(p, (m0 + tid < mend) ? m0 + tid : p.M);
    __syncthreads();
    for (int i = tid; i < P * p.Ktot; i += 256) {
      int pp = i / p.Ktot, k = i - pp * p.Ktot;
      int kt = k / p.Csrc, c = k - kt * p.Csrc;
      uint32_t sp = src_pixel(p, s_rows[pp], s_taps[kt].x);
      xs[i] = (sp != 0xffffffffu) ? __ldg(A + (size_t)sp * p.Csrc + c) : 0.f;
    }
    for (int i = tid; i < P * p.Cn; i += 256) {
      int pp = i / p.Cn;
      gs[i] = (m0 + pp < mend) ? __ldg(G + (size_t)(m0 + pp) * p.Cn + (i - pp * p.Cn)) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SKW_MAXOUT; ++j) {
      if (ok[j] < 0) continue;
      const float* xk = xs + ok[j];
      const float* gn = gs + on[j];
      float a = acc[j];
      for (int pp = 0; pp < P; ++pp) a = fmaf(xk[(size_t)pp * p.Ktot], gn[(size_t)pp * p.Cn], a);
      acc[j] = a;
    }
  }
  float* slab = D + (size_t)blockIdx.x * O;          // per-block partial; sum_slabs_kernel adds the blocks in order
#pragma unroll
  for (int j = 0; j < SKW_MAXOUT; ++j)
    if (ok[j] >= 0) slab[tid + 256 * j] = acc[j];
}

// ------------------------------------------------------------------------------------------------
// Skinny layers, second generation (HBM-bound layers with 3 channels on one side: fromRGB 3->3, D.block0 3->48,
// VGG block1_conv1 3->64, map_final 32->3).  A warp walks a contiguous range of GEMM rows (pixels) with a
// warp-uniform cursor (no per-pixel divisions); its lanes span the WIDE channel dimension, so every global access
// is a coalesced line, and the per-tap source addresses are computed lane-parallel (lane = tap or patch element)
// and broadcast with shuffles.
//   cfew wgrad: Csrc >= 8 wide, Cn <= 4   (lanes = source channels)
//   kfew wgrad: K = ntaps*Csrc <= 32, Cn wide (lanes = output channels)
// (The same warp-per-pixel formulation was measured 2-4x SLOWER than the thread-per-pixel kernels for the forward /
// dgrad direction of these layers - too many instructions per output - so those keep pixel_smalln / igemm_ffma.)
// ------------------------------------------------------------------------------------------------
struct PixCursor { int n, e0, e1, e2; };
__device__ __forceinline__ PixCursor cursor_at(const GemmPlan& p, int m) {
  PixCursor c;
  c.e2 = m % p.E[2]; m /= p.E[2];
  c.e1 = m % p.E[1]; m /= p.E[1];
  c.e0 = m % p.E[0]; c.n = m / p.E[0];
  return c;
}
__device__ __forceinline__ void cursor_next(const GemmPlan& p, PixCursor& c) {
  if (++c.e2 == p.E[2]) { c.e2 = 0; if (++c.e1 == p.E[1]) { c.e1 = 0; if (++c.e0 == p.E[0]) { c.e0 = 0; ++c.n; } } }
}
// source pixel of (cursor, tap offsets) or 0xffffffff in the padding
__device__ __forceinline__ uint32_t cursor_src(const GemmPlan& p, const PixCursor& c, int o0, int o1, int o2) {
  const int u0 = c.e0 * p.mstride + o0, u1 = c.e1 * p.mstride + o1, u2 = c.e2 * p.mstride + o2;
  const bool ok = (unsigned)u0 < (unsigned)p.U[0] && (unsigned)u1 < (unsigned)p.U[1] && (unsigned)u2 < (unsigned)p.U[2];
  if (!ok) return 0xffffffffu;
  return (((uint32_t)c.n * p.S[0] + (uint32_t)(u0 >> p.ushift)) * p.S[1] + (uint32_t)(u1 >> p.ushift)) * p.S[2] + (uint32_t)(u2 >> p.ushift);
}
__device__ __forceinline__ uint32_t cursor_dst(const GemmPlan& p, const PixCursor& c) {
  return (((uint32_t)c.n * p.Q[0] + (uint32_t)(c.e0 * p.ostride + p.ooff[0])) * p.Q[1] + (uint32_t)(c.e1 * p.ostride + p.ooff[1])) * p.Q[2] +
         (uint32_t)(c.e2 * p.ostride + p.ooff[2]);
}
__device__ __forceinline__ void tap_offsets(int pk, int& o0, int& o1, int& o2) {
  o0 = (pk & 1023) - 8; o1 = ((pk >> 10) & 1023) - 8; o2 = (pk >> 20) - 8;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// wgrad, Cn <= 4, ntaps <= 16, one 32-channel chunk per blockIdx.y: D[(tap,c)][n] += sum_pix Src[src(pix,tap)][c] * G[pix][n]
__global__ void __launch_bounds__(256)
skinny_cfew_wgrad_kernel(GemmPlan p, const float* __restrict__ A, const float* __restrict__ G, float* __restrict__ D, int rows_per_warp) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int c = blockIdx.y * 32 + lane;
  const bool cok = c < p.Csrc;
  int o0 = 0, o1 = 0, o2 = 0;
  const bool tapok = lane < p.ntaps;
  if (tapok) tap_offsets(p.taps[lane].x, o0, o1, o2);
  const int gw = blockIdx.x * 8 + (tid >> 5);
  int m = gw * rows_per_warp;
  const int mend = min(p.M, m + rows_per_warp);
  float acc[16][4];
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[t][n] = 0.f;
  if (m < mend) {
    PixCursor cur = cursor_at(p, m);
    for (; m < mend; ++m) {
      const uint32_t sp = tapok ? cursor_src(p, cur, o0, o1, o2) : 0xffffffffu;
      const float gl = lane < p.Cn ? __ldg(G + (size_t)m * p.Cn + lane) : 0.f;
      float g[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) g[n] = __shfl_sync(0xffffffffu, gl, n);
      float x[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        x[t] = 0.f;
        if (t < p.ntaps) {
          const uint32_t s = __shfl_sync(0xffffffffu, sp, t);
          if (s != 0xffffffffu && cok) x[t] = __ldg(A + (size_t)s * p.Csrc + c);
        }
      }
#pragma unroll
      for (int t = 0; t < 16; ++t)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[t][n] = fmaf(x[t], g[n], acc[t][n]);
      cursor_next(p, cur);
    }
  }
  // block reduction over the 8 warps in warp order (deterministic), then the block's partial goes to slab blockIdx.x
  __shared__ float red[16 * 32 * 4];
  for (int i = tid; i < 16 * 32 * 4; i += 256) red[i] = 0.f;
  __syncthreads();
  for (int w = 0; w < 8; ++w) {
    if ((tid >> 5) == w) {
#pragma unroll
      for (int t = 0; t < 16; ++t)
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (t < p.ntaps && n < p.Cn) red[(t * 32 + lane) * 4 + n] += acc[t][n];
    }
    __syncthreads();
  }
  float* slab = D + (size_t)blockIdx.x * ((size_t)p.Ktot * p.Cn);
  for (int i = tid; i < p.ntaps * 32 * 4; i += 256) {
    const int t = i >> 7, l = (i >> 2) & 31, n = i & 3, cc = blockIdx.y * 32 + l;
    if (n < p.Cn && cc < p.Csrc) slab[((size_t)t * p.Csrc + cc) * p.Cn + n] = red[i];
  }
}

// wgrad, K = ntaps*Csrc <= 32, Cn <= 64: D[k][n] += sum_pix patch[pix][k] * G[pix][n]
__global__ void __launch_bounds__(256)
skinny_kfew_wgrad_kernel(GemmPlan p, const float* __restrict__ A, const float* __restrict__ G, float* __restrict__ D, int rows_per_warp) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = p.Ktot;
  const bool n0ok = lane < p.Cn, n1ok = lane + 32 < p.Cn;
  int o0 = 0, o1 = 0, o2 = 0, pc = 0;
  const bool kok = lane < K;
  if (kok) { const int t = lane / p.Csrc; pc = lane - t * p.Csrc; tap_offsets(p.taps[t].x, o0, o1, o2); }
  float a0[32], a1[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) { a0[k] = 0.f; a1[k] = 0.f; }
  const int gw = blockIdx.x * 8 + (tid >> 5);
  int m = gw * rows_per_warp;
  const int mend = min(p.M, m + rows_per_warp);
  if (m < mend) {
    PixCursor cur = cursor_at(p, m);
    auto patch = [&](const PixCursor& c) -> float {
      const uint32_t sp = kok ? cursor_src(p, c, o0, o1, o2) : 0xffffffffu;
      return sp != 0xffffffffu ? __ldg(A + (size_t)sp * p.Csrc + pc) : 0.f;
    };
    float pv_n = patch(cur);
    float g0_n = n0ok ? __ldg(G + (size_t)m * p.Cn + lane) : 0.f, g1_n = n1ok ? __ldg(G + (size_t)m * p.Cn + lane + 32) : 0.f;
    for (; m < mend; ++m) {
      const float pv = pv_n, g0 = g0_n, g1 = g1_n;
      cursor_next(p, cur);
      if (m + 1 < mend) {                              // next pixel's loads in flight while this one is accumulated
        pv_n = patch(cur);
        g0_n = n0ok ? __ldg(G + (size_t)(m + 1) * p.Cn + lane) : 0.f;
        g1_n = n1ok ? __ldg(G + (size_t)(m + 1) * p.Cn + lane + 32) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (k < K) {
          const float x = __shfl_sync(0xffffffffu, pv, k);
          a0[k] = fmaf(x, g0, a0[k]); a1[k] = fmaf(x, g1, a1[k]);
        }
      }
    }
  }
  __shared__ float red[32 * 64];
  for (int i = tid; i < 32 * 64; i += 256) red[i] = 0.f;
  __syncthreads();
  for (int w = 0; w < 8; ++w) {                      // warp order: deterministic
    if ((tid >> 5) == w) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (k < K) {
          if (n0ok) red[k * 64 + lane] += a0[k];
          if (n1ok) red[k * 64 + 32 + lane] += a1[k];
        }
      }
    }
    __syncthreads();
  }
  float* slab = D + (size_t)blockIdx.x * ((size_t)K * p.Cn);
  for (int i = tid; i < K * 64; i += 256) {
    const int k = i >> 6, n = i & 63;
    if (n < p.Cn) slab[(size_t)k * p.Cn + n] = red[i];
  }
}

// wgrad of a 1x1 / stride-1 layer with <= 4 channels on both sides (fromRGB 3->3): a flat reduction over the pixels
__global__ void __launch_bounds__(256)
wgrad_flat_kernel(const float* __restrict__ X, const float* __restrict__ G, int M, int cin, int cout, float* __restrict__ D) {
  float acc[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[c][n] = 0.f;
  for (int m = blockIdx.x * 256 + threadIdx.x; m < M; m += gridDim.x * 256) {
    float x[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < cin) x[c] = __ldg(X + (size_t)m * cin + c);
#pragma unroll
    for (int n = 0; n < 4; ++n) if (n < cout) g[n] = __ldg(G + (size_t)m * cout + n);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[c][n] = fmaf(x[c], g[n], acc[c][n]);
  }
  __shared__ float red[8][16];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float v = warp_sum(acc[c][n]);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c * 4 + n] = v;
    }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int c = threadIdx.x >> 2, n = threadIdx.x & 3;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    if (c < cin && n < cout) D[(size_t)blockIdx.x * (cin * cout) + c * cout + n] = t;     // slab of this block
  }
}

// Dense layers with a handful of rows (the AdaIN / latent MLPs, the style heads, disc_map, latent_predictor:
// M = batch <= 64): y[m][n] = act(sum_k x[m][k] * W(k, n) + b[n]).  The weight matrix is read exactly once
// (lane = output column), x is staged in shared memory and broadcast, the 8 warps of a block and the blocks of
// gridDim.y split K.  W(k, n) = W[k*wsc + n*wsn] covers forward (Keras (in, out)) and dgrad (transposed).
template <int MT>
__global__ void __launch_bounds__(256)
dense_small_kernel(const float* __restrict__ X, int M, int K, const float* __restrict__ W, int wsc, int wsn,
                   const float* __restrict__ bias, float* __restrict__ Y, int N, int act, float alpha,
                   int kchunk, long long part_stride) {
  __shared__ __align__(16) float xs[MT][132];
  __shared__ float outs[MT][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x * 32 + lane;
  const bool nok = n < N;
  const int kbeg = blockIdx.y * kchunk, kend = min(K, kbeg + kchunk);
  for (int i = tid; i < MT * 32; i += 256) (&outs[0][0])[i] = 0.f;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += 128) {
    // this thread's 16 weights of the chunk first (ordered loads: all in flight), then the x tile through cp.async:
    // ONE memory round trip per chunk (the loop used to wait for x, then for the weights group by group - these layers
    // are nothing but latency, ~150 launches per step)
    float wv[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int k = k0 + warp * 16 + e;
      wv[e] = (nok && k < kend) ? cn_ldg1_ordered(W + (size_t)k * wsc + (size_t)n * wsn) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < MT * 128; i += 256) {
      const int m = i >> 7, kk = i & 127;
      const bool in = m < M && k0 + kk < kend;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(&xs[m][kk])), "l"(in ? X + (size_t)m * K + k0 + kk : X),
                   "r"(in ? 4 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
      const int kk = warp * 16 + q;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[m][kk]);
        acc[m] = fmaf(xv.x, wv[q], acc[m]); acc[m] = fmaf(xv.y, wv[q + 1], acc[m]);
        acc[m] = fmaf(xv.z, wv[q + 2], acc[m]); acc[m] = fmaf(xv.w, wv[q + 3], acc[m]);
      }
    }
  }
  // deterministic cross-warp sum: the warps add their partials in a fixed order
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < MT; ++m) outs[m][lane] += acc[m];
    }
    __syncthreads();
  }
  for (int i = tid; i < MT * 32; i += 256) {
    const int m = i >> 5, nn = blockIdx.x * 32 + (i & 31);
    if (m >= M || nn >= N) continue;
    float v = outs[m][i & 31];
    if (part_stride) {             // split-K slab of this block row (splitk_reduce_kernel adds bias)
      Y[(size_t)blockIdx.y * (size_t)part_stride + (size_t)m * N + nn] = v;
    } else {
      if (bias != nullptr) v += bias[nn];
      Y[(size_t)m * N + nn] = cn_apply_act(v, act, alpha);
    }
  }
}

// column sums for narrow matrices (n < 32): flat coalesced walk, every thread stays on one column
__global__ void __launch_bounds__(256)
colsum_narrow_kernel(const float* __restrict__ g, size_t total, int n, int threads_used, float* __restrict__ part) {
  // part[blockIdx.x][n]: the block's column sums, its threads folded in thread order (deterministic)
  __shared__ float s_acc[256];
  const int gt = blockIdx.x * 256 + threadIdx.x;
  float acc = 0.f;
  if (gt < threads_used)
    for (size_t i = gt; i < total; i += threads_used) acc += g[i];
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < n) {
    const int first = (int)(((unsigned)threadIdx.x + (unsigned)n - (unsigned)((blockIdx.x * 256) % n)) % (unsigned)n);   // first thread on this column
    float t = 0.f;
    for (int j = first; j < 256; j += n) t += s_acc[j];
    part[(size_t)blockIdx.x * n + threadIdx.x] = t;
  }
}

// column sums of a (rows, n) matrix: bias gradient
__global__ void colsum_kernel(const float* __restrict__ g, int rows, int n, float* __restrict__ out) {
  // grid.x covers columns in groups of 32, grid.y splits rows; blockDim = (32, 8); out = [gridDim.y][n] partial sums
  __shared__ float sm[8][33];
  int col = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (col < n)
    for (int r = blockIdx.y * 8 + threadIdx.y; r < rows; r += 8 * gridDim.y) acc += g[(size_t)r * n + col];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < n) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += sm[i][threadIdx.x];
    out[(size_t)blockIdx.y * n + col] = s;          // slab of this row split
  }
}

// Dense-layer weight gradient with a handful of rows (the AdaIN / latent MLPs, style heads: M = batch <= 64):
// gW[k][n] = sum_m x[m][k] * gy[m][n].  Thread = (k, column quad); its M row pairs are independent loads, eight in flight;
// the sum runs over m in order (deterministic).  The generic 64x64-tile kernel spent ~9 us per launch on these
// (~80 launches per step: staging, split-K slabs and a reduce for a few kFLOP).
__global__ void __launch_bounds__(256)
dense_wgrad_small_kernel(const float* __restrict__ X, const float* __restrict__ G, float* __restrict__ gW,
                         float* __restrict__ gbias, int M, int K, int N) {
  const int nq = N >> 2;
  const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (size_t)K * nq) return;
  const int k = (int)(t / nq), j = (int)(t - (size_t)k * nq);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), bs = acc;     // bs: the bias gradient (column sums of gy), kept by the k = 0 threads
  int m = 0;
  for (; m + 7 < M; m += 8) {
    float xv[8]; float4 g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { xv[u] = cn_ldg1_ordered(X + (size_t)(m + u) * K + k); g[u] = cn_ldg4_ordered(G + (size_t)(m + u) * N + 4 * j); }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x = fmaf(xv[u], g[u].x, acc.x); acc.y = fmaf(xv[u], g[u].y, acc.y);
      acc.z = fmaf(xv[u], g[u].z, acc.z); acc.w = fmaf(xv[u], g[u].w, acc.w);
      bs.x += g[u].x; bs.y += g[u].y; bs.z += g[u].z; bs.w += g[u].w;
    }
  }
  for (; m < M; ++m) {
    const float xv = cn_ldg1_ordered(X + (size_t)m * K + k);
    const float4 g = cn_ldg4_ordered(G + (size_t)m * N + 4 * j);
    acc.x = fmaf(xv, g.x, acc.x); acc.y = fmaf(xv, g.y, acc.y); acc.z = fmaf(xv, g.z, acc.z); acc.w = fmaf(xv, g.w, acc.w);
    bs.x += g.x; bs.y += g.y; bs.z += g.z; bs.w += g.w;
  }
  *reinterpret_cast<float4*>(gW + (size_t)k * N + 4 * j) = acc;
  if (k == 0 && gbias != nullptr) *reinterpret_cast<float4*>(gbias + 4 * j) = bs;
}

// 16-byte form (n % 4 == 0, n <= 1024): a block owns a run of rows, thread (q, rr) owns column quad q and walks the rows
// rr, rr + R, ... with eight loads in flight (the kernel above keeps ONE 4-byte load in flight per thread and took 34 us
// whatever the matrix, profiles/r02_launches_s2a_summary.txt); the R row groups are folded in order.  out = [gridDim.x][n].
__global__ void __launch_bounds__(256)
colsum4_kernel(const float* __restrict__ g, int rows, int n, int per, float* __restrict__ out) {
  __shared__ __align__(16) float red[1024];
  const int nq = n >> 2, R = 256 / nq;
  const int tid = threadIdx.x, q = tid % nq, rr = tid / nq;
  const int rbeg = blockIdx.x * per, rend = min(rows, rbeg + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rr < R) {
    int r = rbeg + rr;
    for (; r + 7 * R < rend; r += 8 * R) {                // eight rows loaded before the first is added
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = cn_ldg4_ordered(g + (size_t)(r + u * R) * n + 4 * q);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; r < rend; r += R) {
      const float4 v = cn_ldg4_ordered(g + (size_t)r * n + 4 * q);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(red + (rr * nq + q) * 4) = acc;
  }
  __syncthreads();
  if (tid < nq) {
    float4 t = *reinterpret_cast<const float4*>(red + tid * 4);
    for (int k = 1; k < R; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(red + (k * nq + tid) * 4);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (size_t)blockIdx.x * n + 4 * tid) = t;
  }
}

// out[i] = act(bias[i % cn] + sum_z ws[z][i]) in split order: the deterministic second half of every split-K launch.
// n4 = elements / 4 (the slabs are 16-byte aligned and cn % 4 == 0 on this path).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int nsplit, size_t n4, int cn4, const float* __restrict__ bias,
                     float* __restrict__ out, int act, float alpha) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 a = cn_ldg4_ordered(ws + 4 * i);
  int z = 1;
  for (; z + 3 < nsplit; z += 4) {                    // four slabs in flight, added in split order
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = cn_ldg4_ordered(ws + 4 * ((size_t)(z + u) * n4 + i));
#pragma unroll
    for (int u = 0; u < 4; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  for (; z < nsplit; ++z) {
    const float4 v = cn_ldg4_ordered(ws + 4 * ((size_t)z * n4 + i));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  if (bias != nullptr) {
    const float4 b = reinterpret_cast<const float4*>(bias)[i % (size_t)cn4];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  if (act != CN_ACT_NONE) {
    a.x = cn_apply_act(a.x, act, alpha); a.y = cn_apply_act(a.y, act, alpha);
    a.z = cn_apply_act(a.z, act, alpha); a.w = cn_apply_act(a.w, act, alpha);
  }
  reinterpret_cast<float4*>(out)[i] = a;
}
// scalar form for outputs whose size or channel count is not a multiple of 4
__global__ void __launch_bounds__(256)
splitk_reduce1_kernel(const float* __restrict__ ws, int nsplit, size_t n, int cn, const float* __restrict__ bias,
                      float* __restrict__ out, int act, float alpha) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float a = ws[i];
  for (int z = 1; z < nsplit; ++z) a += ws[(size_t)z * n + i];
  if (bias != nullptr) a += bias[i % (size_t)cn];
  out[i] = cn_apply_act(a, act, alpha);
}
// out[i] = sum_b part[b][i] for MANY slabs of FEW outputs (per-block partials of the small weight / bias gradients):
// block (32, 8) - thread (x, y) adds the slabs y, y+8, ... of output 32*blockIdx.x + x in order, then the 8 rows are
// folded in order.  Deterministic; a single thread walking hundreds of slabs would be a chain of dependent L2 round trips.
__global__ void __launch_bounds__(1024)
sum_slabs_kernel(const float* __restrict__ part, int nslabs, int n, float* __restrict__ out) {
  // block (32, Y), Y = 8 or 32 rows (sum_slabs picks 32 above 256 slabs: the 4096 per-k-block slabs of a bias gradient
  // took 48 us with 8 rows and one load in flight); eight slabs are loaded before the first is added
  __shared__ float red[32][33];
  const int i = blockIdx.x * 32 + threadIdx.x, Y = blockDim.y;
  float a = 0.f;
  if (i < n) {
    int b = threadIdx.y;
    for (; b + 7 * Y < nslabs; b += 8 * Y) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = cn_ldg1_ordered(part + (size_t)(b + u * Y) * n + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) a += v[u];
    }
    for (; b < nslabs; b += Y) a += cn_ldg1_ordered(part + (size_t)b * n + i);
  }
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = 0.f;
    for (int k = 0; k < Y; ++k) t += red[k][threadIdx.x];
    out[i] = t;
  }
}
static inline void sum_slabs(const float* part, int nslabs, int n, float* out, cudaStream_t st) {
  sum_slabs_kernel<<<(unsigned)((n + 31) / 32), dim3(32, nslabs > 256 ? 32 : 8), 0, st>>>(part, nslabs, n, out);
}

// ------------------------------------------------------------------------------------------------
// Host dispatch
// ------------------------------------------------------------------------------------------------
static int g_cluster = 1;   // CTAs per cluster in the tcgen05 pixel kernel (cn_debug_set_cluster: 1, 2 or 4); measured per layer in
                            // profiles/r01_cluster_ab_v5.txt: sharing the weight stream by multicast no longer pays (c2/c1 = 1.00..1.06)
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_cluster(int v) { g_cluster = (v == 4 || v == 2) ? v : 1; return CN_OK; }
#endif
static long long* g_prof = nullptr;   // TEST HOOK: device buffer (16 x int64) receiving CTA 0's role timings
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_prof(void* p) { g_prof = (long long*)p; return CN_OK; }
#endif
static int g_dbg = 0;   // TEST HOOK (cn_debug_set): bit 0/1 skip A/B global loads, bit 2 skip MMA issue, bit 3 skip STS
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set(int v) { g_dbg = v; return CN_OK; }
#endif
// how a finished tile leaves the tcgen05 kernel (cn_debug_set_coal): 0 = every thread stores its row, 1 = channel-major
// pass by rule (launch_tc), 2 = channel-major pass always, 3 (default) = TMA tensor store from shared memory wherever the
// output rows are consecutive and the launch does not split K, else as 1
static int g_coal = 3;
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_coal(int v) { g_coal = v; return CN_OK; }
#endif

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}
// the layer output as a (rows, cn) fp32 matrix, box = 32 columns x 128 rows, 128-byte swizzle (the running-total layout)
static bool make_out_map(CUtensorMap* map, float* dst, long long rows, int cn) {
  EncodeTiledFn enc = tma_encode_fn();
  if (enc == nullptr || ((uintptr_t)dst & 15) != 0 || (cn & 3) != 0 || rows <= 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cn, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cn * 4};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dst, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the output of a 2-D stride-2 phase plan (q = 2 e + parity): (channels, x parity, x / 2, y parity, sample * H / 2 + y / 2),
// box {32, 1, bw, 1, 128 / bw}: one M tile of one phase
static bool make_out_map_phased(CUtensorMap* map, float* dst, const GemmPlan& g) {
  EncodeTiledFn enc = tma_encode_fn();
  if (enc == nullptr || ((uintptr_t)dst & 15) != 0 || (g.Cn & 3) != 0) return false;
  if (g.ostride != 2 || g.E[2] != 1 || g.Q[2] != 1 || g.Q[0] != 2 * g.E[0] || g.Q[1] != 2 * g.E[1]) return false;
  for (int ph = 0; ph < g.nphase; ++ph) if (g.ph_ooff[ph] & 4) return false;
  const int bw = g.E[1] < 128 ? g.E[1] : 128;
  if ((g.E[1] >= 128 && g.E[1] % 128 != 0) || (g.E[1] < 128 && 128 % g.E[1] != 0)) return false;
  const cuuint64_t C4 = (cuuint64_t)g.Cn * 4, W = (cuuint64_t)g.Q[1];
  const cuuint64_t dims[5] = {(cuuint64_t)g.Cn, 2, (cuuint64_t)g.E[1], 2, (cuuint64_t)g.n_img * g.E[0]};
  const cuuint64_t strides[4] = {C4, 2 * C4, W * C4, 2 * W * C4};
  const cuuint32_t box[5] = {32, 1, (cuuint32_t)bw, 1, (cuuint32_t)(128 / bw)};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, dst, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// TMA tensor stores for 2-D stride-2 phase plans too (environment CN_TMA_PHASED=0 switches them off for A/B runs)
static int g_tma_phased = [] { const char* e = getenv("CN_TMA_PHASED"); return e ? atoi(e) : 1; }();
static int g_persistent = 1;   // persistent CTAs in the tcgen05 pixel kernel (cn_debug_set_persistent)
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_persistent(int v) { g_persistent = v; return CN_OK; }
#endif
static int g_fold = 1;     // folded upsample+conv plans (cn_debug_set_fold)
static int g_act_split = [] { const char* e = getenv("CN_ACT_SPLIT"); return e ? atoi(e) : 1; }();     // split-K of launches with a fused activation (A/B knob)
static int g_fold_split = [] { const char* e = getenv("CN_FOLD_SPLIT"); return e ? atoi(e) : 1; }();   // split-K of the folded forward at small batch (A/B knob)
static int g_s2all = 1;    // all parity phases of a stride-2 dgrad in one launch
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_fold(int fold, int s2all) { g_fold = fold; g_s2all = s2all; return CN_OK; }
#endif
static int g_chunk_kb = 0;   // k-blocks per tensor-core accumulation chunk (0 = TC_CHUNK_KB); cn_debug_set_chunk
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_chunk(int v) { g_chunk_kb = (v >= 1 && v <= 64) ? v : 0; return CN_OK; }
#endif
static thread_local int g_last_impl = 0;     // 1 = CUDA-core, 2 = tcgen05: what the last conv call on this thread ran
extern "C" int cn_last_conv_impl(void) { return g_last_impl; }
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}


template <typename K>
static int set_smem(K kernel, int bytes) {
  // once per (kernel, size): keeps the attribute call out of captured CUDA graphs
  static std::map<const void*, int> done;
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  int& have = done[(const void*)kernel];
  if (have >= bytes) return CN_OK;
  CN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  have = bytes;
  return CN_OK;
}

static bool tc_pixel_eligible(const GemmPlan& g, bool b_mn) {
  if (g.M < 128 || g.ntaps == 0 || g.Ktot < 8) return false;
  if (g.Csrc % 4 != 0) return false;
  if (b_mn) return g.Cn % 4 == 0 && g.Cn >= 16;
  return g.Cn % 16 == 0;
}

static void pick_bn_mn(int cn, int* bn, int* bn_smem) {
  // MN-major B tiles are built from 32-column swizzle atoms (any multiple of 32 up to 128); the discriminator
  // widths 96 / 192 / 288 take 96-column tiles instead of padding to 128
  int c = cn > 64 ? 128 : (cn > 32 ? 64 : 32);
  if (cn > 64 && cn % 96 == 0 && cn % 128 != 0) c = 96;
  *bn_smem = c;
  *bn = c;
}
static void pick_bn_k(int cn, int* bn, int* bn_smem) {
  int best = 16;
  for (int c = 16; c <= 128; c += 16) if (cn % c == 0) best = c;
  *bn = best; *bn_smem = best;
}

static std::map<const void*, std::pair<float*, size_t>> g_wpack;    // per plan (keyed by its tap table): packed-weight buffer
static float* g_gpack = nullptr;       // packed-gradient workspace of the wgrad GEMM (grow-only; launches on one stream serialise)
static size_t g_gpack_bytes = 0;

// Registered parameter buffers (ParamGroup flat buffers): their packed stage images are cached per (plan, weight
// pointer) and reused until THAT buffer's epoch changes - VGG19's frozen weights are packed once per run, a
// discriminator kernel once per optimizer step instead of once per conv call.
//
// CUDA graphs.  A capture records launches without running them, and a replay runs the optimizer kernels without the
// host code that advances the epochs.  Two rules keep a captured step self-contained:
//   * inside a capture an image counts as current only if it was packed INSIDE THIS capture (cudaStreamGetCaptureInfo's
//     id), so every graph re-packs the trainable kernels it uses at their first use - except images of buffers the
//     caller declared frozen (cn_set_params_frozen), which only change through cn_params_changed();
//   * outside a capture an image recorded inside one is never trusted; the caller marks the buffers a replayed graph
//     updated with cn_params_changed() after the replay (runtime.GraphedFn does).
struct ParamRange { const char* lo; const char* hi; unsigned long long epoch; bool frozen; };
static std::vector<ParamRange> g_param_ranges;
struct PackedEntry { float* buf; size_t bytes; unsigned long long epoch; unsigned long long capture_id; };
static std::map<std::pair<const void*, const void*>, PackedEntry> g_wcache;

// Once a CUDA graph has been captured over these launches (cn_graphs_captured), buffers the graph may reference
// are never freed: growth / invalidation abandons them instead.
static bool g_no_free = false;
extern "C" int cn_graphs_captured(void) { g_no_free = true; return CN_OK; }
static void cn_free(void* p) { if (!g_no_free && p) cudaFree(p); }
static ParamRange* find_range_locked(const void* w) {
  for (auto& r : g_param_ranges) if ((const char*)w >= r.lo && (const char*)w < r.hi) return &r;
  return nullptr;
}
// drops the cached images of the kernels inside [lo, hi) only (the other networks keep theirs)
static void drop_wcache_range_locked(const char* lo, const char* hi) {
  bool synced = false;
  for (auto it = g_wcache.begin(); it != g_wcache.end();) {
    const char* w = (const char*)it->first.second;
    if (w >= lo && w < hi) {
      if (!g_no_free) { if (!synced) { cudaDeviceSynchronize(); synced = true; } cn_free(it->second.buf); }
      it = g_wcache.erase(it);
    } else ++it;
  }
}
extern "C" int cn_register_params(const void* p, size_t bytes) {
  CN_REQUIRE(p != nullptr && bytes > 0, CN_ERR_BAD_SHAPE, "cn_register_params: bad range");
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  drop_wcache_range_locked((const char*)p, (const char*)p + bytes);      // a new buffer may reuse the address of a dead one
  for (auto& r : g_param_ranges) if (r.lo == (const char*)p) { r.hi = r.lo + bytes; ++r.epoch; r.frozen = false; return CN_OK; }
  g_param_ranges.push_back({(const char*)p, (const char*)p + bytes, g_cn_weight_epoch, false});
  return CN_OK;
}
extern "C" int cn_unregister_params(const void* p) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  for (size_t i = 0; i < g_param_ranges.size(); ++i)
    if (g_param_ranges[i].lo == (const char*)p) {
      drop_wcache_range_locked(g_param_ranges[i].lo, g_param_ranges[i].hi);
      g_param_ranges.erase(g_param_ranges.begin() + i);
      break;
    }
  return CN_OK;
}
// `p` = any address inside a registered buffer (the optimizer / EMA entry points pass the buffer they write)
void cn_mark_params_changed(const void* p) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  ParamRange* r = find_range_locked(p);
  if (r) ++r->epoch;
}
extern "C" int cn_params_changed(const void* p) { cn_mark_params_changed(p); return CN_OK; }
extern "C" long long cn_params_epoch(const void* p) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  ParamRange* r = find_range_locked(p);
  return r ? (long long)(r->epoch + g_cn_weight_epoch) : -1;
}
extern "C" int cn_set_params_frozen(const void* p, int frozen) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  ParamRange* r = find_range_locked(p);
  CN_REQUIRE(r != nullptr, CN_ERR_BAD_SHAPE, "cn_set_params_frozen: not a registered buffer");
  r->frozen = frozen != 0;
  return CN_OK;
}
static bool is_registered_param(const void* w) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  return find_range_locked(w) != nullptr;
}
static bool g_wcache_on = true;
#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_wcache(int on) { g_wcache_on = on != 0; return CN_OK; }
#endif
static unsigned long long capture_id_of(cudaStream_t st) {
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  if (cudaStreamGetCaptureInfo(st, &status, &id) != cudaSuccess) { cudaGetLastError(); return 0; }
  return status == cudaStreamCaptureStatusActive ? (id ? id : ~0ull) : 0;
}
// is the cached image `e` of a kernel in range `r` current for a launch on a stream whose capture id is `cap`?
static bool entry_current(const PackedEntry& e, const ParamRange& r, unsigned long long cap) {
  if (e.epoch != r.epoch + g_cn_weight_epoch) return false;
  if (r.frozen) return true;
  return e.capture_id == cap;       // eager: packed eagerly; capturing: packed inside this very capture
}
// -> *hit: the cached image is current (skip the pack).  Otherwise the caller packs into *out on the launch stream.
static int cached_packed(const void* plan_key, const void* w, size_t bytes, cudaStream_t st, float** out, bool* hit) {
  const unsigned long long cap = capture_id_of(st);
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  ParamRange* r = find_range_locked(w);
  CN_REQUIRE(r != nullptr, CN_ERR_BAD_SHAPE, "cached_packed: unregistered weight pointer");
  auto key = std::make_pair(plan_key, w);
  auto it = g_wcache.find(key);
  if (it == g_wcache.end() || it->second.bytes < bytes) {
    if (it != g_wcache.end()) { if (!g_no_free) { CN_CHECK_CUDA(cudaDeviceSynchronize()); cn_free(it->second.buf); } g_wcache.erase(it); }
    PackedEntry e; e.bytes = bytes; e.epoch = 0; e.capture_id = 0; e.buf = nullptr;
    CN_CHECK_CUDA(cudaMalloc(&e.buf, bytes));
    it = g_wcache.insert(std::make_pair(key, e)).first;
  }
  *hit = entry_current(it->second, *r, cap);
  if (!*hit) { it->second.epoch = r->epoch + g_cn_weight_epoch; it->second.capture_id = cap; }
  *out = it->second.buf;
  return CN_OK;
}
static bool packed_is_current(const void* plan_key, const void* w, cudaStream_t st) {
  if (!g_wcache_on) return false;
  const unsigned long long cap = capture_id_of(st);
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  ParamRange* r = find_range_locked(w);
  if (r == nullptr) return false;
  auto it = g_wcache.find(std::make_pair(plan_key, w));
  return it != g_wcache.end() && entry_current(it->second, *r, cap);
}

static int packed_buffer(const void* key, size_t bytes, float** out) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  if (key == nullptr) {
    if (g_gpack_bytes < bytes) {
      if (g_gpack) { CN_CHECK_CUDA(cudaDeviceSynchronize()); cn_free(g_gpack); }
      CN_CHECK_CUDA(cudaMalloc(&g_gpack, bytes));
      g_gpack_bytes = bytes;
    }
    *out = g_gpack;
    return CN_OK;
  }
  auto it = g_wpack.find(key);
  if (it == g_wpack.end() || it->second.second < bytes) {
    float* wp = nullptr;
    if (it != g_wpack.end()) { CN_CHECK_CUDA(cudaDeviceSynchronize()); cn_free(it->second.first); }
    CN_CHECK_CUDA(cudaMalloc(&wp, bytes));
    g_wpack[key] = std::make_pair(wp, bytes);
    *out = wp;
  } else {
    *out = it->second.first;
  }
  return CN_OK;
}

// Scratch slots of the deterministic reductions (grow-only like the pack buffers; launches on one stream serialise):
// split-K slabs of the GEMM kernels, per-block partials of the small weight-gradient kernels, bias-gradient partials.
static char k_ws_splitk, k_ws_small, k_ws_bias;

int cn_scratch(const void* key, size_t bytes, float** out) {
  CN_REQUIRE(key != nullptr, CN_ERR_BAD_SHAPE, "cn_scratch: null key");
  return packed_buffer(key, bytes, out);
}
int cn_sum_slabs(const float* part, int nslabs, int n, float* out, cudaStream_t st) {
  sum_slabs(part, nslabs, n, out, st);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

static int launch_splitk_reduce(const float* ws, int nsplit, size_t n, int cn, const float* bias, float* out, cudaStream_t st,
                                int act = CN_ACT_NONE, float alpha = 0.f) {
  if ((n & 3) == 0 && (cn & 3) == 0 && ((uintptr_t)out & 15) == 0 && (bias == nullptr || ((uintptr_t)bias & 15) == 0)) {
    const size_t n4 = n >> 2;
    splitk_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(ws, nsplit, n4, cn >> 2, bias, out, act, alpha);
  } else {
    splitk_reduce1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, nsplit, n, cn, bias, out, act, alpha);
  }
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// Split-K factor for `tiles` output tiles of `total_kb` k-blocks: minimise waves x (k-blocks per CTA + fixed
// per-CTA cost), i.e. fill whole waves of SMs without shredding K.  OVH ~ prologue + epilogue in k-block units.
static int pick_split(int tiles, int total_kb, int maxsplit) {
  const int sms = num_sms(), OVH = 6;
  int best = 1; long long best_cost = -1;
  int cap = (4 * sms + tiles - 1) / tiles;
  if (cap > maxsplit) cap = maxsplit;
  for (int s = 1; s <= cap; ++s) {
    const long long waves = ((long long)tiles * s + sms - 1) / sms;
    const long long cost = waves * ((total_kb + s - 1) / s + OVH + (s > 1 ? 1 : 0));
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

// B stages fill the shared memory (A lives in tensor memory); >113 KB also keeps it at one CTA per SM, which
// the 512-column TMEM allocation needs.
static void tc_smem_config(int bn_smem, int* nb, int* smem) {
  int n = 6;
  TcpLayout L = tcp_layout(n, bn_smem);
  while (n > 2 && L.total + 1024 > 200 * 1024) { --n; L = tcp_layout(n, bn_smem); }
  *nb = n;
  *smem = (int)L.total + 1024 < 120 * 1024 ? 120 * 1024 : (int)L.total + 1024;
}

template <int B_MN, int WG>
static int launch_tc(const GemmPlan& g, const float* src, const float* packed, const float* bias, float* dst, int act,
                     float alpha, int bn, int bn_smem, dim3 grid, int per, int split, cudaStream_t st, size_t part_stride = 0) {
  int nb, smem;
  tc_smem_config(bn_smem, &nb, &smem);
  // Final pass of the epilogue with the lanes along the channels (whole 128-byte lines per store instruction) where it
  // was measured to pay: forward launches with full 128-column tiles (profiles/r02_epilogue_ab.txt: +6..17 % there, a loss
  // on narrow tiles and strided outputs, where one row is only 1-3 store instructions).  g_coal: 0 never, 1 rule, 2 always.
  const bool coal = g_coal == 2 || ((g_coal == 1 || g_coal == 3) && B_MN && !WG && bn == 128 && g.nphase <= 1);
  // TMA tensor store of the finished tile: the GEMM rows must be consecutive rows of the output matrix (wgrad: always;
  // pixel mode: plain plans, i.e. no phases / strided output) and the launch must write final values (no split-K slabs)
  CUtensorMap omap;
  memset(&omap, 0, sizeof(omap));
  bool tma_out = false;
  if (g_coal == 3 && split == 1) {
    const bool rows_consecutive = WG || (g.nphase <= 1 && g.ostride == 1 && g.ooff[0] == 0 && g.ooff[1] == 0 && g.ooff[2] == 0 &&
                                         g.Q[0] == g.E[0] && g.Q[1] == g.E[1] && g.Q[2] == g.E[2]);
    // a 32-column store box must stay inside this tile's columns (or fall off the tensor, where the unit clips it)
    const bool boxes_ok = (bn % 32 == 0) || bn >= g.Cn;
    if (rows_consecutive && boxes_ok) tma_out = make_out_map(&omap, dst, WG ? (long long)g.Ktot : (long long)g.M, g.Cn);
    else if (!WG && g.nphase > 1 && boxes_ok && g_tma_phased) tma_out = make_out_map_phased(&omap, dst, g);
  }
  // thread-block cluster along M: the CTAs of a cluster share the B stream by multicast
  int csize = g_cluster;
  while (csize > 1 && (int)grid.x < csize) csize >>= 1;
  grid.x = (grid.x + csize - 1) / csize * csize;
  grid.z = split;
  int ny = 0;
  if (!WG && csize == 1 && g_persistent) {
    // persistent CTAs over the (M tile, n-tile/phase) items: one CTA per SM, G coprime with ny (see the kernel)
    ny = (int)grid.y;
    const int n_items = (int)grid.x * ny;
    int G = n_items;
    if (split == 1 && n_items > num_sms()) {
      G = num_sms();
      auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
      while (G > 1 && gcd(G, ny) != 1) --G;
    }
    grid.x = G; grid.y = 1;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(TCP_THREADS, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
#ifdef CN_TEST_HOOKS
  auto kern = g_prof ? igemm_tc_pixel_kernel<B_MN, WG, 1> : igemm_tc_pixel_kernel<B_MN, WG, 0>;   // role timers: hooks build only
#else
  auto kern = igemm_tc_pixel_kernel<B_MN, WG, 0>;
#endif
  if (set_smem(kern, smem)) return CN_ERR_CUDA;
  CN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, g, src, packed, bias, dst, act, alpha, bn, bn_smem,
                                   nb, 512, per, (long long)(split > 1 ? part_stride : 0), csize, ny,
                                   ((g_dbg & 0x9f) | (tma_out ? 64 : (coal ? 32 : 0))) | (g_chunk_kb << 8), g_prof, omap));
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// zero_mode: 0 = this launch covers all of dst and may run split-K (slabs + ordered reduction), 2 = no split-K allowed,
// 3 = phased plan covering all of dst in this one launch: split-K allowed when its phases have equal tap counts
static int launch_pixel(const GemmPlan& g, bool b_mn, const float* src, const float* w, const float* bias,
                        float* dst, int act, float alpha, int impl, cudaStream_t st, int zero_mode = 0,
                        const float* w_ident = nullptr) {
  if (g.M == 0) return CN_OK;
  bool tc = tc_pixel_eligible(g, b_mn);
  CN_REQUIRE(!(impl == CN_IMPL_TC && !tc), CN_ERR_UNSUPPORTED, "shape not eligible for the tcgen05 kernel");
  if (impl == CN_IMPL_FFMA) tc = false;
  g_last_impl = tc ? 2 : 1;
  if (tc) {
    int bn, bn_smem;
    if (b_mn) pick_bn_mn(g.Cn, &bn, &bn_smem); else pick_bn_k(g.Cn, &bn, &bn_smem);
    const int nph = g.nphase > 1 ? g.nphase : 1;
    dim3 grid((g.M + TC_BM - 1) / TC_BM, ((g.Cn + bn - 1) / bn) * nph, 1);
    const int total_kb = (g.Ktot + TC_BK - 1) / TC_BK;      // phased plans: Ktot is the longest phase (= kb_stride)
    int per = total_kb, split = 1;
    const size_t out_elems = (size_t)g.n_img * g.Q[0] * g.Q[1] * g.Q[2] * g.Cn;
    float* ws = nullptr;
    // zero_mode 3: a phased plan whose ONE launch writes every pixel of dst (the sub-pixel phases of a folded upsample + conv)
    // and whose phases all have the same number of taps - every split of every phase then has k-blocks to work on and every
    // slab is written completely.  At batch 1 (generate_images, the demo loop) the folded 3-D convs of the generator are
    // 16 / 32 tiles of 128 / 64 k-blocks on 148 SMs: 265 + 164 us of a 0.83 ms call (profiles/r02_m7_generate_timeline_b1.txt).
    bool phased_split = false;
    if (zero_mode == 3 && nph > 1) {
      phased_split = true;
      for (int ph = 1; ph < nph; ++ph) phased_split = phased_split && g.ph_ntaps[ph] == g.ph_ntaps[0];
    }
    // A fused activation does not rule split-K out: the slabs hold raw sums and splitk_reduce_kernel finishes them (bias,
    // activation).  g_act_split: 0 = only activation-free launches split (rounds 1-2), 1 = launches with an activation too
    // when their tiles do not fill one wave of SMs (the small-batch path), 2 = under the same rule as activation-free ones.
    const int tiles = (int)(grid.x * grid.y);
    const bool act_ok = act == CN_ACT_NONE || g_act_split == 2 || (g_act_split == 1 && tiles < num_sms());
    if (((nph == 1 && zero_mode == 0 && g.ostride == 1) || phased_split) && act_ok && tiles < 4 * num_sms() && total_kb >= 32) {
      // split-K without atomics: every split writes its own slab, splitk_reduce_kernel adds them in order (+ bias)
      split = pick_split((int)(grid.x * grid.y), total_kb, total_kb / 16);
      per = (total_kb + split - 1) / split;
      split = (total_kb + per - 1) / per;
      if (split > 1) { int rcw = packed_buffer(&k_ws_splitk, (size_t)split * out_elems * sizeof(float), &ws); if (rcw) return rcw; }
    }
    // pre-pack the weights into per-(n-tile, k-block) stage images (buffer cached per plan)
    float* wp = nullptr;
    bool hit = false;
    const size_t pbytes = (size_t)grid.y * total_kb * 2 * bn_smem * TC_BK * 4;
    const float* ident = w_ident ? w_ident : w;       // folded plans: identity = the original Keras kernel
    int rc;
    if (g_wcache_on && is_registered_param(ident)) rc = cached_packed((const void*)g.taps, ident, pbytes, st, &wp, &hit);
    else rc = packed_buffer((const void*)g.taps, pbytes, &wp);
    if (rc) return rc;
    if (!hit) {
      dim3 pgrid(total_kb, grid.y);
      if (b_mn) pack_weights_kernel<1><<<pgrid, 256, 0, st>>>(g, w, wp, bn, bn_smem, total_kb);
      else pack_weights_kernel<0><<<pgrid, 256, 0, st>>>(g, w, wp, bn, bn_smem, total_kb);
      CN_CHECK_LAUNCH();
    }
    float* tdst = split > 1 ? ws : dst;
    if (b_mn) rc = launch_tc<1, 0>(g, src, wp, bias, tdst, act, alpha, bn, bn_smem, grid, per, split, st, out_elems);
    else rc = launch_tc<0, 0>(g, src, wp, bias, tdst, act, alpha, bn, bn_smem, grid, per, split, st, out_elems);
    if (rc) return rc;
    if (split > 1) return launch_splitk_reduce(ws, split, out_elems, g.Cn, bias, dst, st, act, alpha);
    return CN_OK;
  }
  // CUDA-core path
  const bool flat = g.ntaps == 1 && g.E[0] * g.E[1] * g.E[2] == 1 && g.S[0] * g.S[1] * g.S[2] == 1 && g.Q[0] * g.Q[1] * g.Q[2] == 1;
  if (flat && g.M <= 64) {
    // Dense layer on a few rows
    int nt = (g.Cn + 31) / 32;
    int split = 1;
    if (act == CN_ACT_NONE && zero_mode == 0 && g.Ktot >= 1024) {
      split = (num_sms() + nt - 1) / nt;
      int maxsplit = g.Ktot / 256;
      if (split > maxsplit) split = maxsplit;
      if (split < 1) split = 1;
    }
    int kchunk = ((g.Ktot + split - 1) / split + 127) / 128 * 128;
    split = (g.Ktot + kchunk - 1) / kchunk;
    const size_t out_elems = (size_t)g.M * g.Cn;
    float* ws = nullptr;
    if (split > 1) { int rcw = packed_buffer(&k_ws_splitk, (size_t)split * out_elems * sizeof(float), &ws); if (rcw) return rcw; }
    float* tdst = split > 1 ? ws : dst;
    const long long ps = split > 1 ? (long long)out_elems : 0;
    dim3 grid(nt, split);
    if (g.M <= 16) dense_small_kernel<16><<<grid, 256, 0, st>>>(src, g.M, g.Ktot, w, g.wsc, g.wsn, bias, tdst, g.Cn, act, alpha, kchunk, ps);
    else if (g.M <= 32) dense_small_kernel<32><<<grid, 256, 0, st>>>(src, g.M, g.Ktot, w, g.wsc, g.wsn, bias, tdst, g.Cn, act, alpha, kchunk, ps);
    else dense_small_kernel<64><<<grid, 256, 0, st>>>(src, g.M, g.Ktot, w, g.wsc, g.wsn, bias, tdst, g.Cn, act, alpha, kchunk, ps);
    if (split > 1) { CN_CHECK_LAUNCH(); return launch_splitk_reduce(ws, split, out_elems, g.Cn, bias, dst, st); }
  } else if (g.Cn <= 4 && g.M >= 4096 && g.Ktot * 16 <= 96 * 1024) {
    dim3 grid((g.M + 255) / 256, 1, 1);
    int smem = g.Ktot * 16;
    if (smem > 48 * 1024 && set_smem(pixel_smalln_kernel, smem)) return CN_ERR_CUDA;
    pixel_smalln_kernel<<<grid, 256, smem, st>>>(g, src, w, bias, dst, act, alpha);
  } else if (g.Cn <= 4 && g.M >= 4096) {
    dim3 grid((g.M + 255) / 256, 1, 1);
    igemm_ffma_kernel<MODE_PIXEL, 256, 4, 1, 4><<<grid, 256, 0, st>>>(g, src, w, bias, dst, act, alpha, g.Ktot > 0 ? g.Ktot : 1, 0ll);
  } else {
    int mt = (g.M + 63) / 64, nt = (g.Cn + 63) / 64;
    int split = 1;
    if (act == CN_ACT_NONE && zero_mode == 0 && g.ostride == 1 && mt * nt < num_sms() && g.Ktot >= 2048) {
      split = (2 * num_sms() + mt * nt - 1) / (mt * nt);
      int maxsplit = g.Ktot / 256; if (maxsplit < 1) maxsplit = 1;
      if (split > maxsplit) split = maxsplit;
    }
    int kchunk = g.Ktot > 0 ? ((g.Ktot + split - 1) / split + 15) / 16 * 16 : 16;
    split = g.Ktot > 0 ? (g.Ktot + kchunk - 1) / kchunk : 1;
    const size_t out_elems = (size_t)g.n_img * g.Q[0] * g.Q[1] * g.Q[2] * g.Cn;
    float* ws = nullptr;
    if (split > 1) { int rcw = packed_buffer(&k_ws_splitk, (size_t)split * out_elems * sizeof(float), &ws); if (rcw) return rcw; }
    dim3 grid(mt, nt, split);
    igemm_ffma_kernel<MODE_PIXEL, 64, 64, 4, 4><<<grid, 256, 0, st>>>(g, src, w, bias, split > 1 ? ws : dst, act, alpha, kchunk,
                                                                      split > 1 ? (long long)out_elems : 0ll);
    if (split > 1) { CN_CHECK_LAUNCH(); return launch_splitk_reduce(ws, split, out_elems, g.Cn, bias, dst, st); }
  }
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// skinny.cu: HBM-bound layers with three channels on one side (return 1 = launched, 0 = not such a layer)
int cn_skinny_fwd(const cn_conv_desc* d, const float* x, const float* w, const float* bias, int act, float alpha,
                  float* y, cudaStream_t st);
int cn_skinny_dgrad(const cn_conv_desc* d, const float* gy, const float* w, float* gx, cudaStream_t st);
size_t cn_skinny_wgrad_scratch(const cn_conv_desc* d);
int cn_skinny_wgrad(const cn_conv_desc* d, const float* x, const float* gy, float* gw, float* gbias, float* scratch, cudaStream_t st);

static size_t conv_numel_x(const cn_conv_desc* d) {
  return (size_t)d->batch * d->in_dims[0] * d->in_dims[1] * d->in_dims[2] * d->cin;
}

extern "C" int cn_conv_out_dims(const cn_conv_desc* d, int out_dims[3]) {
  if (validate_desc(d)) return CN_ERR_BAD_SHAPE;
  int U[3], O[3], pb[3];
  same_geometry(d, U, O, pb);
  for (int i = 0; i < 3; ++i) out_dims[i] = O[i];
  return CN_OK;
}

extern "C" int cn_conv_fwd(const cn_conv_desc* d, const float* x, const float* w, const float* bias,
                           int act, float alpha, float* y, int impl, void* stream) {
  int rc = validate_desc(d); if (rc) return rc;
  CN_REQUIRE(x && w && y, CN_ERR_BAD_SHAPE, "null tensor pointer");
  CN_REQUIRE(act >= CN_ACT_NONE && act <= CN_ACT_TANH, CN_ERR_UNSUPPORTED,
             "activation code %d is not one of the conv epilogue's four (none, lrelu, relu, tanh); see cn_act_ext", act);
  if (impl != CN_IMPL_TC) {
    rc = cn_skinny_fwd(d, x, w, bias, act, alpha, y, (cudaStream_t)stream);
    if (rc < 0) return rc;
    if (rc == 1) { g_last_impl = 1; return CN_OK; }
  }
  GemmPlan g;
  if (impl != CN_IMPL_FFMA && g_fold && fold_ok(d)) {
    // nearest x2 upsample + conv on the low-resolution tensor: sub-pixel phases with pre-summed taps
    FoldInfo* fi = nullptr;
    rc = get_special_plan(d, KIND_FWD_FOLD, &g, &fi);
    if (rc == CN_OK && tc_pixel_eligible(g, true)) {
      cudaStream_t st = (cudaStream_t)stream;
      float* wf = nullptr;
      const int cc = d->cin * d->cout;
      if (!packed_is_current((const void*)g.taps, w, st)) {
        rc = packed_buffer((const char*)g.taps + 1, (size_t)fi->nfold * cc * sizeof(float), &wf); if (rc) return rc;
        fold_weights_kernel<<<dim3((cc / 4 + 255) / 256, fi->nfold), 256, 0, st>>>(w, fi->d_spec, d->ksize[1], d->ksize[2], cc / 4, wf);
        CN_CHECK_LAUNCH();
      }
      return launch_pixel(g, true, x, wf, bias, y, act, alpha, CN_IMPL_TC, st, g_fold_split ? 3 : 2, w);
    }
  }
  rc = get_plan(d, KIND_FWD, 0, &g); if (rc) return rc;
  return launch_pixel(g, true, x, w, bias, y, act, alpha, impl, (cudaStream_t)stream);
}

extern "C" int cn_conv_dgrad(const cn_conv_desc* d, const float* gy, const float* w, float* gx,
                             int impl, void* stream) {
  int rc = validate_desc(d); if (rc) return rc;
  CN_REQUIRE(gy && w && gx, CN_ERR_BAD_SHAPE, "null tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl != CN_IMPL_TC) {
    rc = cn_skinny_dgrad(d, gy, w, gx, st);
    if (rc < 0) return rc;
    if (rc == 1) { g_last_impl = 1; return CN_OK; }
  }
  if (impl != CN_IMPL_FFMA && g_fold && fold_ok(d)) {
    GemmPlan g; FoldInfo* fi = nullptr;
    rc = get_special_plan(d, KIND_DGRAD_FOLD, &g, &fi);
    if (rc == CN_OK && tc_pixel_eligible(g, false)) {
      float* wf = nullptr;
      const int cc = d->cin * d->cout;
      if (!packed_is_current((const void*)g.taps, w, st)) {
        rc = packed_buffer((const char*)g.taps + 1, (size_t)fi->nfold * cc * sizeof(float), &wf); if (rc) return rc;
        fold_weights_kernel<<<dim3((cc / 4 + 255) / 256, fi->nfold), 256, 0, st>>>(w, fi->d_spec, d->ksize[1], d->ksize[2], cc / 4, wf);
        CN_CHECK_LAUNCH();
      }
      return launch_pixel(g, false, gy, wf, nullptr, gx, CN_ACT_NONE, 0.f, CN_IMPL_TC, st, 0, w);
    }
  }
  if (impl != CN_IMPL_FFMA && g_s2all && d->stride == 2 && d->nd >= 2) {
    bool even = true;
    for (int i = 0; i < d->nd; ++i) even = even && (d->in_dims[i] % 2 == 0) && d->ksize[i] >= 2;
    GemmPlan g;
    if (even && get_special_plan(d, KIND_DGRAD_S2ALL, &g, nullptr) == CN_OK && tc_pixel_eligible(g, false))
      return launch_pixel(g, false, gy, w, nullptr, gx, CN_ACT_NONE, 0.f, CN_IMPL_TC, st, 2);   // one launch, all parity phases
  }
  const int nphase = (d->stride == 2) ? (1 << d->nd) : 1;
  const int zero_mode = nphase > 1 ? 2 : 0;       // parity phases write disjoint pixels of one tensor: no split-K slabs for them
  for (int ph = 0; ph < nphase; ++ph) {
    GemmPlan g;
    rc = get_plan(d, KIND_DGRAD, ph, &g); if (rc) return rc;
    if (g.M == 0) continue;
    int use_impl = impl;
    if (g.ntaps == 0) use_impl = CN_IMPL_FFMA;       // phase without taps: writes zeros
    rc = launch_pixel(g, false, gy, w, nullptr, gx, CN_ACT_NONE, 0.f, use_impl, st, zero_mode);
    if (rc) return rc;
  }
  return CN_OK;
}

static bool wgrad_tc_eligible(const GemmPlan& g) {
  return g.M >= 256 && g.Csrc % 4 == 0 && g.Cn % 4 == 0 && g.Cn >= 16 && g.Ktot >= 64;
}

// Weight-gradient GEMM on the tensor cores: out[(tap,c)][n] = sum_m src[pix(m,tap)][c] * G[m][n]
// gbias != nullptr (and G is the layer's output gradient): the bias gradient = column sums of G comes out of the pack pass
static int launch_wgrad_tc(const GemmPlan& g, const float* src, const float* G, float* out, cudaStream_t st, float* gbias = nullptr) {
  const size_t wn = (size_t)g.Ktot * g.Cn;
  int bn, bn_smem; pick_bn_mn(g.Cn, &bn, &bn_smem);
  int mt = (g.Ktot + TC_BM - 1) / TC_BM, nt = (g.Cn + bn - 1) / bn;
  int total_kb = (g.M + TC_BK - 1) / TC_BK;
  // split-K over the pixels so that the CTAs fill whole waves of SMs (each CTA keeps >= 8 k-blocks)
  int mtc = (mt + g_cluster - 1) / g_cluster * g_cluster;
  int maxsplit = total_kb / 8; if (maxsplit < 1) maxsplit = 1;
  int split = pick_split(mtc * nt, total_kb, maxsplit);
  int per = (total_kb + split - 1) / split;
  split = (total_kb + per - 1) / per;
  float* ws = nullptr;          // split-K slabs, summed in split order by splitk_reduce_kernel (no atomics)
  int rc;
  if (split > 1) { rc = packed_buffer(&k_ws_splitk, (size_t)split * wn * sizeof(float), &ws); if (rc) return rc; }
  // the gradient pre-split into per-(n-tile, k-block) stage images: every (tap,c)-tile CTA streams it
  float* gp = nullptr;
  rc = packed_buffer(nullptr, (size_t)nt * total_kb * 2 * bn_smem * TC_BK * 4, &gp);
  if (rc) return rc;
  float* colpart = nullptr;
  if (gbias != nullptr) { rc = packed_buffer(&k_ws_bias, (size_t)total_kb * g.Cn * sizeof(float), &colpart); if (rc) return rc; }
  pack_grad_kernel<<<dim3(total_kb, nt), 256, 0, st>>>(G, g.M, g.Cn, gp, bn, bn_smem, total_kb, colpart);
  CN_CHECK_LAUNCH();
  if (gbias != nullptr) {
    sum_slabs(colpart, total_kb, g.Cn, gbias, st);
    CN_CHECK_LAUNCH();
  }
  rc = launch_tc<1, 1>(g, src, gp, nullptr, split > 1 ? ws : out, CN_ACT_NONE, 0.f, bn, bn_smem, dim3(mt, nt, 1), per, split, st, wn);
  if (rc) return rc;
  if (split > 1) return launch_splitk_reduce(ws, split, wn, g.Cn, nullptr, out, st);
  return CN_OK;
}

extern "C" int cn_conv_wgrad(const cn_conv_desc* d, const float* x, const float* gy, float* gw,
                             float* gbias, int impl, void* stream) {
  int rc = validate_desc(d); if (rc) return rc;
  CN_REQUIRE(x && gy && gw, CN_ERR_BAD_SHAPE, "null tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  GemmPlan g;
  rc = get_plan(d, KIND_FWD, 0, &g); if (rc) return rc;
  const size_t wn = (size_t)g.Ktot * g.Cn;
  bool tc = g.M >= 256 && g.Csrc % 4 == 0 && g.Cn % 4 == 0 && g.Cn >= 16 && g.Ktot >= 64;
  CN_REQUIRE(!(impl == CN_IMPL_TC && !tc), CN_ERR_UNSUPPORTED, "shape not eligible for the tcgen05 wgrad kernel");
  if (impl == CN_IMPL_FFMA) tc = false;
  g_last_impl = tc ? 2 : 1;
  const size_t skinny_ws = tc ? 0 : cn_skinny_wgrad_scratch(d);
  if (skinny_ws > 0) {
    float* ws = nullptr;
    rc = packed_buffer(nullptr, skinny_ws, &ws); if (rc) return rc;
    rc = cn_skinny_wgrad(d, x, gy, gw, gbias, ws, st);
    if (rc < 0) return rc;
    if (rc == 2) return CN_OK;           // weight and bias gradient both done (1x1, 3 -> 3)
  }
  bool folded = false;
  if (skinny_ws == 0 && impl != CN_IMPL_FFMA && g_fold && fold_ok(d)) {
    // folded weight gradient: gWd[l] = sum_r gy[2r+l] (x) x[r] on the low-resolution grid, then unfold to the k^nd taps
    GemmPlan gf; FoldInfo* fi = nullptr;
    int rf = get_special_plan(d, KIND_DGRAD_FOLD, &gf, &fi);
    if (rf == CN_OK && wgrad_tc_eligible(gf)) {
      float* dp = nullptr;
      rc = packed_buffer((const char*)gf.taps + 2, (size_t)gf.Ktot * gf.Cn * sizeof(float), &dp); if (rc) return rc;
      rc = launch_wgrad_tc(gf, gy, x, dp, st); if (rc) return rc;
      const int kvol = d->ksize[0] * d->ksize[1] * d->ksize[2];
      unfold_wgrad_kernel<<<dim3((d->cin + 31) / 32, (d->cout + 31) / 32, kvol), dim3(32, 8), 0, st>>>(dp, fi->d_unfold, d->cin, d->cout, gw);
      CN_CHECK_LAUNCH();
      folded = true;
      g_last_impl = 2;
    }
  }
  if (folded) {
    // done above
  } else if (skinny_ws > 0 && rc == 1) {
    // launched by skinny.cu
  } else if (tc) {
    rc = launch_wgrad_tc(g, x, gy, gw, st, gbias);
    if (rc) return rc;
    if (gbias != nullptr) return CN_OK;      // the bias gradient came out of the gradient pack pass
  } else if (d->nd == 0 && g.ntaps == 1 && g.M <= 64 && g.Cn % 4 == 0 && (((uintptr_t)gy | (uintptr_t)gw) & 15) == 0) {
    // Dense layer, a handful of rows: one thread per (input, output quad), no slabs
    const size_t threads = (size_t)g.Ktot * (g.Cn / 4);
    const bool bias_here = gbias != nullptr && ((uintptr_t)gbias & 15) == 0;
    dense_wgrad_small_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, gy, gw, bias_here ? gbias : nullptr, g.M, g.Ktot, g.Cn);
    CN_CHECK_LAUNCH();
    if (bias_here) return CN_OK;           // the bias gradient came out of the same launch
  } else if (g.ntaps == 1 && g.Csrc <= 4 && g.Cn <= 4 && g.mstride == 1 && g.ushift == 0 && g.M >= 4096) {   // 1x1, stride 1: source pixel = output pixel
    // the small weight-gradient kernels write per-block partials (slabs); sum_slabs_kernel adds them in block order
    const int blocks = 4 * num_sms();
    float* ws = nullptr;
    rc = packed_buffer(&k_ws_small, (size_t)blocks * wn * sizeof(float), &ws); if (rc) return rc;
    wgrad_flat_kernel<<<blocks, 256, 0, st>>>(x, gy, g.M, g.Csrc, g.Cn, ws);
    CN_CHECK_LAUNCH();
    sum_slabs(ws, blocks, (int)wn, gw, st);
    CN_CHECK_LAUNCH();
  } else if (g.Cn <= 4 && g.Csrc >= 8 && g.ntaps <= 16 && g.M >= 4096) {
    int warps = 16 * num_sms();
    int per = (g.M + warps - 1) / warps; if (per < 64) per = 64;
    int blocks = ((g.M + per - 1) / per + 7) / 8;
    float* ws = nullptr;
    rc = packed_buffer(&k_ws_small, (size_t)blocks * wn * sizeof(float), &ws); if (rc) return rc;
    skinny_cfew_wgrad_kernel<<<dim3(blocks, (g.Csrc + 31) / 32), 256, 0, st>>>(g, x, gy, ws, per);
    CN_CHECK_LAUNCH();
    sum_slabs(ws, blocks, (int)wn, gw, st);
    CN_CHECK_LAUNCH();
  } else if (g.Ktot <= 32 && g.Cn <= 64 && g.M >= 4096) {
    int warps = 16 * num_sms();
    int per = (g.M + warps - 1) / warps;
    int blocks = ((g.M + per - 1) / per + 7) / 8;
    float* ws = nullptr;
    rc = packed_buffer(&k_ws_small, (size_t)blocks * wn * sizeof(float), &ws); if (rc) return rc;
    skinny_kfew_wgrad_kernel<<<blocks, 256, 0, st>>>(g, x, gy, ws, per);
    CN_CHECK_LAUNCH();
    sum_slabs(ws, blocks, (int)wn, gw, st);
    CN_CHECK_LAUNCH();
  } else if ((long long)g.Ktot * g.Cn <= 256 * SKW_MAXOUT && g.M >= 4096) {
    int P = 8192 / (g.Ktot + g.Cn); if (P > 64) P = 64; if (P < 4) P = 4;
    int smem = P * (g.Ktot + g.Cn) * (int)sizeof(float);
    int blocks = 4 * num_sms();
    int per = ((g.M + blocks - 1) / blocks + P - 1) / P * P;
    blocks = (g.M + per - 1) / per;
    float* ws = nullptr;
    rc = packed_buffer(&k_ws_small, (size_t)blocks * wn * sizeof(float), &ws); if (rc) return rc;
    if (smem > 48 * 1024 && set_smem(wgrad_skinny_kernel, smem)) return CN_ERR_CUDA;
    wgrad_skinny_kernel<<<blocks, 256, smem, st>>>(g, x, gy, ws, P, per);
    CN_CHECK_LAUNCH();
    sum_slabs(ws, blocks, (int)wn, gw, st);
    CN_CHECK_LAUNCH();
  } else {
    int mt = (g.Ktot + 63) / 64, nt = (g.Cn + 63) / 64;
    int split = (2 * num_sms() + mt * nt - 1) / (mt * nt);
    int maxsplit = g.M / 128; if (maxsplit < 1) maxsplit = 1;
    if (split > maxsplit) split = maxsplit;
    int kchunk = ((g.M + split - 1) / split + 15) / 16 * 16;
    split = (g.M + kchunk - 1) / kchunk;
    float* ws = nullptr;
    if (split > 1) { rc = packed_buffer(&k_ws_splitk, (size_t)split * wn * sizeof(float), &ws); if (rc) return rc; }
    dim3 grid(mt, nt, split);
    igemm_ffma_kernel<MODE_WGRAD, 64, 64, 4, 4><<<grid, 256, 0, st>>>(g, x, gy, nullptr, split > 1 ? ws : gw, CN_ACT_NONE, 0.f, kchunk,
                                                                      split > 1 ? (long long)wn : 0ll);
    CN_CHECK_LAUNCH();
    if (split > 1) { rc = launch_splitk_reduce(ws, split, wn, g.Cn, nullptr, gw, st); if (rc) return rc; }
  }
  if (gbias != nullptr) {
    // bias gradient = column sums of gy: per-block partial rows, then sum_slabs_kernel (fixed order, no atomics); a
    // matrix that one block row covers (Dense layers: 16-32 rows) is summed straight into gbias by ONE launch
    float* ws = nullptr;
    int nslabs;
    if (g.Cn < 32) {
      size_t total = (size_t)g.M * g.Cn;
      int blocks = (int)((total / 16 + 255) / 256); if (blocks > 4 * num_sms()) blocks = 4 * num_sms(); if (blocks < 1) blocks = 1;
      int used = blocks * 256 / g.Cn * g.Cn;
      if (blocks > 1) { rc = packed_buffer(&k_ws_bias, (size_t)blocks * g.Cn * sizeof(float), &ws); if (rc) return rc; }
      colsum_narrow_kernel<<<blocks, 256, 0, st>>>(gy, total, g.Cn, used, blocks > 1 ? ws : gbias);
      nslabs = blocks;
    } else if (g.M > 512 && g.Cn % 4 == 0 && g.Cn <= 1024 && ((uintptr_t)gy & 15) == 0) {
      const int R = 256 / (g.Cn / 4);
      int blocks = 8 * num_sms();
      int per = (g.M + blocks - 1) / blocks; if (per < 4 * R) per = 4 * R;
      blocks = (g.M + per - 1) / per;
      rc = packed_buffer(&k_ws_bias, (size_t)blocks * g.Cn * sizeof(float), &ws); if (rc) return rc;
      colsum4_kernel<<<blocks, 256, 0, st>>>(gy, g.M, g.Cn, per, ws);
      nslabs = blocks;
    } else {
      int ysplit = (g.M + 511) / 512; if (ysplit > 8 * num_sms()) ysplit = 8 * num_sms(); if (ysplit < 1) ysplit = 1;
      dim3 grid((g.Cn + 31) / 32, ysplit), block(32, 8);
      if (ysplit > 1) { rc = packed_buffer(&k_ws_bias, (size_t)ysplit * g.Cn * sizeof(float), &ws); if (rc) return rc; }
      colsum_kernel<<<grid, block, 0, st>>>(gy, g.M, g.Cn, ysplit > 1 ? ws : gbias);
      nslabs = ysplit;
    }
    CN_CHECK_LAUNCH();
    if (nslabs > 1) {
      sum_slabs(ws, nslabs, g.Cn, gbias, st);
      CN_CHECK_LAUNCH();
    }
  }
  (void)conv_numel_x;
  return CN_OK;
}

#ifdef CN_TEST_HOOKS
// ------------------------------------------------------------------------------------------------
// Host evaluation of a plan with scalar loops.  TEST HOOK ONLY (tests/test_plan_host.py): lets the
// CPU-only test suite check the integer geometry (SAME padding, phases, fused upsample, tap tables)
// that the kernels share, without a GPU.  Not used by any product path.
// kind: 0 forward, 1 dgrad, 2 wgrad.  All pointers are HOST pointers here.
// ------------------------------------------------------------------------------------------------
// host evaluation of one (possibly phased) pixel-mode plan
static void host_eval_pixel(const GemmPlan& p0, const std::vector<int2>& taps, const float* a, const float* b, float* out) {
  const int nph = p0.nphase > 1 ? p0.nphase : 1;
  for (int ph = 0; ph < nph; ++ph) {
    GemmPlan p = p0;
    int t0 = 0;
    if (p0.nphase > 1) {
      t0 = p0.ph_tap0[ph]; p.ntaps = p0.ph_ntaps[ph];
      p.ooff[0] = p0.ph_ooff[ph] & 1; p.ooff[1] = (p0.ph_ooff[ph] >> 1) & 1; p.ooff[2] = (p0.ph_ooff[ph] >> 2) & 1;
    }
    for (int m = 0; m < p.M; ++m) {
      RowInfo ri = decode_row(p, m);
      size_t ro = (size_t)dest_pixel(p, m) * p.Cn;
      for (int n = 0; n < p.Cn; ++n) {
        double acc = 0;
        for (int kt = 0; kt < p.ntaps; ++kt) {
          uint32_t sp = src_pixel(p, ri, taps[t0 + kt].x);
          if (sp == 0xffffffffu) continue;
          for (int c = 0; c < p.Csrc; ++c)
            acc += (double)a[(size_t)sp * p.Csrc + c] * b[(size_t)taps[t0 + kt].y + (size_t)c * p.wsc + (size_t)n * p.wsn];
        }
        out[ro + n] = (float)acc;
      }
    }
  }
}

static void host_fold_weights(const cn_conv_desc* d, const FoldInfo& fi, const float* w, std::vector<float>& wf) {
  const int cc = d->cin * d->cout;
  wf.assign((size_t)fi.nfold * cc, 0.f);
  for (int f = 0; f < fi.nfold; ++f) {
    int lo[3], hi[3]; fold_decode(fi.spec[f], lo, hi);
    for (int t0 = lo[0]; t0 <= hi[0]; ++t0) for (int t1 = lo[1]; t1 <= hi[1]; ++t1) for (int t2 = lo[2]; t2 <= hi[2]; ++t2)
      for (int i = 0; i < cc; ++i) wf[(size_t)f * cc + i] += w[(size_t)((t0 * d->ksize[1] + t1) * d->ksize[2] + t2) * cc + i];
  }
}

// kind: 0 forward, 1 dgrad (per-phase plans), 2 wgrad, 3 folded forward, 4 folded dgrad, 5 folded wgrad,
//       6 stride-2 dgrad with all parity phases in one phased plan
extern "C" int cn_debug_conv_host(const cn_conv_desc* d, int kind, const float* a, const float* b, float* out) {
  int rc = validate_desc(d); if (rc) return rc;
  if (kind >= 3 && kind <= 5) {
    CN_REQUIRE(fold_ok(d), CN_ERR_UNSUPPORTED, "descriptor has no folded form");
    GemmPlan p; std::vector<int2> taps; FoldInfo fi;
    rc = build_fold_plan(d, kind == 3 ? KIND_FWD_FOLD : KIND_DGRAD_FOLD, &p, taps, &fi); if (rc) return rc;
    if (kind == 5) {       // a = x, b = gy -> out = gw
      std::vector<double> dp((size_t)p.Ktot * p.Cn, 0.0);
      for (int r = 0; r < p.Ktot; ++r) {
        const int kt = r / p.Csrc, c = r % p.Csrc;
        for (int m = 0; m < p.M; ++m) {
          uint32_t sp = src_pixel(p, decode_row(p, m), taps[kt].x);
          if (sp == 0xffffffffu) continue;
          for (int n = 0; n < p.Cn; ++n) dp[(size_t)r * p.Cn + n] += (double)b[(size_t)sp * p.Csrc + c] * a[(size_t)m * p.Cn + n];
        }
      }
      const int kvol = d->ksize[0] * d->ksize[1] * d->ksize[2];
      for (int t = 0; t < kvol; ++t)
        for (int ci = 0; ci < d->cin; ++ci)
          for (int co = 0; co < d->cout; ++co) {
            double acc = 0;
            for (int j = 0; j < 8; ++j) { int f = fi.unfold[(size_t)t * 8 + j]; if (f >= 0) acc += dp[((size_t)f * d->cout + co) * d->cin + ci]; }
            out[((size_t)t * d->cin + ci) * d->cout + co] = (float)acc;
          }
      return CN_OK;
    }
    std::vector<float> wf;
    host_fold_weights(d, fi, b, wf);
    host_eval_pixel(p, taps, a, wf.data(), out);
    return CN_OK;
  }
  if (kind == 6) {
    GemmPlan p; std::vector<int2> taps;
    rc = build_s2all_plan(d, &p, taps); if (rc) return rc;
    host_eval_pixel(p, taps, a, b, out);
    return CN_OK;
  }
  const int nphase = (kind == 1 && d->stride == 2) ? (1 << d->nd) : 1;
  for (int ph = 0; ph < nphase; ++ph) {
    GemmPlan p; std::vector<int2> taps;
    rc = build_plan(d, kind == 1 ? KIND_DGRAD : KIND_FWD, ph, &p, taps); if (rc) return rc;
    if (kind == 2) {
      for (int r = 0; r < p.Ktot; ++r) {
        int kt = r / p.Csrc, c = r % p.Csrc;
        for (int n = 0; n < p.Cn; ++n) {
          double acc = 0;
          for (int m = 0; m < p.M; ++m) {
            uint32_t sp = src_pixel(p, decode_row(p, m), taps[kt].x);
            if (sp != 0xffffffffu) acc += (double)a[(size_t)sp * p.Csrc + c] * b[(size_t)m * p.Cn + n];
          }
          out[(size_t)r * p.Cn + n] = (float)acc;
        }
      }
    } else {
      for (int m = 0; m < p.M; ++m) {
        RowInfo ri = decode_row(p, m);
        size_t ro = (size_t)dest_pixel(p, m) * p.Cn;
        for (int n = 0; n < p.Cn; ++n) {
          double acc = 0;
          for (int kt = 0; kt < p.ntaps; ++kt) {
            uint32_t sp = src_pixel(p, ri, taps[kt].x);
            if (sp == 0xffffffffu) continue;
            for (int c = 0; c < p.Csrc; ++c)
              acc += (double)a[(size_t)sp * p.Csrc + c] * b[(size_t)taps[kt].y + (size_t)c * p.wsc + (size_t)n * p.wsn];
          }
          out[ro + n] = (float)acc;
        }
      }
    }
  }
  return CN_OK;
}
#endif  // CN_TEST_HOOKS
