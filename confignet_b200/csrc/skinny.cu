// Skinny convolution layers of ConfigNet's hot path: three channels on one side, so the layer is bound by
// the HBM traffic of its WIDE tensor (D.block0 3->48 AI 11, VGG block1_conv1 3->64 AI 13,
// G.map_final 32->3 AI 70 FLOP/B), not by the tensor pipe.  Every kernel here follows one recipe:
//   * a block owns one image row (or row pair) and stages the rows it needs in shared memory with coalesced
//     128-bit loads, each HBM byte read once per block (neighbouring rows come from L2);
//   * threads own register tiles chosen so that FFMA issue, not shared-memory wavefronts, bounds the inner loop;
//   * stores are whole 32-byte sectors;
//   * weight gradients are accumulated in registers by persistent blocks (grid = 2 x SM count), written as
//     per-block partials and summed in a fixed order by a second kernel: deterministic, no atomics.
//
//   c3_*    : Conv2D(3 -> 48|64, k3, stride 1|2, SAME)            hologan_discriminator.py:28-40 (DiscrBlock 0),
//                                                                  hologan_discriminator.py:77-97 (regressor trunk),
//                                                                  perceptual_loss.py:19-24 (VGG block1_conv1)
//   up4c3_* : UpSampling2D(2) + Conv2D(32 -> 3, k4, SAME) + tanh   hologan_generator.py:101,170-172 (map_final)
//             evaluated on the LOW-resolution tensor: output pixel (2R+dy, 2C+dx) only sees source pixels
//             (R+sy, C+sx), sy,sx in {-1,0,1}, and the taps that land on the same source pixel are pre-summed
//             (sub-pixel folding, SURVEY.md section 7 hard part 5): 25 folded taps per 2x2 output cell instead of 64.
#include "common.cuh"
#include <stdlib.h>
#include <map>

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& a, float x, const float4& w) {
  a.x = fmaf(x, w.x, a.x); a.y = fmaf(x, w.y, a.y); a.z = fmaf(x, w.z, a.z); a.w = fmaf(x, w.w, a.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// cp.async staging (global -> shared without a register round trip): every piece a thread issues is in flight at once.  The
// plain "load, store to shared" loops below it replaced kept ONE load in flight per thread (the compiler places each store
// right behind its load), i.e. a chain of ~10 dependent HBM / L2 round trips per block before the first FFMA.
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// c3 forward: block = (128 output pixels of one output row); thread = (pixel pair, channel quarter)
// ------------------------------------------------------------------------------------------------
template <int S, int COUT, int PXT>
__global__ void __launch_bounds__(512 / PXT)
c3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              float* __restrict__ y, int H, int W, int OH, int OW, int pby, int pbx, int act, float alpha) {
  // thread = (PXT consecutive pixels, channel quarter).  The kernel is bound by shared-memory wavefronts, not FFMAs (ncu:
  // L1 data pipe 87 % busy, FMA pipe 40 % with two pixels per thread - a weight float4 costs four wavefronts and fed only
  // 8 FFMAs): PXT = 4 pixels per thread halve the weight reads per FFMA.
  constexpr int NT = 512 / PXT;           // threads per block: (128 / PXT pixel groups) x 4 channel quarters
  constexpr int CPT = COUT / 16;          // float4 channel groups per thread
  constexpr int C4 = COUT / 4;
  constexpr int NCOL = 127 * S + 3;       // input columns under 128 output pixels
  constexpr int RS = NCOL * 3;
  __shared__ __align__(16) float s_w[27 * COUT];
  __shared__ float s_x[3 * RS];
  const int tid = threadIdx.x, n = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * 128;
  for (int i = tid; i < 27 * C4; i += NT) cp_async16(s_w + 4 * i, w + 4 * i);
  const int ix0 = ox0 * S - pbx;
  for (int i = tid; i < 3 * RS; i += NT) {
    const int r = i / RS, j = i - r * RS, col = j / 3;
    const int iy = oy * S + r - pby, ix = ix0 + col;
    const bool in = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
    cp_async4_zfill(s_x + i, in ? x + ((long long)(n * H + iy) * W + ix0) * 3 + j : x, in);
  }
  cp_async_wait_all();
  __syncthreads();
  const int q = tid & 3, pp = tid >> 2, la = PXT * pp * S;
  float4 acc[PXT][CPT];
#pragma unroll
  for (int p = 0; p < PXT; ++p)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[p][j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* w4 = reinterpret_cast<const float4*>(s_w);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int e = 0; e < 9; ++e) {           // e = kx*3 + ci: 9 contiguous floats of the input row
      float xs[PXT];
#pragma unroll
      for (int p = 0; p < PXT; ++p) xs[p] = s_x[ky * RS + (la + p * S) * 3 + e];
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float4 wv = w4[(ky * 9 + e) * C4 + q + 4 * j];
#pragma unroll
        for (int p = 0; p < PXT; ++p) fma4(acc[p][j], xs[p], wv);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PXT; ++p) {
    const int ox = ox0 + PXT * pp + p;
    if (ox >= OW) continue;
    float* out = y + ((size_t)(n * OH + oy) * OW + ox) * COUT;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int f = q + 4 * j;
      float4 v = acc[p][j];
      if (bias != nullptr) { const float4 b = ldg4(bias + 4 * f); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
      v.x = cn_apply_act(v.x, act, alpha); v.y = cn_apply_act(v.y, act, alpha);
      v.z = cn_apply_act(v.z, act, alpha); v.w = cn_apply_act(v.w, act, alpha);
      *reinterpret_cast<float4*>(out + 4 * f) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// c3 input gradient: gx[iy][ix][ci] = sum_{ky,kx,co} gy[(iy+pb-ky)/S][(ix+pb-kx)/S][co] * w[ky][kx][ci][co]
// (terms whose division is not exact do not exist).  stride 2 (even H, W; pb = 0): thread = 2x2 cell of gx;
// stride 1 (pb = 1): thread = 1x2 pixels.  The four lanes of a cell split the output channels of gy and are
// summed with two shuffles.  PS = padded pixel stride of the staged gy rows (conflict-free float4 reads).
// ------------------------------------------------------------------------------------------------
template <int S, int COUT, int PS>
__global__ void __launch_bounds__(256)
c3_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ gx,
                int H, int W, int OH, int OW) {
  constexpr int TY = (S == 2) ? 2 : 1, TX = 2;
  constexpr int PB = (S == 2) ? 0 : 1;
  constexpr int CELLS = 64, GXW = CELLS * TX;
  constexpr int GYW = (S == 2) ? CELLS + 1 : GXW + 2;
  constexpr int GYR = (S == 2) ? 2 : 3;
  constexpr int CPT = COUT / 16, C4 = COUT / 4, PS4 = PS / 4;
  extern __shared__ __align__(16) float sm[];
  float* s_w = sm;                          // [27][COUT]
  float* s_gy = sm + 27 * COUT;             // [GYR][GYW][PS]
  const int tid = threadIdx.x, n = blockIdx.z, by = blockIdx.y, x0 = blockIdx.x * GXW;
  const int iy0 = by * TY;
  const int gr0 = (S == 2) ? by - 1 : iy0 - 1, gc0 = (S == 2) ? x0 / 2 - 1 : x0 - 1;
  for (int i = tid; i < 27 * C4; i += 256) reinterpret_cast<float4*>(s_w)[i] = ldg4(w + 4 * i);
  for (int i = tid; i < GYR * GYW * C4; i += 256) {
    const int f = i % C4, pc = i / C4, col = pc % GYW, row = pc / GYW;
    const int gr = gr0 + row, gc = gc0 + col;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((unsigned)gr < (unsigned)OH && (unsigned)gc < (unsigned)OW) v = ldg4(gy + ((size_t)(n * OH + gr) * OW + gc) * COUT + 4 * f);
    reinterpret_cast<float4*>(s_gy)[(row * GYW + col) * PS4 + f] = v;
  }
  __syncthreads();
  const int q = tid & 3, cell = tid >> 2;
  float acc[TY][TX][3];
#pragma unroll
  for (int a = 0; a < TY; ++a)
#pragma unroll
    for (int b = 0; b < TX; ++b)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[a][b][c] = 0.f;
  const float4* w4 = reinterpret_cast<const float4*>(s_w);
  const float4* g4 = reinterpret_cast<const float4*>(s_gy);
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int f = q + 4 * j;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int dy = 0; dy < TY; ++dy) {
        if (((dy + PB - ky + 2 * S) % S) != 0) continue;
        const int lrow = (S == 2) ? (dy - ky) / 2 + 1 : 2 - ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int dx = 0; dx < TX; ++dx) {
            if (((dx + PB - kx + 2 * S) % S) != 0) continue;
            const int lcol = (S == 2) ? cell + (dx - kx) / 2 + 1 : cell * TX + dx + 2 - kx;
            const float4 g = g4[(lrow * GYW + lcol) * PS4 + f];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
              acc[dy][dx][ci] = dot4(g, w4[((ky * 3 + kx) * 3 + ci) * C4 + f], acc[dy][dx][ci]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < TY; ++a)
#pragma unroll
    for (int b = 0; b < TX; ++b)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = acc[a][b][c];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        acc[a][b][c] = v;
      }
  // lane q of the cell writes pixel (dy, dx) = (q >> 1, q & 1) (stride 2) or (0, q) for q < 2 (stride 1)
  const int dy = (S == 2) ? (q >> 1) : 0, dx = q & 1;
  if (S == 2 || q < 2) {
    const int iy = iy0 + dy, ix = x0 + cell * TX + dx;
    if (iy < H && ix < W) {
      float* out = gx + ((size_t)(n * H + iy) * W + ix) * 3;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
      for (int a = 0; a < TY; ++a)
#pragma unroll
        for (int b = 0; b < TX; ++b)
          if (a == dy && b == dx) { v0 = acc[a][b][0]; v1 = acc[a][b][1]; v2 = acc[a][b][2]; }
      out[0] = v0; out[1] = v1; out[2] = v2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// c3 weight gradient: gw[ky][kx][ci][co] = sum_pixels x[S*oy+ky-pb][S*ox+kx-pb][ci] * gy[oy][ox][co].
// Persistent blocks walk output rows; thread = (pixel group g, kernel row ky, 8 output channels) holds a
// 9 x 8 register tile; the 14 pixel groups of a block are folded in a fixed order at the end.
// ------------------------------------------------------------------------------------------------
// The rows are staged with cp.async into TWO buffers: row i + 1 is in flight while row i is multiplied (one buffer and plain
// loads left every row iteration waiting on ~9 dependent scalar loads per thread: 93-113 us per launch for a layer whose
// FFMAs take ~40 us, profiles/r02_launches_s2c_summary.txt).
template <int S, int COUT>
__global__ void __launch_bounds__(256, 2)
c3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ part,
                int B, int H, int W, int OH, int OW, int pby, int pbx) {
  constexpr int NC8 = COUT / 8, ITEMS = 3 * NC8, GROUPS = 256 / ITEMS, C4 = COUT / 4;
  constexpr int NCOL = 127 * S + 3, RS = NCOL * 3, XBUF = (3 * RS + 8 + 3) & ~3;
  extern __shared__ __align__(16) float sm_w[];          // [2][128 * COUT] gy rows, then [2][XBUF] input rows
  float* s_gy = sm_w;
  float* s_xb = sm_w + 2 * 128 * COUT;
  const int tid = threadIdx.x;
  const int g = tid / ITEMS, item = tid - g * ITEMS, ky = item / NC8, c8 = item - ky * NC8;
  const bool active = g < GROUPS;
  float acc[9][8];
#pragma unroll
  for (int a = 0; a < 9; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  const int nseg = (OW + 127) / 128;
  const int nrows = B * OH * nseg;
  auto stage = [&](int row, int buf) {
    const int seg = row % nseg, r2 = row / nseg, oy = r2 % OH, n = r2 / OH;
    const int ox0 = seg * 128, npx = min(128, OW - ox0);
    const float* grow = gy + ((size_t)(n * OH + oy) * OW + ox0) * COUT;
    float* dg = s_gy + buf * 128 * COUT;
    for (int i = tid; i < npx * C4; i += 256) cp_async16(dg + 4 * i, grow + 4 * i);
    const int ix0 = ox0 * S - pbx;
    float* dx = s_xb + buf * XBUF;
    for (int i = tid; i < 3 * RS; i += 256) {
      const int r = i / RS, j = i - r * RS, col = j / 3;
      const int iy = oy * S + r - pby, ix = ix0 + col;
      const bool in = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
      cp_async4_zfill(dx + i, in ? x + ((long long)(n * H + iy) * W + ix0) * 3 + j : x, in);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if ((int)blockIdx.x < nrows) stage(blockIdx.x, 0);
  for (int row = blockIdx.x; row < nrows; row += gridDim.x, buf ^= 1) {
    const int seg = row % nseg;
    const int npx = min(128, OW - seg * 128);
    const int next = row + gridDim.x;
    if (next < nrows) {
      stage(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // this row has landed, the next one stays in flight
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      const float* bg = s_gy + buf * 128 * COUT;
      const float* bx = s_xb + buf * XBUF;
      for (int p = g; p < npx; p += GROUPS) {
        const float4 ga = reinterpret_cast<const float4*>(bg)[p * C4 + 2 * c8];
        const float4 gb = reinterpret_cast<const float4*>(bg)[p * C4 + 2 * c8 + 1];
        const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float* xr = bx + ky * RS + p * S * 3;
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const float xv = xr[a];
#pragma unroll
          for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(xv, gv[b], acc[a][b]);
        }
      }
    }
    __syncthreads();                                       // this buffer is refilled by the next iteration's stage()
  }
  // fixed-order fold of the pixel groups, then one partial per block
  __syncthreads();
  float* s_acc = s_gy;                      // 27*COUT floats
  for (int g2 = 0; g2 < GROUPS; ++g2) {
    if (active && g == g2) {
#pragma unroll
      for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int o = (ky * 9 + a) * COUT + c8 * 8 + b;
          s_acc[o] = (g2 == 0) ? acc[a][b] : s_acc[o] + acc[a][b];
        }
    }
    __syncthreads();
  }
  for (int i = tid; i < 27 * COUT; i += 256) part[(size_t)blockIdx.x * 27 * COUT + i] = s_acc[i];
}

// out[i] = sum_b part[b][i], deterministic: block (32, 8) - thread (x, y) adds the partials y, y+8, ... of output
// 32*blockIdx.x + x in order, then the 8 rows are folded in order (a single thread walking all ~300 partials is a
// chain of dependent L2 round trips: 23 us for 1296 outputs, measured)
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ part, int nblocks, int n, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (i < n)
    for (int b = threadIdx.y; b < nblocks; b += 8) a += part[(size_t)b * n + i];
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    out[i] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// up4c3: UpSampling2D(2) + Conv2D(CIN -> 3, k4, SAME), folded onto the low-resolution tensor.
// With lr = 2 - 2*sy + dy (0..4):  y[2R+dy][2C+dx] = sum_{sy,sx} x[R+sy][C+sx] . Wd[lr][lc]
//                                   gx[r][c]       = sum_{lr,lc} gy[2r-2+lr][2c-2+lc] . Wd[lr][lc]
// Wd[lr][lc] = sum_{ty in T(lr), tx in T(lc)} w[ty][tx],  T = {3}, {2,3}, {1,2}, {0,1}, {0}.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fold_taps(int l, int& t0, int& t1) {   // T(l) as [t0, t1]
  t0 = l == 0 ? 3 : (l == 1 ? 2 : (l == 2 ? 1 : 0));
  t1 = l == 0 ? 3 : (l == 1 ? 3 : (l == 2 ? 2 : (l == 3 ? 1 : 0)));
}

// s_wd[(lr*5+lc)][ci] = float4(Wd[..][ci][0..2], 0)
template <int CIN>
__device__ __forceinline__ void build_folded(const float* __restrict__ w, float4* s_wd, int tid, int nthreads) {
  for (int i = tid; i < 25 * CIN; i += nthreads) {
    const int ci = i % CIN, l = i / CIN, lr = l / 5, lc = l - lr * 5;
    int y0, y1, x0, x1;
    fold_taps(lr, y0, y1); fold_taps(lc, x0, x1);
    float a[3] = {0.f, 0.f, 0.f};
    for (int ty = y0; ty <= y1; ++ty)
      for (int tx = x0; tx <= x1; ++tx)
        for (int co = 0; co < 3; ++co) a[co] += __ldg(w + ((size_t)(ty * 4 + tx) * CIN + ci) * 3 + co);
    s_wd[i] = make_float4(a[0], a[1], a[2], 0.f);
  }
}

// forward: block = one source row (128 cells); warp parity = dy; thread = (cell, dy) -> 2 output pixels x 3 channels
template <int CIN>
__global__ void __launch_bounds__(256)
up4c3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y, int H, int W, int act, float alpha) {
  constexpr int PS = CIN + 4, PS4 = PS / 4, XW = 130, CI4 = CIN / 4;
  extern __shared__ __align__(16) float sm[];
  float4* s_wd = reinterpret_cast<float4*>(sm);               // [25][CIN]
  float4* s_x = s_wd + 25 * CIN;                              // [3][XW][PS4]
  const int tid = threadIdx.x, n = blockIdx.z, R = blockIdx.y, c0 = blockIdx.x * 128;
  for (int i = tid; i < 3 * XW * CI4; i += 256) {
    const int f = i % CI4, pc = i / CI4, col = pc % XW, row = pc / XW;
    const int r = R - 1 + row, c = c0 - 1 + col;
    const bool in = (unsigned)r < (unsigned)H && (unsigned)c < (unsigned)W;
    cp_async16_zfill(reinterpret_cast<float*>(s_x + (row * XW + col) * PS4 + f), in ? x + ((size_t)(n * H + r) * W + c) * CIN + 4 * f : x, in);
  }
  build_folded<CIN>(w, s_wd, tid, 256);                       // overlaps the copies in flight
  cp_async_wait_all();
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int dy = warp & 1, cell = (warp >> 1) * 32 + lane;
  float acc[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
  for (int sy = -1; sy <= 1; ++sy) {
    const int lr = 2 - 2 * sy + dy;
    if (lr > 4) continue;                                      // warp-uniform (dy = warp parity)
#pragma unroll
    for (int sx = -1; sx <= 1; ++sx) {
      const float4* xp = s_x + ((sy + 1) * XW + cell + 1 + sx) * PS4;
      const float4* w0 = s_wd + (lr * 5 + 2 - 2 * sx) * CIN;         // dx = 0: lc = 2 - 2 sx
      const float4* w1 = s_wd + (lr * 5 + 3 - 2 * sx) * CIN;         // dx = 1: lc = 3 - 2 sx (sx = -1 -> 5: absent)
#pragma unroll
      for (int f = 0; f < CI4; ++f) {
        const float4 xv = xp[f];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 a = w0[4 * f + e];
          acc[0][0] = fmaf(xs[e], a.x, acc[0][0]); acc[0][1] = fmaf(xs[e], a.y, acc[0][1]); acc[0][2] = fmaf(xs[e], a.z, acc[0][2]);
          if (sx >= 0) {
            const float4 b = w1[4 * f + e];
            acc[1][0] = fmaf(xs[e], b.x, acc[1][0]); acc[1][1] = fmaf(xs[e], b.y, acc[1][1]); acc[1][2] = fmaf(xs[e], b.z, acc[1][2]);
          }
        }
      }
    }
  }
  const int C = c0 + cell;
  if (C < W) {
    float* out = y + ((size_t)(n * 2 * H + 2 * R + dy) * (2 * W) + 2 * C) * 3;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        float v = acc[dx][co] + (bias != nullptr ? __ldg(bias + co) : 0.f);
        out[dx * 3 + co] = cn_apply_act(v, act, alpha);
      }
  }
}

// input gradient: block = one source row, 64 pixels; thread = (pixel, 8 source channels)
template <int CIN>
__global__ void __launch_bounds__(256)
up4c3_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ gx, int H, int W) {
  constexpr int NG = CIN / 8;               // channel groups per pixel
  constexpr int PXB = 256 / NG;             // pixels per block
  constexpr int GW = 2 * PXB + 3;           // gy columns 2c0-2 .. 2(c0+PXB-1)+2
  extern __shared__ __align__(16) float sm[];
  float* s_wt = sm;                         // [25][3][CIN]  (transposed: source channels contiguous)
  float* s_gy = sm + 25 * 3 * CIN;          // [5][GW][3]
  const int tid = threadIdx.x, n = blockIdx.z, r = blockIdx.y, c0 = blockIdx.x * PXB;
  {
    const int OHh = 2 * H, OWw = 2 * W;
    for (int i = tid; i < 5 * GW * 3; i += 256) {           // cp.async: all pieces in flight while the kernel is folded below
      const int row = i / (GW * 3), j = i - row * GW * 3, col = j / 3;
      const int oy = 2 * r - 2 + row, ox = 2 * c0 - 2 + col;
      const bool in = (unsigned)oy < (unsigned)OHh && (unsigned)ox < (unsigned)OWw;
      cp_async4_zfill(s_gy + i, in ? gy + ((long long)(n * OHh + oy) * OWw + (2 * c0 - 2)) * 3 + j : gy, in);
    }
  }
  for (int i = tid; i < 25 * CIN; i += 256) {
    const int ci = i % CIN, l = i / CIN, lr = l / 5, lc = l - lr * 5;
    int y0, y1, x0, x1;
    fold_taps(lr, y0, y1); fold_taps(lc, x0, x1);
    float a[3] = {0.f, 0.f, 0.f};
    for (int ty = y0; ty <= y1; ++ty)
      for (int tx = x0; tx <= x1; ++tx)
        for (int co = 0; co < 3; ++co) a[co] += __ldg(w + ((size_t)(ty * 4 + tx) * CIN + ci) * 3 + co);
    for (int co = 0; co < 3; ++co) s_wt[(l * 3 + co) * CIN + ci] = a[co];
  }
  cp_async_wait_all();
  __syncthreads();
  const int grp = tid % NG, px = tid / NG;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
  for (int lr = 0; lr < 5; ++lr) {
#pragma unroll
    for (int lc = 0; lc < 5; ++lc) {
      const float* gp = s_gy + (lr * GW + 2 * px + lc) * 3;
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        const float g = gp[co];
        const float4* wp = reinterpret_cast<const float4*>(s_wt + ((lr * 5 + lc) * 3 + co) * CIN + grp * 8);
        fma4(a0, g, wp[0]);
        fma4(a1, g, wp[1]);
      }
    }
  }
  const int c = c0 + px;
  if (c < W) {
    float4* out = reinterpret_cast<float4*>(gx + ((size_t)(n * H + r) * W + c) * CIN + grp * 8);
    out[0] = a0; out[1] = a1;
  }
}

// weight gradient, folded: gWd[lr][lc][ci][co] = sum_{n,r,c} x[r][c][ci] * gy[2r-2+lr][2c-2+lc][co].
// Persistent blocks walk source rows; lane = source channel (CIN = 32), warp wv owns folded taps wv, wv+8, ...
template <int CIN>
__global__ void __launch_bounds__(256, 2)
up4c3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ part, int B, int H, int W) {
  static_assert(CIN == 32, "lane = source channel");
  constexpr int SEG = 128, GW = 2 * SEG + 3;
  __shared__ __align__(16) float s_x[SEG * CIN];
  __shared__ __align__(16) float4 s_gy[5 * GW];              // gy pixel padded to float4
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a) { acc[a][0] = 0.f; acc[a][1] = 0.f; acc[a][2] = 0.f; }
  const int nseg = (W + SEG - 1) / SEG, nrows = B * H * nseg;
  const int OHh = 2 * H, OWw = 2 * W;
  for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int seg = row % nseg, r2 = row / nseg, r = r2 % H, n = r2 / H;
    const int c0 = seg * SEG, npx = min(SEG, W - c0);
    __syncthreads();
    const float* xrow = x + ((size_t)(n * H + r) * W + c0) * CIN;
    for (int i = tid; i < npx * (CIN / 4); i += 256) reinterpret_cast<float4*>(s_x)[i] = ldg4(xrow + 4 * i);
    for (int i = tid; i < 5 * GW; i += 256) {
      const int rr = i / GW, col = i - rr * GW;
      const int oy = 2 * r - 2 + rr, ox = 2 * c0 - 2 + col;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)oy < (unsigned)OHh && (unsigned)ox < (unsigned)OWw) {
        const float* gp = gy + ((size_t)(n * OHh + oy) * OWw + ox) * 3;
        v = make_float4(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), 0.f);
      }
      s_gy[i] = v;
    }
    __syncthreads();
    for (int p = 0; p < npx; ++p) {
      const float xv = s_x[p * CIN + lane];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int l = warp + 8 * a;
        if (l < 25) {
          const int lr = l / 5, lc = l - lr * 5;
          const float4 g = s_gy[lr * GW + 2 * p + lc];
          acc[a][0] = fmaf(xv, g.x, acc[a][0]); acc[a][1] = fmaf(xv, g.y, acc[a][1]); acc[a][2] = fmaf(xv, g.z, acc[a][2]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int l = warp + 8 * a;
    if (l < 25)
      for (int co = 0; co < 3; ++co) part[(size_t)blockIdx.x * 25 * CIN * 3 + (l * CIN + lane) * 3 + co] = acc[a][co];
  }
}

// gw[ty][tx][ci][co] = sum over the folded taps that contain (ty, tx), partials summed in a fixed order
// (block (32, 8) as in sum_partials_kernel)
template <int CIN>
__global__ void __launch_bounds__(256)
up4c3_unfold_kernel(const float* __restrict__ part, int nblocks, float* __restrict__ gw) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  const bool ok = i < 16 * CIN * 3;
  const int e = i % (CIN * 3), t = i / (CIN * 3), ty = t >> 2, tx = t & 3;
  // T^-1: tap 0 -> {3,4}, 1 -> {2,3}, 2 -> {1,2}, 3 -> {0,1}
  const int lr0 = 3 - ty, lc0 = 3 - tx;
  float a = 0.f;
  if (ok)
    for (int b = threadIdx.y; b < nblocks; b += 8) {
      const float* p = part + (size_t)b * 25 * CIN * 3;
      a += (p[((lr0 * 5 + lc0)) * CIN * 3 + e] + p[((lr0 * 5 + lc0 + 1)) * CIN * 3 + e]) +
           (p[(((lr0 + 1) * 5 + lc0)) * CIN * 3 + e] + p[(((lr0 + 1) * 5 + lc0 + 1)) * CIN * 3 + e]);
    }
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && ok) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    gw[i] = s;
  }
}

int g_sms = 0;
int sm_count() {
  if (g_sms == 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

// ------------------------------------------------------------------------------------------------
// c3k: the c3 forward and stride-2 input gradient with the KERNEL IN THE CONSTANT BANK.  Per output pixel these layers
// do 27 x COUT FMAs on 27 input values: the layer is FFMA-bound (0.68 GFMA for D.block0 at batch 32 = 19 us at the FP32
// peak, the same 19 us its 125 MB take at the HBM rate) if - and only if - the inner loop is nothing but FFMAs.  The
// first-generation kernels above read the weights from shared memory (3 - 4 LDS.128 per 12 - 24 FFMAs: shared-memory
// bound, 72 / 109 us measured).  Here a thread owns ALL output channels of its pixel(s) and every FFMA takes its weight
// as a constant-bank operand c[bank][imm] (the 27 x COUT kernel is copied into c_c3w, stream-ordered, before the launch):
// no weight loads at all, 27 (forward) or 48 (gradient) shared-memory reads per 1296 FFMAs.
//   forward : thread = one output pixel x COUT channels; the tile leaves through shared memory as whole lines
//   dgrad s2: thread = one 2x2 cell of gx; walks its 4 neighbouring gy pixels (48 channels in registers each)
// ------------------------------------------------------------------------------------------------
__constant__ float c_c3w[27 * 64];

template <int S, int COUT>
__global__ void __launch_bounds__(128)
c3k_fwd_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ y,
               int H, int W, int OH, int OW, int pby, int pbx, int act, float alpha) {
  constexpr int C4 = COUT / 4, PSO = COUT + 4;          // padded pixel stride of the output tile: conflict-free float4 rows
  constexpr int NCOL = 127 * S + 3, RS = NCOL * 3;
  extern __shared__ __align__(16) float sm[];
  float* s_x = sm;                                       // [3][RS]
  float* s_o = sm + ((3 * RS + 3) & ~3);                 // [128][PSO]
  const int tid = threadIdx.x, n = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * 128;
  const int ix0 = ox0 * S - pbx;
  for (int i = tid; i < 3 * RS; i += 128) {
    const int r = i / RS, j = i - r * RS, col = j / 3;
    const int iy = oy * S + r - pby, ix = ix0 + col;
    const bool in = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
    cp_async4_zfill(s_x + i, in ? x + ((long long)(n * H + iy) * W + ix0) * 3 + j : x, in);
  }
  cp_async_wait_all();
  __syncthreads();
  float xv[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int e = 0; e < 9; ++e) xv[ky * 9 + e] = s_x[ky * RS + tid * S * 3 + e];
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = bias != nullptr ? __ldg(bias + co) : 0.f;
#pragma unroll
  for (int k = 0; k < 27; ++k)
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = fmaf(xv[k], c_c3w[k * COUT + co], acc[co]);
#pragma unroll
  for (int f = 0; f < C4; ++f)
    *reinterpret_cast<float4*>(s_o + tid * PSO + 4 * f) =
        make_float4(cn_apply_act(acc[4 * f], act, alpha), cn_apply_act(acc[4 * f + 1], act, alpha),
                    cn_apply_act(acc[4 * f + 2], act, alpha), cn_apply_act(acc[4 * f + 3], act, alpha));
  __syncthreads();
  float* out = y + ((size_t)(n * OH + oy) * OW + ox0) * COUT;       // the 128 pixels of this block are contiguous in y
  const int npx = min(128, OW - ox0);
  for (int i = tid; i < npx * C4; i += 128) {
    const int px = i / C4, f = i - px * C4;
    *reinterpret_cast<float4*>(out + 4 * (size_t)i) = *reinterpret_cast<const float4*>(s_o + px * PSO + 4 * f);
  }
}

// stride 2, even H and W, TF-SAME (pad 0 in front): gx[2a+dy][2b+dx][ci] = sum over ky = dy (mod 2), kx = dx (mod 2) of
// gy[a + (dy-ky)/2][b + (dx-kx)/2][:] . w[ky][kx][ci][:]
template <int COUT>
__global__ void __launch_bounds__(128)
c3k_dgrad_s2_kernel(const float* __restrict__ gy, float* __restrict__ gx, int H, int W, int OH, int OW) {
  constexpr int C4 = COUT / 4, PS = COUT + 4, GYW = 129;
  extern __shared__ __align__(16) float sm[];            // [2][GYW][PS]: gy rows a-1, a; columns b0-1 .. b0+127
  const int tid = threadIdx.x, n = blockIdx.z, a = blockIdx.y, b0 = blockIdx.x * 128;
  for (int i = tid; i < 2 * GYW * C4; i += 128) {
    const int f = i % C4, pc = i / C4, col = pc % GYW, row = pc / GYW;
    const int gr = a - 1 + row, gc = b0 - 1 + col;
    const bool in = (unsigned)gr < (unsigned)OH && (unsigned)gc < (unsigned)OW;
    cp_async16_zfill(sm + (row * GYW + col) * PS + 4 * f, in ? gy + ((size_t)(n * OH + gr) * OW + gc) * COUT + 4 * f : gy, in);
  }
  cp_async_wait_all();
  __syncthreads();
  float acc[2][2][3];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) acc[dy][dx][ci] = 0.f;
#pragma unroll
  for (int lrow = 0; lrow < 2; ++lrow) {
#pragma unroll
    for (int lc = 0; lc < 2; ++lc) {
      float g[COUT];
      const float* src = sm + (lrow * GYW + tid + lc) * PS;
#pragma unroll
      for (int f = 0; f < C4; ++f) {
        const float4 t = *reinterpret_cast<const float4*>(src + 4 * f);
        g[4 * f] = t.x; g[4 * f + 1] = t.y; g[4 * f + 2] = t.z; g[4 * f + 3] = t.w;
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int dy = ky & 1;                            // parity of the input row this tap reaches
        if ((dy - ky) / 2 + 1 != lrow) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int dx = kx & 1;
          if ((dx - kx) / 2 + 1 != lc) continue;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            float t = acc[dy][dx][ci];
#pragma unroll
            for (int co = 0; co < COUT; ++co) t = fmaf(g[co], c_c3w[((ky * 3 + kx) * 3 + ci) * COUT + co], t);
            acc[dy][dx][ci] = t;
          }
        }
      }
    }
  }
  const int b = b0 + tid;
  if (2 * b < W) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      float* out = gx + ((size_t)(n * H + 2 * a + dy) * W + 2 * b) * 3;     // 6 contiguous floats: pixels (2b, 2b+1)
      *reinterpret_cast<float2*>(out) = make_float2(acc[dy][0][0], acc[dy][0][1]);
      *reinterpret_cast<float2*>(out + 2) = make_float2(acc[dy][0][2], acc[dy][1][0]);
      *reinterpret_cast<float2*>(out + 4) = make_float2(acc[dy][1][1], acc[dy][1][2]);
    }
  }
}

// stride 1, TF-SAME (pad 1): gx[y][x][ci] = sum_{ky,kx,co} gy[y+1-ky][x+1-kx][co] . w[ky][kx][ci][co]  (VGG block1_conv1's
// input gradient, 64 -> 3 at 256x256: the one perceptual-loss layer whose gradient reaches the image).
// block = 128 pixels x R = 4 rows, 64 R threads: thread (t, r) owns the pixel PAIR (2t, 2t+1) of row r.  The output channels
// of gy go through shared memory in two halves (R + 2 rows x 130 columns x 32 channels, fetched with cp.async: every 16-byte
// piece of a half is in flight at once, rows are read 1.5 times from L2 instead of 3); a gy pixel is read once (8 LDS.128)
// and feeds both pixels of the pair: 96 LDS.128 per 3456 FFMAs, every FFMA with its weight as a constant-bank operand (the
// first-generation kernel: 4 LDS.128 per 12 FFMAs, 490 us at batch 16).  The staged columns are split by parity
// (even | odd) so that the lanes of a warp - which walk columns 2t + j - read consecutive records: pixel stride 36 floats =
// 4 banks, conflict-free 16-byte reads.
template <int COUT, int HALF>
__device__ __forceinline__ void c3k_dgrad_s1_body(const float* __restrict__ s, int t, float (&acc)[2][3]) {
  constexpr int HC = COUT / 2, PS = HC + 4, KS = 65;
#pragma unroll
  for (int lrow = 0; lrow < 3; ++lrow) {
    const int ky = 2 - lrow;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* src = s + ((size_t)((lrow * 2 + (j & 1)) * KS + t + (j >> 1))) * PS;
      float g[HC];
#pragma unroll
      for (int f = 0; f < HC / 4; ++f) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * f);
        g[4 * f] = v.x; g[4 * f + 1] = v.y; g[4 * f + 2] = v.z; g[4 * f + 3] = v.w;
      }
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int kx = d + 2 - j;
        if (kx < 0 || kx > 2) continue;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          float a = acc[d][ci];
#pragma unroll
          for (int co = 0; co < HC; ++co) a = fmaf(g[co], c_c3w[((ky * 3 + kx) * 3 + ci) * COUT + HALF * HC + co], a);
          acc[d][ci] = a;
        }
      }
    }
  }
}
template <int COUT, int R>
__global__ void __launch_bounds__(64 * R)
c3k_dgrad_s1_kernel(const float* __restrict__ gy, float* __restrict__ gx, int H, int W) {
  constexpr int HC = COUT / 2, H4 = HC / 4, PS = HC + 4, KS = 65, NCOL = 130, NT = 64 * R;
  extern __shared__ __align__(16) float sm[];            // [R + 2 rows y0-1 ..][parity][KS][PS]; columns x0-1 .. x0+128
  const int tid = threadIdx.x, n = blockIdx.z, y0 = blockIdx.y * R, x0 = blockIdx.x * 128;
  const int t = tid & 63, r = tid >> 6;
  float acc[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (half) __syncthreads();                            // every thread is done with the first half's rows
    for (int i = tid; i < (R + 2) * NCOL * H4; i += NT) {
      const int f = i % H4, pc = i / H4, col = pc % NCOL, row = pc / NCOL;
      const int gr = y0 - 1 + row, gc = x0 - 1 + col;
      const bool in = (unsigned)gr < (unsigned)H && (unsigned)gc < (unsigned)W;
      const float* src = in ? gy + ((size_t)(n * H + gr) * W + gc) * COUT + half * HC + 4 * f : gy;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + ((size_t)((row * 2 + (col & 1)) * KS + (col >> 1))) * PS + 4 * f);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");   // outside: zeros
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float* rows = sm + (size_t)r * 2 * KS * PS;     // this thread's output row y0 + r reads staged rows r .. r + 2
    if (half == 0) c3k_dgrad_s1_body<COUT, 0>(rows, t, acc);
    else c3k_dgrad_s1_body<COUT, 1>(rows, t, acc);
  }
  const int px = x0 + 2 * t, y = y0 + r;
  if (px < W && y < H) {
    float* out = gx + ((size_t)(n * H + y) * W + px) * 3;                  // 6 contiguous floats: pixels (px, px + 1)
    *reinterpret_cast<float2*>(out) = make_float2(acc[0][0], acc[0][1]);
    *reinterpret_cast<float2*>(out + 2) = make_float2(acc[0][2], acc[1][0]);
    *reinterpret_cast<float2*>(out + 4) = make_float2(acc[1][1], acc[1][2]);
  }
}

// ------------------------------------------------------------------------------------------------
// p3: Conv2D(3 -> 3, 1x1, stride 1) - the discriminators' / latent regressor's fromRGB layer
// (hologan_discriminator.py:20-26,75-81, initial_1x1_conv).  A pixel is 12 bytes in and 12 bytes out: the tensor is walked
// as a flat stream, 4 pixels = 3 float4 per thread per step, so every load / store is a full-width coalesced access and
// the kernel runs at the copy rate (the generic thread-per-pixel kernel spent its time in row decodes and 4-byte accesses:
// 40 us per launch for 50 MB, profiles/r01_conv_breakdown_final.txt).
//   forward : y[p][n] = b[n] + sum_c x[p][c] W[c][n]        (T = 0)
//   dgrad   : gx[p][c] = sum_n gy[p][n] W[c][n]             (T = 1: the same kernel with W transposed)
//   wgrad   : gW[c][n] = sum_p x[p][c] gy[p][n], gb[n] = sum_p gy[p][n]: per-block partials + ordered sum
// ------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(256)
p3_map_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
              size_t ngroups, size_t npix, int act, float alpha) {
  float W[3][3], B[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int n = 0; n < 3; ++n) W[c][n] = T ? __ldg(w + n * 3 + c) : __ldg(w + c * 3 + n);
#pragma unroll
  for (int n = 0; n < 3; ++n) B[n] = bias != nullptr ? __ldg(bias + n) : 0.f;
  for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ngroups; g += (size_t)gridDim.x * 256) {
    float v[12], o[12];
    if (4 * g + 4 <= npix) {
      const float4 a = ldg4(x + 12 * g), b = ldg4(x + 12 * g + 4), c = ldg4(x + 12 * g + 8);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) v[i] = (12 * g + i < 3 * npix) ? __ldg(x + 12 * g + i) : 0.f;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        float r = B[n];
#pragma unroll
        for (int c = 0; c < 3; ++c) r = fmaf(v[3 * p + c], W[c][n], r);
        o[3 * p + n] = cn_apply_act(r, act, alpha);
      }
    if (4 * g + 4 <= npix) {
      float4* dst = reinterpret_cast<float4*>(y + 12 * g);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]); dst[1] = make_float4(o[4], o[5], o[6], o[7]); dst[2] = make_float4(o[8], o[9], o[10], o[11]);
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) if (12 * g + i < 3 * npix) y[12 * g + i] = o[i];
    }
  }
}

// part[block][12]: 9 weight-gradient sums then 3 bias-gradient sums of the block's pixels (fixed order inside the block)
__global__ void __launch_bounds__(256)
p3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ part, size_t ngroups, size_t npix) {
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  for (size_t g = (size_t)blockIdx.x * 256 + threadIdx.x; g < ngroups; g += (size_t)gridDim.x * 256) {
    float v[12], u[12];
    if (4 * g + 4 <= npix) {
      const float4 a = ldg4(x + 12 * g), b = ldg4(x + 12 * g + 4), c = ldg4(x + 12 * g + 8);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
      const float4 d = ldg4(gy + 12 * g), e = ldg4(gy + 12 * g + 4), f = ldg4(gy + 12 * g + 8);
      u[0] = d.x; u[1] = d.y; u[2] = d.z; u[3] = d.w; u[4] = e.x; u[5] = e.y; u[6] = e.z; u[7] = e.w;
      u[8] = f.x; u[9] = f.y; u[10] = f.z; u[11] = f.w;
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const bool ok = 12 * g + i < 3 * npix;
        v[i] = ok ? __ldg(x + 12 * g + i) : 0.f; u[i] = ok ? __ldg(gy + 12 * g + i) : 0.f;
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int n = 0; n < 3; ++n) acc[c * 3 + n] = fmaf(v[3 * p + c], u[3 * p + n], acc[c * 3 + n]);
#pragma unroll
      for (int n = 0; n < 3; ++n) acc[9 + n] += u[3 * p + n];
    }
  }
  __shared__ float red[8][12];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float t = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) t += red[wq][threadIdx.x];
    part[(size_t)blockIdx.x * 12 + threadIdx.x] = t;
  }
}

// second-generation c3 kernels (cn_debug_set_c3k), bit 0 = forward, bit 1 = stride-2 input gradient.  Measured on B200
// (profiles/r02_skinny_c3k.txt, D.block0 at batch 32): dgrad 119.8 -> 89.1 us; forward 85.0 -> 85.0 us (both generations sit
// on the FP32 pipe: ptxas keeps the kernel out of the FFMA operand slot and loads it with LDC.128, and three-register
// FFMAs issue at half rate) - so only the gradient kernel is on by default.
int g_c3k = [] { const char* e = getenv("CN_C3K"); return e ? atoi(e) : 6; }();   // environment CN_C3K overrides for A/B runs

bool is_p3(const cn_conv_desc* d) {
  return d->nd == 2 && d->cin == 3 && d->cout == 3 && d->ksize[0] == 1 && d->ksize[1] == 1 && d->stride == 1 && d->upsample == 1 &&
         d->pad <= 0;
}

void same_pad(int in, int k, int s, int* out, int* pb) {
  *out = (in + s - 1) / s;
  int tot = (*out - 1) * s + k - in;
  if (tot < 0) tot = 0;
  *pb = tot / 2;
}

bool is_c3(const cn_conv_desc* d) {
  return d->nd == 2 && d->cin == 3 && d->ksize[0] == 3 && d->ksize[1] == 3 && d->upsample == 1 && d->pad < 0 &&
         (d->cout == 48 || d->cout == 64);
}
bool is_up4c3(const cn_conv_desc* d) {
  return d->nd == 2 && d->cin == 32 && d->cout == 3 && d->ksize[0] == 4 && d->ksize[1] == 4 && d->upsample == 2 &&
         d->stride == 1 && d->pad < 0;
}

template <typename K>
int opt_in_smem(K kernel, int bytes) {
  // once per (kernel, size): keeps the attribute call out of captured CUDA graphs.  Keyed by the kernel's ADDRESS: two
  // instantiations with the same signature (c3_wgrad_kernel<2, 48> / <1, 48>) share this function template's statics.
  static std::map<const void*, int> have;
  int& h = have[reinterpret_cast<const void*>(kernel)];
  if (bytes > 48 * 1024 && h < bytes) {
    CN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    h = bytes;
  }
  return CN_OK;
}

}  // namespace

// Each cn_skinny_* returns 1 when it launched the layer, 0 when the shape is not one of the skinny layers,
// or a negative CN_ERR_* code.
int cn_skinny_fwd(const cn_conv_desc* d, const float* x, const float* w, const float* bias, int act, float alpha,
                  float* y, cudaStream_t st) {
  const int H = d->in_dims[0], W = d->in_dims[1];
  if (is_p3(d)) {
    const size_t npix = (size_t)d->batch * H * W, ng = (npix + 3) / 4;
    int blocks = (int)((ng + 255) / 256); if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    p3_map_kernel<0><<<blocks, 256, 0, st>>>(x, w, bias, y, ng, npix, act, alpha);
    CN_CHECK_LAUNCH();
    return 1;
  }
  if (is_c3(d) && (g_c3k & 1)) {
    int OH, OW, pby, pbx;
    same_pad(H, 3, d->stride, &OH, &pby); same_pad(W, 3, d->stride, &OW, &pbx);
    CN_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_c3w, w, (size_t)27 * d->cout * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
    dim3 grid((OW + 127) / 128, OH, d->batch);
    const int ncol = 127 * d->stride + 3;
    const int smem = (((3 * ncol * 3 + 3) & ~3) + 128 * (d->cout + 4)) * (int)sizeof(float);
    if (d->stride == 2 && d->cout == 48) c3k_fwd_kernel<2, 48><<<grid, 128, smem, st>>>(x, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
    else if (d->stride == 2) c3k_fwd_kernel<2, 64><<<grid, 128, smem, st>>>(x, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
    else if (d->cout == 48) c3k_fwd_kernel<1, 48><<<grid, 128, smem, st>>>(x, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
    else c3k_fwd_kernel<1, 64><<<grid, 128, smem, st>>>(x, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
    CN_CHECK_LAUNCH();
    return 1;
  }
  if (is_c3(d)) {
    int OH, OW, pby, pbx;
    same_pad(H, 3, d->stride, &OH, &pby); same_pad(W, 3, d->stride, &OW, &pbx);
    dim3 grid((OW + 127) / 128, OH, d->batch);
    static const int pxt = [] { const char* e = getenv("CN_C3_PXT"); return e ? atoi(e) : 4; }();     // pixels per thread (A/B: 2)
    if (pxt == 2) {
      if (d->stride == 2 && d->cout == 48) c3_fwd_kernel<2, 48, 2><<<grid, 256, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
      else if (d->stride == 2) c3_fwd_kernel<2, 64, 2><<<grid, 256, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
      else if (d->cout == 48) c3_fwd_kernel<1, 48, 2><<<grid, 256, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
      else c3_fwd_kernel<1, 64, 2><<<grid, 256, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
    } else {
      if (d->stride == 2 && d->cout == 48) c3_fwd_kernel<2, 48, 4><<<grid, 128, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
      else if (d->stride == 2) c3_fwd_kernel<2, 64, 4><<<grid, 128, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
      else if (d->cout == 48) c3_fwd_kernel<1, 48, 4><<<grid, 128, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
      else c3_fwd_kernel<1, 64, 4><<<grid, 128, 0, st>>>(x, w, bias, y, H, W, OH, OW, pby, pbx, act, alpha);
    }
    CN_CHECK_LAUNCH();
    return 1;
  }
  if (is_up4c3(d)) {
    const int smem = (25 * 32 + 3 * 130 * 9) * 16;
    int rc = opt_in_smem(up4c3_fwd_kernel<32>, smem); if (rc) return rc;
    dim3 grid((W + 127) / 128, H, d->batch);
    up4c3_fwd_kernel<32><<<grid, 256, smem, st>>>(x, w, bias, y, H, W, act, alpha);
    CN_CHECK_LAUNCH();
    return 1;
  }
  return 0;
}

int cn_skinny_dgrad(const cn_conv_desc* d, const float* gy, const float* w, float* gx, cudaStream_t st) {
  const int H = d->in_dims[0], W = d->in_dims[1];
  if (is_p3(d)) {
    const size_t npix = (size_t)d->batch * H * W, ng = (npix + 3) / 4;
    int blocks = (int)((ng + 255) / 256); if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    p3_map_kernel<1><<<blocks, 256, 0, st>>>(gy, w, nullptr, gx, ng, npix, CN_ACT_NONE, 0.f);
    CN_CHECK_LAUNCH();
    return 1;
  }
  if (is_c3(d)) {
    int OH, OW, pby, pbx;
    same_pad(H, 3, d->stride, &OH, &pby); same_pad(W, 3, d->stride, &OW, &pbx);
    if ((g_c3k & 2) && d->stride == 2 && d->cout == 48 && H % 2 == 0 && W % 2 == 0) {
      const int smem = 2 * 129 * (48 + 4) * (int)sizeof(float);
      int rc = opt_in_smem(c3k_dgrad_s2_kernel<48>, smem); if (rc) return rc;
      CN_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_c3w, w, (size_t)27 * 48 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
      dim3 grid((W / 2 + 127) / 128, H / 2, d->batch);
      c3k_dgrad_s2_kernel<48><<<grid, 128, smem, st>>>(gy, gx, H, W, OH, OW);
      CN_CHECK_LAUNCH();
      return 1;
    }
    if (d->stride == 2 && d->cout == 48 && H % 2 == 0 && W % 2 == 0) {
      constexpr int PS = 48;
      const int smem = (27 * 48 + 2 * 65 * PS) * 4;
      int rc = opt_in_smem(c3_dgrad_kernel<2, 48, PS>, smem); if (rc) return rc;
      dim3 grid((W + 127) / 128, H / 2, d->batch);
      c3_dgrad_kernel<2, 48, PS><<<grid, 256, smem, st>>>(gy, w, gx, H, W, OH, OW);
      CN_CHECK_LAUNCH();
      return 1;
    }
    if ((g_c3k & 4) && d->stride == 1 && d->cout == 64 && W % 2 == 0) {
      constexpr int R = 4;
      const int smem = (R + 2) * 2 * 65 * (32 + 4) * (int)sizeof(float);
      int rc = opt_in_smem(c3k_dgrad_s1_kernel<64, R>, smem); if (rc) return rc;
      CN_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_c3w, w, (size_t)27 * 64 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
      dim3 grid((W + 127) / 128, (H + R - 1) / R, d->batch);
      c3k_dgrad_s1_kernel<64, R><<<grid, 64 * R, smem, st>>>(gy, gx, H, W);
      CN_CHECK_LAUNCH();
      return 1;
    }
    if (d->stride == 1 && d->cout == 64) {
      constexpr int PS = 72;
      const int smem = (27 * 64 + 3 * 130 * PS) * 4;
      int rc = opt_in_smem(c3_dgrad_kernel<1, 64, PS>, smem); if (rc) return rc;
      dim3 grid((W + 127) / 128, H, d->batch);
      c3_dgrad_kernel<1, 64, PS><<<grid, 256, smem, st>>>(gy, w, gx, H, W, OH, OW);
      CN_CHECK_LAUNCH();
      return 1;
    }
    return 0;
  }
  if (is_up4c3(d)) {
    const int smem = (25 * 3 * 32 + 5 * (2 * 64 + 3) * 3) * 4;
    dim3 grid((W + 63) / 64, H, d->batch);
    up4c3_dgrad_kernel<32><<<grid, 256, smem, st>>>(gy, w, gx, H, W);
    CN_CHECK_LAUNCH();
    return 1;
  }
  return 0;
}

size_t cn_skinny_wgrad_scratch(const cn_conv_desc* d) {
  if (is_p3(d)) return (size_t)(4 * sm_count() + 1) * 12 * sizeof(float);
  if (is_c3(d)) return (size_t)2 * sm_count() * 27 * d->cout * sizeof(float);
  if (is_up4c3(d)) return (size_t)2 * sm_count() * 25 * 32 * 3 * sizeof(float);
  return 0;
}

// p3 (1x1, 3 -> 3) also produces the bias gradient when gbias != nullptr and then returns 2
__global__ void p3_finish_kernel(const float* __restrict__ sum12, float* __restrict__ gw, float* __restrict__ gbias) {
  const int i = threadIdx.x;
  if (i < 9) gw[i] = sum12[i];
  else if (i < 12 && gbias != nullptr) gbias[i - 9] = sum12[i];
}
int cn_skinny_wgrad(const cn_conv_desc* d, const float* x, const float* gy, float* gw, float* gbias, float* scratch, cudaStream_t st) {
  const int H = d->in_dims[0], W = d->in_dims[1];
  if (is_p3(d)) {
    const size_t npix = (size_t)d->batch * H * W, ng = (npix + 3) / 4;
    int blocks = (int)((ng + 255) / 256); if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
    p3_wgrad_kernel<<<blocks, 256, 0, st>>>(x, gy, scratch, ng, npix);
    CN_CHECK_LAUNCH();
    float* sum12 = scratch + (size_t)4 * sm_count() * 12;
    sum_partials_kernel<<<1, dim3(32, 8), 0, st>>>(scratch, blocks, 12, sum12);
    CN_CHECK_LAUNCH();
    p3_finish_kernel<<<1, 32, 0, st>>>(sum12, gw, gbias);
    CN_CHECK_LAUNCH();
    return gbias != nullptr ? 2 : 1;
  }
  if (is_c3(d) && d->cout == 48) {
    int OH, OW, pby, pbx;
    same_pad(H, 3, d->stride, &OH, &pby); same_pad(W, 3, d->stride, &OW, &pbx);
    int rows = d->batch * OH * ((OW + 127) / 128);
    int blocks = 2 * sm_count(); if (blocks > rows) blocks = rows;
    if (d->stride == 2) {
      const int smem = (2 * 128 * 48 + 2 * ((3 * (127 * 2 + 3) * 3 + 8 + 3) & ~3)) * (int)sizeof(float);
      int rc = opt_in_smem(c3_wgrad_kernel<2, 48>, smem); if (rc) return rc;
      c3_wgrad_kernel<2, 48><<<blocks, 256, smem, st>>>(x, gy, scratch, d->batch, H, W, OH, OW, pby, pbx);
    } else {
      const int smem = (2 * 128 * 48 + 2 * ((3 * (127 * 1 + 3) * 3 + 8 + 3) & ~3)) * (int)sizeof(float);
      int rc = opt_in_smem(c3_wgrad_kernel<1, 48>, smem); if (rc) return rc;
      c3_wgrad_kernel<1, 48><<<blocks, 256, smem, st>>>(x, gy, scratch, d->batch, H, W, OH, OW, pby, pbx);
    }
    CN_CHECK_LAUNCH();
    sum_partials_kernel<<<(27 * 48 + 31) / 32, dim3(32, 8), 0, st>>>(scratch, blocks, 27 * 48, gw);
    CN_CHECK_LAUNCH();
    return 1;
  }
  if (is_up4c3(d)) {
    int rows = d->batch * H * ((W + 127) / 128);
    int blocks = 2 * sm_count(); if (blocks > rows) blocks = rows;
    up4c3_wgrad_kernel<32><<<blocks, 256, 0, st>>>(x, gy, scratch, d->batch, H, W);
    CN_CHECK_LAUNCH();
    up4c3_unfold_kernel<32><<<(16 * 32 * 3 + 31) / 32, dim3(32, 8), 0, st>>>(scratch, blocks, gw);
    CN_CHECK_LAUNCH();
    return 1;
  }
  return 0;
}

#ifdef CN_TEST_HOOKS
extern "C" int cn_debug_set_c3k(int v) { g_c3k = v; return CN_OK; }
#endif
