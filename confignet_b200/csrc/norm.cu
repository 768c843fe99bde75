// Per-(sample, channel) statistics and broadcast-affine kernels over channels-last tensors, plus the
// closed-form coefficient kernels of the three normalisations on ConfigNet's hot path and their
// first- and second-order gradients (the R1 penalty differentiates the discriminator's backward
// pass, losses.py:42-43,75-82).
//
//   AdaIN          building_blocks.py:135-149   (x-mu)*rsqrt(var+1e-3)*(1+s)+b
//   InstanceNorm   instance_normalization.py:108-131   gamma*(x-mu)/(std+1e-3)+beta  (eps on the STD)
//   layer style    confignet_utils.py:147-159   concat(mean, sqrt(var+1e-6))
//
// Every op is "7 sums per (n,c)" -> "a few scalars per (n,c)" -> "out = ka*a + kb*b + kc*c + k0".
// HBM-bound: one read of each operand per pass, float4 along the channel axis, warp-shuffle-free
// column reductions (threads own channels, so no cross-lane traffic until the 8-row smem fold).
#include "common.cuh"
#include <stdlib.h>

#define CN_SUMS_LD 8        // floats per (slice, sample, channel) record of the statistics buffer: 7 sums + 1 pad
#define CN_FLAG_LRELU_A 1   // a := lrelu(a, alpha) before use
#define CN_FLAG_MASK_OUT 2  // affine result *= lrelu'(a_raw)
#define CN_FLAG_MASK_C 4    // c := c * lrelu'(a_raw)

__device__ __forceinline__ float lrelu_f(float v, float alpha) { return v > 0.f ? v : v * alpha; }
__device__ __forceinline__ float lrelu_d(float v, float alpha) { return v > 0.f ? 1.f : alpha; }

// grid: (ceil(ch/32), n, psplit)  block: (32, 8).  Every block writes its own slice sums[z][n][col][0..6]
// (no atomics, no zero-fill); the coefficient kernel adds the psplit slices in a fixed order, so the
// statistics - and with them generate_images / encode_images - are bit-reproducible from run to run.
template <int NT>
__global__ void chan_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                 int p, int ch, int flags, float alpha, float* __restrict__ sums) {
  __shared__ float sm[7][8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int n = blockIdx.y;
  float s[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < ch) {
    const int per = (p + gridDim.z - 1) / gridDim.z;
    const int pbeg = blockIdx.z * per, pend = min(p, pbeg + per);
    const size_t base = (size_t)n * p * ch + col;
    for (int r = pbeg + threadIdx.y; r < pend; r += 8) {
      const size_t i = base + (size_t)r * ch;
      float araw = a[i];
      float av = (flags & CN_FLAG_LRELU_A) ? lrelu_f(araw, alpha) : araw;
      s[0] += av; s[3] += av * av;
      if (NT >= 2) { float bv = b[i]; s[1] += bv; s[4] += av * bv;
        if (NT >= 3) { float cv = c[i]; if (flags & CN_FLAG_MASK_C) cv *= lrelu_d(araw, alpha);
          s[2] += cv; s[5] += av * cv; s[6] += bv * cv; } }
    }
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) sm[j][threadIdx.y][threadIdx.x] = s[j];
  __syncthreads();
  if (threadIdx.y < 7 && col < ch) {
    const int j = threadIdx.y;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[j][i][threadIdx.x];
    sums[(((size_t)blockIdx.z * gridDim.y + n) * ch + col) * CN_SUMS_LD + j] = t;
  }
}

// float4 form (ch % 4 == 0, ch <= 1024): block = one (sample, pixel slice); thread (q, rr) owns channel quad q and walks
// the pixels rr, rr + R, ... - consecutive threads read consecutive 16-byte pieces of the tensor (a warp instruction is
// one contiguous 512-byte run whatever the channel count: 48 and 96 channels no longer leave lanes idle), four
// independent loads in flight per thread.  grid (n, psplit), 256 threads, smem R*ch*7 floats (<= 28 KB).
template <int NT>
__global__ void __launch_bounds__(256)
chan_sums4_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                  int p, int ch, int flags, float alpha, float* __restrict__ sums) {
  extern __shared__ __align__(16) float red[];   // [R][chq][7][4]
  const int chq = ch >> 2, R = 256 / chq;
  const int tid = threadIdx.x, q = tid % chq, rr = tid / chq;
  const int n = blockIdx.x;
  const int per = (p + gridDim.y - 1) / gridDim.y;
  const int pbeg = blockIdx.y * per, pend = min(p, pbeg + per);
  float s[7][4];
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
  if (rr < R) {
    const size_t base = (size_t)n * p * ch + 4 * q;
    auto add = [&](const float4& a4, const float4& b4, const float4& c4) {
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float av = (flags & CN_FLAG_LRELU_A) ? lrelu_f(ar[e], alpha) : ar[e];
        s[0][e] += av; s[3][e] = fmaf(av, av, s[3][e]);       // explicit FMAs: the dual kernel below must round identically
        if (NT >= 2) { s[1][e] += bv[e]; s[4][e] = fmaf(av, bv[e], s[4][e]); }
        if (NT >= 3) {
          float cc = cv[e];
          if (flags & CN_FLAG_MASK_C) cc *= lrelu_d(ar[e], alpha);
          s[2][e] += cc; s[5][e] = fmaf(av, cc, s[5][e]); s[6][e] = fmaf(bv[e], cc, s[6][e]);
        }
      }
    };
    // U pixels of every operand are loaded before the first is used: written as "#pragma unroll" of a load-use loop the
    // compiler keeps ONE load in flight per thread (each use waits for its load) and the pass runs at 0.3-0.5 of the HBM rate
    constexpr int U = (NT == 1) ? 8 : 4;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int r = pbeg + rr;
    for (; r + (U - 1) * R < pend; r += U * R) {          // full batches: unconditional loads, all issued before the first use
      float4 A[U], B[U], C[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = base + (size_t)(r + u * R) * ch;
        A[u] = cn_ldg4_ordered(a + i);
        B[u] = (NT >= 2) ? cn_ldg4_ordered(b + i) : z4;
        C[u] = (NT >= 3) ? cn_ldg4_ordered(c + i) : z4;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) add(A[u], B[u], C[u]);
    }
    for (; r < pend; r += R) {
      const size_t i = base + (size_t)r * ch;
      add(cn_ldg4_ordered(a + i), (NT >= 2) ? cn_ldg4_ordered(b + i) : z4, (NT >= 3) ? cn_ldg4_ordered(c + i) : z4);
    }
#pragma unroll
    for (int j = 0; j < 7; ++j)
      *reinterpret_cast<float4*>(red + ((size_t)(rr * chq + q) * 7 + j) * 4) = make_float4(s[j][0], s[j][1], s[j][2], s[j][3]);
  }
  __syncthreads();
  for (int o = tid; o < ch * 7; o += 256) {
    const int cc = o / 7, j = o - cc * 7, qq = cc >> 2, e = cc & 3;
    float t = 0.f;
    for (int k = 0; k < R; ++k) t += red[((size_t)(k * chq + qq) * 7 + j) * 4 + e];
    sums[(((size_t)blockIdx.y * gridDim.x + n) * ch + cc) * CN_SUMS_LD + j] = t;
  }
}

// One pass over a for BOTH statistics a DiscrBlock needs of its conv output (building_blocks.py:100-106): the raw sums
// (get_layer_style) into sums_raw and the sums of lrelu(a) (InstanceNormalization after LeakyReLU) into sums_act, records
// laid out as by chan_sums4_kernel<1> (j = 0: sum, j = 3: sum of squares).  Same thread mapping, eight pixels in flight.
__global__ void __launch_bounds__(256)
chan_sums4_dual_kernel(const float* __restrict__ a, int p, int ch, float alpha, float* __restrict__ sums_act,
                       float* __restrict__ sums_raw) {
  extern __shared__ __align__(16) float red[];   // [R][chq][4][4]: act sum, act squares, raw sum, raw squares
  const int chq = ch >> 2, R = 256 / chq;
  const int tid = threadIdx.x, q = tid % chq, rr = tid / chq;
  const int n = blockIdx.x;
  const int per = (p + gridDim.y - 1) / gridDim.y;
  const int pbeg = blockIdx.y * per, pend = min(p, pbeg + per);
  float s[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
  if (rr < R) {
    const size_t base = (size_t)n * p * ch + 4 * q;
    auto add = [&](const float4& a4) {
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float av = lrelu_f(ar[e], alpha);
        s[0][e] += av; s[1][e] = fmaf(av, av, s[1][e]);
        s[2][e] += ar[e]; s[3][e] = fmaf(ar[e], ar[e], s[3][e]);
      }
    };
    constexpr int U = 8;
    int r = pbeg + rr;
    for (; r + (U - 1) * R < pend; r += U * R) {
      float4 A[U];
#pragma unroll
      for (int u = 0; u < U; ++u) A[u] = cn_ldg4_ordered(a + base + (size_t)(r + u * R) * ch);
#pragma unroll
      for (int u = 0; u < U; ++u) add(A[u]);
    }
    for (; r < pend; r += R) add(cn_ldg4_ordered(a + base + (size_t)r * ch));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(red + ((size_t)(rr * chq + q) * 4 + j) * 4) = make_float4(s[j][0], s[j][1], s[j][2], s[j][3]);
  }
  __syncthreads();
  for (int o = tid; o < ch * 4; o += 256) {
    const int cc = o >> 2, j = o & 3, qq = cc >> 2, e = cc & 3;
    float t = 0.f;
    for (int k = 0; k < R; ++k) t += red[((size_t)(k * chq + qq) * 4 + j) * 4 + e];
    float* dst = (j < 2 ? sums_act : sums_raw) + (((size_t)blockIdx.y * gridDim.x + n) * ch + cc) * CN_SUMS_LD;
    dst[(j & 1) ? 3 : 0] = t;
  }
}

extern "C" int cn_chan_sums_splits(int n, int p, int ch) {
  if (n <= 0 || p <= 0 || ch <= 0) return 1;
  int cb = (ch + 31) / 32;
  if (ch % 4 == 0 && ch <= 1024) cb = 1;           // float4 kernel: one block per (sample, slice), ~2 blocks per SM
  // ~4 blocks of 256 threads per SM: a one-operand pass needs that many 16-byte loads in flight to approach the HBM rate
  // (measured with 2 per SM: 2.1 TB/s on one operand, 4.9 TB/s on two - profiles/r02_launches_v11_summary.txt)
  static int per_sm = 0;
  if (per_sm == 0) { const char* e = getenv("CN_SUMS_BLOCKS_PER_SM"); per_sm = e ? atoi(e) : 4; if (per_sm < 1) per_sm = 4; }
  int psplit = (per_sm * 148 + cb * n - 1) / (cb * n);
  // at least 8 pixels (float4 kernel: 4 per row group) per block: a thread that walks 64 pixels alone is 16 dependent
  // DRAM round trips - 20 us on the 4 MB tensors of the last discriminator blocks, whatever the grid
  int minpix = 64;
  if (ch % 4 == 0 && ch <= 1024) { minpix = 4 * (256 / (ch / 4)); if (minpix < 8) minpix = 8; if (minpix > 64) minpix = 64; }
  int maxsplit = (p + minpix - 1) / minpix;
  if (psplit > maxsplit) psplit = maxsplit;
  if (psplit < 1) psplit = 1;
  return psplit;
}

extern "C" int cn_chan_sums(const float* a, const float* b, const float* c, int n, int p, int ch,
                            int flags, float alpha, float* sums, void* stream) {
  CN_REQUIRE(a && sums && n > 0 && p > 0 && ch > 0, CN_ERR_BAD_SHAPE, "cn_chan_sums: bad arguments");
  CN_REQUIRE(!(c && !b), CN_ERR_BAD_SHAPE, "cn_chan_sums: c given without b");
  cudaStream_t st = (cudaStream_t)stream;
  int cb = (ch + 31) / 32;
  int psplit = cn_chan_sums_splits(n, p, ch);
  if (ch % 4 == 0 && ch <= 1024) {
    const int chq = ch / 4, R = 256 / chq;
    const int smem = R * ch * 7 * (int)sizeof(float);
    dim3 grid4(n, psplit);
    if (c) chan_sums4_kernel<3><<<grid4, 256, smem, st>>>(a, b, c, p, ch, flags, alpha, sums);
    else if (b) chan_sums4_kernel<2><<<grid4, 256, smem, st>>>(a, b, c, p, ch, flags, alpha, sums);
    else chan_sums4_kernel<1><<<grid4, 256, smem, st>>>(a, b, c, p, ch, flags, alpha, sums);
    CN_CHECK_LAUNCH();
    return CN_OK;
  }
  dim3 grid(cb, n, psplit), block(32, 8);
  if (c) chan_sums_kernel<3><<<grid, block, 0, st>>>(a, b, c, p, ch, flags, alpha, sums);
  else if (b) chan_sums_kernel<2><<<grid, block, 0, st>>>(a, b, c, p, ch, flags, alpha, sums);
  else chan_sums_kernel<1><<<grid, block, 0, st>>>(a, b, c, p, ch, flags, alpha, sums);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// sums of lrelu(a, alpha) into sums_act and of a itself into sums_raw in one pass (records as cn_chan_sums writes them for
// one operand; the j = 1, 2, 4, 5, 6 entries are left untouched - the one-operand closed forms do not read them).
// CN_ERR_UNSUPPORTED when ch % 4 != 0 or ch > 1024: the caller then runs two cn_chan_sums passes.
extern "C" int cn_chan_sums_dual(const float* a, int n, int p, int ch, float alpha, float* sums_act, float* sums_raw, void* stream) {
  CN_REQUIRE(a && sums_act && sums_raw && n > 0 && p > 0 && ch > 0, CN_ERR_BAD_SHAPE, "cn_chan_sums_dual: bad arguments");
  CN_REQUIRE(ch % 4 == 0 && ch <= 1024, CN_ERR_UNSUPPORTED, "cn_chan_sums_dual: channel count needs the two-pass form");
  const int psplit = cn_chan_sums_splits(n, p, ch);
  const int chq = ch / 4, R = 256 / chq;
  const int smem = R * ch * 4 * (int)sizeof(float);
  chan_sums4_dual_kernel<<<dim3(n, psplit), 256, smem, (cudaStream_t)stream>>>(a, p, ch, alpha, sums_act, sums_raw);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// out[n,r,ch] = (ka*a + kb*b + kc*c + k0) [* lrelu'(a_raw)], coef (n, ch, 4)
template <int VEC>
__global__ void chan_affine_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                   const float4* __restrict__ coef, int p, int ch, int flags, float alpha,
                                   float* __restrict__ out, size_t total_vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int chv = ch / VEC;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total_vec; v += stride) {
    const int cv = (int)(v % chv);
    const size_t n = v / ((size_t)chv * p);
    const size_t i = v * VEC;
    float av[VEC], bv[VEC], cvv[VEC], ov[VEC];
    if (VEC == 4) {
      float4 t = *reinterpret_cast<const float4*>(a + i); av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
      if (b) { t = *reinterpret_cast<const float4*>(b + i); bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w; }
      if (c) { t = *reinterpret_cast<const float4*>(c + i); cvv[0] = t.x; cvv[1] = t.y; cvv[2] = t.z; cvv[3] = t.w; }
    } else {
      av[0] = a[i]; if (b) bv[0] = b[i]; if (c) cvv[0] = c[i];
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const float4 k = coef[n * ch + cv * VEC + q];
      const float araw = av[q];
      const float aa = (flags & CN_FLAG_LRELU_A) ? lrelu_f(araw, alpha) : araw;
      float r = fmaf(k.x, aa, k.w);
      if (b) r = fmaf(k.y, bv[q], r);
      if (c) { float cc = cvv[q]; if (flags & CN_FLAG_MASK_C) cc *= lrelu_d(araw, alpha); r = fmaf(k.z, cc, r); }
      if (flags & CN_FLAG_MASK_OUT) r *= lrelu_d(araw, alpha);
      ov[q] = r;
    }
    if (VEC == 4) *reinterpret_cast<float4*>(out + i) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    else out[i] = ov[0];
  }
}

// Row-walking form (ch % 4 == 0, ch <= 1024): block = one (sample, pixel slice), thread (q, rr) owns channel quad q and
// walks the pixels rr, rr + R, ... of the slice.  Its four coefficient records are loaded ONCE into registers (the
// grid-stride kernel above re-reads 64 bytes of coefficients and does two 64-bit divisions per 16 bytes of data: it ran at
// 0.2-0.45 of the HBM rate, profiles/r02_launches_v11_summary.txt), a warp instruction is one contiguous 512-byte run,
// four pixels are in flight per thread.  grid (n, psplit), 256 threads.
template <bool HB, bool HC>
__global__ void __launch_bounds__(256)
chan_affine_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                        const float4* __restrict__ coef, const float4* __restrict__ coef2, int p, int ch, int flags, float alpha,
                        float* __restrict__ out) {
  const int chq = ch >> 2, R = 256 / chq;
  const int tid = threadIdx.x, q = tid % chq, rr = tid / chq;
  if (rr >= R) return;
  const int n = blockIdx.x;
  const int per = (p + gridDim.y - 1) / gridDim.y;
  const int pbeg = blockIdx.y * per, pend = min(p, pbeg + per);
  float4 k[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) k[e] = coef[(size_t)n * ch + 4 * q + e];
  // second coefficient set (cn_chan_affine2): + k2.x * a_raw + k2.w after the mask - the layer-style gradient riding on
  // the InstanceNorm gradient's pass over the same tensor
  float k2x[4] = {0.f, 0.f, 0.f, 0.f}, k2w[4] = {0.f, 0.f, 0.f, 0.f};
  const bool two = coef2 != nullptr;
  if (two) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float4 t = coef2[(size_t)n * ch + 4 * q + e]; k2x[e] = t.x; k2w[e] = t.w; }
  }
  const size_t base = (size_t)n * p * ch + 4 * q;
  constexpr int U = 4;                                   // four pixels of every operand in flight per thread
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto one = [&](const float4& A, const float4& B, const float4& C, float* o4) {
    const float ar[4] = {A.x, A.y, A.z, A.w}, br[4] = {B.x, B.y, B.z, B.w}, cr[4] = {C.x, C.y, C.z, C.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float araw = ar[e];
      const float aa = (flags & CN_FLAG_LRELU_A) ? lrelu_f(araw, alpha) : araw;
      float v = fmaf(k[e].x, aa, k[e].w);
      if (HB) v = fmaf(k[e].y, br[e], v);
      if (HC) { float cc = cr[e]; if (flags & CN_FLAG_MASK_C) cc *= lrelu_d(araw, alpha); v = fmaf(k[e].z, cc, v); }
      if (flags & CN_FLAG_MASK_OUT) v *= lrelu_d(araw, alpha);
      if (two) v += fmaf(k2x[e], araw, k2w[e]);
      o[e] = v;
    }
    *reinterpret_cast<float4*>(o4) = make_float4(o[0], o[1], o[2], o[3]);
  };
  int r = pbeg + rr;
  for (; r + (U - 1) * R < pend; r += U * R) {           // full batches, no branch between the loads and the stores
    float4 A[U], B[U], C[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + (size_t)(r + u * R) * ch;
      A[u] = cn_ldg4_ordered(a + i);
      B[u] = HB ? cn_ldg4_ordered(b + i) : z4;
      C[u] = HC ? cn_ldg4_ordered(c + i) : z4;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(A[u], B[u], C[u], out + base + (size_t)(r + u * R) * ch);
  }
  for (; r < pend; r += R) {
    const size_t i = base + (size_t)r * ch;
    one(cn_ldg4_ordered(a + i), HB ? cn_ldg4_ordered(b + i) : z4, HC ? cn_ldg4_ordered(c + i) : z4, out + i);
  }
}

static bool affine_rows_ok(int n, int ch) {
  static int rows_form = -1;
  if (rows_form < 0) { const char* e = getenv("CN_AFFINE_ROWS"); rows_form = e ? atoi(e) : 1; }
  return rows_form && ch % 4 == 0 && ch <= 1024 && n <= 65535;
}
static void launch_affine_rows(const float* a, const float* b, const float* c, const float* coef, const float* coef2,
                               int n, int p, int ch, int flags, float alpha, float* out, cudaStream_t st) {
  const int R = 256 / (ch / 4);
  int psplit = (8 * 148 + n - 1) / n;                    // ~8 blocks of 256 threads per SM
  int maxsplit = p / (2 * R); if (maxsplit < 1) maxsplit = 1;
  if (psplit > maxsplit) psplit = maxsplit;
  dim3 grid(n, psplit);
  const float4 *k4 = (const float4*)coef, *k2 = (const float4*)coef2;
  if (b && c) chan_affine_rows_kernel<true, true><<<grid, 256, 0, st>>>(a, b, c, k4, k2, p, ch, flags, alpha, out);
  else if (b) chan_affine_rows_kernel<true, false><<<grid, 256, 0, st>>>(a, b, c, k4, k2, p, ch, flags, alpha, out);
  else if (c) chan_affine_rows_kernel<false, true><<<grid, 256, 0, st>>>(a, b, c, k4, k2, p, ch, flags, alpha, out);
  else chan_affine_rows_kernel<false, false><<<grid, 256, 0, st>>>(a, b, c, k4, k2, p, ch, flags, alpha, out);
}

extern "C" int cn_chan_affine(const float* a, const float* b, const float* c, const float* coef,
                              int n, int p, int ch, int flags, float alpha, float* out, void* stream) {
  CN_REQUIRE(a && coef && out && n > 0 && p > 0 && ch > 0, CN_ERR_BAD_SHAPE, "cn_chan_affine: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)n * p * ch;
  if (affine_rows_ok(n, ch)) {
    launch_affine_rows(a, b, c, coef, nullptr, n, p, ch, flags, alpha, out, st);
    CN_CHECK_LAUNCH();
    return CN_OK;
  }
  if (ch % 4 == 0) {
    size_t tv = total / 4;
    int blocks = (int)((tv + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    chan_affine_kernel<4><<<blocks, 256, 0, st>>>(a, b, c, (const float4*)coef, p, ch, flags, alpha, out, tv);
  } else {
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    chan_affine_kernel<1><<<blocks, 256, 0, st>>>(a, b, c, (const float4*)coef, p, ch, flags, alpha, out, total);
  }
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// out = [cn_chan_affine(a, b, c, coef, flags)] + coef2.x * a_raw + coef2.w: two broadcast-affine results over the same
// tensor in ONE pass (the InstanceNorm and layer-style gradients of a DiscrBlock's conv output, building_blocks.py:100-106).
// Channel counts the row-walking kernel does not take (ch % 4 != 0 or ch > 1024) return CN_ERR_UNSUPPORTED: the caller
// then issues the two passes and adds them.
extern "C" int cn_chan_affine2(const float* a, const float* b, const float* c, const float* coef, const float* coef2,
                               int n, int p, int ch, int flags, float alpha, float* out, void* stream) {
  CN_REQUIRE(a && coef && coef2 && out && n > 0 && p > 0 && ch > 0, CN_ERR_BAD_SHAPE, "cn_chan_affine2: bad arguments");
  CN_REQUIRE(affine_rows_ok(n, ch), CN_ERR_UNSUPPORTED, "cn_chan_affine2: channel count needs the two-pass form");
  launch_affine_rows(a, b, c, coef, coef2, n, p, ch, flags, alpha, out, (cudaStream_t)stream);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// Two broadcast-affine results of the SAME three operands in one pass: the second-order terms of InstanceNorm(LeakyReLU(a))
// (R1 penalty, losses.py:75-82) need, from (a, gy, h),
//   out_a = (ka.x lrelu(a) + ka.y gy + ka.z h' + ka.w) * lrelu'(a)      (gradient wrt the conv output)
//   out_g =  kg.x lrelu(a)            + kg.z h' + kg.w                   (gradient wrt the incoming gradient)
// with h' = h * lrelu'(a): 3 reads + 2 writes instead of two passes (4 + 3 tensor moves).  Same arithmetic as two
// cn_chan_affine calls with flags LRELU_A | MASK_C (| MASK_OUT for out_a).
__global__ void __launch_bounds__(256)
chan_affine_pair_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                        const float4* __restrict__ coef_a, const float4* __restrict__ coef_g, int p, int ch, float alpha,
                        float* __restrict__ out_a, float* __restrict__ out_g) {
  const int chq = ch >> 2, R = 256 / chq;
  const int tid = threadIdx.x, q = tid % chq, rr = tid / chq;
  if (rr >= R) return;
  const int n = blockIdx.x;
  const int per = (p + gridDim.y - 1) / gridDim.y;
  const int pbeg = blockIdx.y * per, pend = min(p, pbeg + per);
  float4 ka[4], kg[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) { ka[e] = coef_a[(size_t)n * ch + 4 * q + e]; kg[e] = coef_g[(size_t)n * ch + 4 * q + e]; }
  const size_t base = (size_t)n * p * ch + 4 * q;
  auto one = [&](const float4& A, const float4& B, const float4& C, size_t i) {
    const float ar[4] = {A.x, A.y, A.z, A.w}, br[4] = {B.x, B.y, B.z, B.w}, cr[4] = {C.x, C.y, C.z, C.w};
    float oa[4], og[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float araw = ar[e], aa = lrelu_f(araw, alpha), d = lrelu_d(araw, alpha);
      const float cc = cr[e] * d;
      float v = fmaf(ka[e].x, aa, ka[e].w);
      v = fmaf(ka[e].y, br[e], v);
      v = fmaf(ka[e].z, cc, v);
      oa[e] = v * d;
      float g = fmaf(kg[e].x, aa, kg[e].w);
      og[e] = fmaf(kg[e].z, cc, g);
    }
    *reinterpret_cast<float4*>(out_a + i) = make_float4(oa[0], oa[1], oa[2], oa[3]);
    *reinterpret_cast<float4*>(out_g + i) = make_float4(og[0], og[1], og[2], og[3]);
  };
  constexpr int U = 4;
  int r = pbeg + rr;
  for (; r + (U - 1) * R < pend; r += U * R) {
    float4 A[U], B[U], C[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + (size_t)(r + u * R) * ch;
      A[u] = cn_ldg4_ordered(a + i); B[u] = cn_ldg4_ordered(b + i); C[u] = cn_ldg4_ordered(c + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(A[u], B[u], C[u], base + (size_t)(r + u * R) * ch);
  }
  for (; r < pend; r += R) {
    const size_t i = base + (size_t)r * ch;
    one(cn_ldg4_ordered(a + i), cn_ldg4_ordered(b + i), cn_ldg4_ordered(c + i), i);
  }
}
extern "C" int cn_chan_affine_pair(const float* a, const float* b, const float* c, const float* coef_a, const float* coef_g,
                                   int n, int p, int ch, float alpha, float* out_a, float* out_g, void* stream) {
  CN_REQUIRE(a && b && c && coef_a && coef_g && out_a && out_g && n > 0 && p > 0 && ch > 0, CN_ERR_BAD_SHAPE,
             "cn_chan_affine_pair: bad arguments");
  CN_REQUIRE(affine_rows_ok(n, ch), CN_ERR_UNSUPPORTED, "cn_chan_affine_pair: channel count needs the two-pass form");
  const int R = 256 / (ch / 4);
  int psplit = (8 * 148 + n - 1) / n;
  int maxsplit = p / (2 * R); if (maxsplit < 1) maxsplit = 1;
  if (psplit > maxsplit) psplit = maxsplit;
  chan_affine_pair_kernel<<<dim3(n, psplit), 256, 0, (cudaStream_t)stream>>>(a, b, c, (const float4*)coef_a, (const float4*)coef_g, p, ch,
                                                                             alpha, out_a, out_g);
  CN_CHECK_LAUNCH();
  return CN_OK;
}

// ------------------------------------------------------------------------------------------------
// Coefficient kernels: one thread per channel, loop over samples (so per-channel parameter
// gradients need no second reduction).  S(n,c,j) = sums[(n*ch+c)*8+j] (7 sums, padded to 32 bytes so that a slice is two
// 16-byte loads), N = pixels per sample.
// ------------------------------------------------------------------------------------------------
enum {
  CN_COEF_IN_FWD = 0,       // p0=gamma p1=beta            -> coef0
  CN_COEF_IN_BWD = 1,       // sums(a,gy) p0=gamma         -> coef0 (input grad), out0=ggamma[c], out1=gbeta[c]
  CN_COEF_IN_BWDBWD = 2,    // sums(a,gy,h) p0=gamma       -> coef0 (d/da), coef1 (d/dgy), out0=ggamma[c]
  CN_COEF_STYLE_FWD = 3,    // sums(a)                     -> out0 = style (n,2ch)
  CN_COEF_STYLE_BWD = 4,    // sums(a) p0=gstyle (n,2ch)   -> coef0
  CN_COEF_STYLE_BWDBWD = 5, // sums(a,-,h)[b slot=h] p0=gstyle -> coef0 (d/da), out0 = d/dgstyle (n,2ch)
  CN_COEF_ADAIN_FWD = 6,    // sums(a) p0=sb (n,2ch)       -> coef0
  CN_COEF_ADAIN_BWD = 7,    // sums(a,gy) p0=sb            -> coef0, out0 = gsb (n,2ch)
};

__global__ void __launch_bounds__(1024)
norm_coef_kernel(int kind, const float* __restrict__ sums, int nsplit, const float* __restrict__ p0,
                                 const float* __restrict__ p1, int n, int ch, float N, float eps,
                                 float4* __restrict__ coef0, float4* __restrict__ coef1,
                                 float* __restrict__ out0, float* __restrict__ out1, int spb) {
  // block (8, 128): x = channel (8 records = one 256-byte run), the 128 rows are spb samples x G = 128 / spb slice groups
  // (spb = 32, 16, 8 or 4, the largest that the batch fills).  The kernel is a chain of dependent L2 round trips, not
  // bandwidth: a row adds the slices g, g + G, ... of its sample, the G partial records are folded in group order through
  // shared memory, then the row of group 0 evaluates the closed form (fixed order everywhere: bit-reproducible).
  // grid = ch / 8 blocks (6 .. 96 on the step's layers; 32-channel blocks left the GPU to 2 .. 24 blocks of serial walks).
  __shared__ float red[2][32][9];
  __shared__ float fold[7][128][9];
  const int G = 128 / spb;
  const int si = threadIdx.y % spb, gz = threadIdx.y / spb;
  const int c = blockIdx.x * 8 + threadIdx.x;
  const bool cok = c < ch;
  const float invN = 1.f / N;
  float acc0 = 0.f, acc1 = 0.f;
  for (int i0 = 0; i0 < n; i0 += spb) {
    const int i = i0 + si;
    const bool ok = cok && i < n;
    float S[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) S[j] = 0.f;
    if (ok) {
      int z = gz;
      for (; z + 3 * G < nsplit; z += 4 * G) {    // four slices (eight 16-byte loads) in flight
        float4 lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* Sz = sums + (((size_t)(z + u * G) * n + i) * ch + c) * CN_SUMS_LD;
          lo[u] = cn_ldg4_ordered(Sz); hi[u] = cn_ldg4_ordered(Sz + 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          S[0] += lo[u].x; S[1] += lo[u].y; S[2] += lo[u].z; S[3] += lo[u].w; S[4] += hi[u].x; S[5] += hi[u].y; S[6] += hi[u].z;
        }
      }
      for (; z < nsplit; z += G) {
        const float* Sz = sums + (((size_t)z * n + i) * ch + c) * CN_SUMS_LD;
        const float4 lo = cn_ldg4_ordered(Sz), hi = cn_ldg4_ordered(Sz + 4);
        S[0] += lo.x; S[1] += lo.y; S[2] += lo.z; S[3] += lo.w; S[4] += hi.x; S[5] += hi.y; S[6] += hi.z;
      }
    }
    if (G > 1) {
      if (i0 > 0) __syncthreads();
#pragma unroll
      for (int j = 0; j < 7; ++j) fold[j][threadIdx.y][threadIdx.x] = S[j];
      __syncthreads();
      if (gz == 0) {
        for (int g = 1; g < G; ++g)
#pragma unroll
          for (int j = 0; j < 7; ++j) S[j] += fold[j][g * spb + si][threadIdx.x];
      }
    }
    if (!ok || gz != 0) continue;
    const size_t nc = (size_t)i * ch + c;
    const float mu = S[0] * invN;
    float var = S[3] * invN - mu * mu;
    if (var < 0.f) var = 0.f;
    switch (kind) {
      case CN_COEF_IN_FWD: {
        float d = sqrtf(var) + eps, g = p0[c];
        coef0[nc] = make_float4(g / d, 0.f, 0.f, p1[c] - mu * g / d);
      } break;
      case CN_COEF_IN_BWD: {
        float s = sqrtf(var), d = s + eps, g = p0[c];
        float m1 = S[1] * invN, m2 = S[4] * invN - mu * m1;            // mean(gy), mean(gy*xc)
        float ka = (s > 0.f) ? -g * m2 / (s * d * d) : 0.f;
        coef0[nc] = make_float4(ka, g / d, 0.f, -g * m1 / d - ka * mu);
        acc0 += N * m2 / d;                                            // sum gy*xhat
        acc1 += S[1];
      } break;
      case CN_COEF_IN_BWDBWD: {
        float s = sqrtf(var), d = s + eps, g = p0[c];
        float m1 = S[1] * invN, mh = S[2] * invN;
        float m2 = S[4] * invN - mu * m1, mh2 = S[5] * invN - mu * mh;  // mean(gy*xc), mean(h*xc)
        float Ab = S[6] * invN - m1 * mh;                                // mean(h*gy) - mh*m1
        float cx = 0.f, cg = 0.f, chh = 0.f, kb = 0.f;
        if (s > 0.f) {
          float isd2 = 1.f / (s * d * d);
          cx = -g * Ab * isd2 + g * m2 * mh2 * (d + 2.f * s) / (s * s * s * d * d * d);
          cg = -g * mh2 * isd2;
          chh = -g * m2 * isd2;
          kb = -g * mh2 * isd2;
        }
        coef0[nc] = make_float4(cx, cg, chh, -cx * mu - cg * m1 - chh * mh);
        coef1[nc] = make_float4(kb, 0.f, g / d, -g * mh / d - kb * mu);
        acc0 += N * (Ab / d - ((s > 0.f) ? m2 * mh2 / (s * d * d) : 0.f));
      } break;
      case CN_COEF_STYLE_FWD: {
        out0[(size_t)i * 2 * ch + c] = mu;
        out0[(size_t)i * 2 * ch + ch + c] = sqrtf(var + eps);
      } break;
      case CN_COEF_STYLE_BWD: {
        float sd = sqrtf(var + eps);
        float gm = p0[(size_t)i * 2 * ch + c], gs = p0[(size_t)i * 2 * ch + ch + c];
        float ka = gs * invN / sd;
        coef0[nc] = make_float4(ka, 0.f, 0.f, gm * invN - ka * mu);
      } break;
      case CN_COEF_STYLE_BWDBWD: {
        float sd = sqrtf(var + eps);
        float gs = p0[(size_t)i * 2 * ch + ch + c];
        float mh = S[1] * invN, mhx = S[4] * invN - mu * mh;            // mean(h), mean(h*xc)
        float ka = -gs * mhx * invN / (sd * sd * sd), kb = gs * invN / sd;
        coef0[nc] = make_float4(ka, kb, 0.f, -kb * mh - ka * mu);
        out0[(size_t)i * 2 * ch + c] = mh;
        out0[(size_t)i * 2 * ch + ch + c] = mhx / sd;
      } break;
      case CN_COEF_ADAIN_FWD: {
        float r = rsqrtf(var + eps);
        float sc = p0[(size_t)i * 2 * ch + c], bi = p0[(size_t)i * 2 * ch + ch + c];
        float ka = r * (1.f + sc);
        coef0[nc] = make_float4(ka, 0.f, 0.f, bi - mu * ka);
      } break;
      case CN_COEF_ADAIN_BWD: {
        float r = rsqrtf(var + eps);
        float sc = p0[(size_t)i * 2 * ch + c];
        float m1 = S[1] * invN, m2 = S[4] * invN - mu * m1;            // mean(gy), mean(gy*xc)
        float A = r * (1.f + sc);
        float ka = -A * r * r * m2;
        coef0[nc] = make_float4(ka, A, 0.f, -A * m1 - ka * mu);
        out0[(size_t)i * 2 * ch + c] = N * m2 * r;                      // d/ds = sum gy*xhat
        out0[(size_t)i * 2 * ch + ch + c] = S[1];                       // d/db = sum gy
      } break;
    }
  }
  if (kind == CN_COEF_IN_BWD || kind == CN_COEF_IN_BWDBWD) {     // per-channel parameter gradients: fold the 8 rows in order
    if (threadIdx.y < 32) {                                      // rows of group 0 (spb <= 32 of them hold a sample)
      red[0][threadIdx.y][threadIdx.x] = threadIdx.y < spb ? acc0 : 0.f;
      red[1][threadIdx.y][threadIdx.x] = threadIdx.y < spb ? acc1 : 0.f;
    }
    __syncthreads();
    if (threadIdx.y == 0 && cok) {
      float t0 = 0.f, t1 = 0.f;
      for (int i = 0; i < 32; ++i) { t0 += red[0][i][threadIdx.x]; t1 += red[1][i][threadIdx.x]; }
      out0[c] = t0;
      if (kind == CN_COEF_IN_BWD) out1[c] = t1;
    }
  }
}

extern "C" int cn_norm_coef(int kind, const float* sums, int nsplit, const float* p0, const float* p1, int n, int ch,
                            int npix, float eps, float* coef0, float* coef1, float* out0, float* out1,
                            void* stream) {
  CN_REQUIRE(kind >= 0 && kind <= 7 && sums && nsplit >= 1 && n > 0 && ch > 0 && npix > 0, CN_ERR_BAD_SHAPE, "cn_norm_coef: bad arguments");
  const int spb = n >= 32 ? 32 : n >= 16 ? 16 : n >= 8 ? 8 : 4;
  norm_coef_kernel<<<(ch + 7) / 8, dim3(8, 128), 0, (cudaStream_t)stream>>>(kind, sums, nsplit, p0, p1, n, ch, (float)npix, eps,
                                                                              (float4*)coef0, (float4*)coef1, out0, out1, spb);
  CN_CHECK_LAUNCH();
  return CN_OK;
}
