// Elementwise / gather / reduction kernels of the ConfigNet hot path that are not convolutions:
// activations' backward, 2x2 max-pool (VGG), the rigid 3-D feature rotation, loss reductions,
// uint8 <-> float image conversion, VGG preprocessing and the fused Keras-Adam + EMA update.
// All HBM-bound: float4 accesses, grid-stride loops over a grid sized in multiples of the SM count.
#include "common.cuh"

static inline int grid_for(size_t n, int per_thread = 1) {
  size_t b = (n + (size_t)256 * per_thread - 1) / ((size_t)256 * per_thread);
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void lrelu_fwd_kernel(const float* __restrict__ x, float alpha, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i]; y[i] = v > 0.f ? v : v * alpha;
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ ref, int act, float alpha,
                               float* __restrict__ gx, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float r = ref[i], g = gy[i];
    if (act == CN_ACT_LRELU) g *= (r > 0.f ? 1.f : alpha);
    else if (act == CN_ACT_RELU) g = r > 0.f ? g : 0.f;
    else if (act == CN_ACT_TANH) g *= (1.f - r * r);
    gx[i] = g;
  }
}
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b,
                             float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}

extern "C" int cn_lrelu_fwd(const float* x, float alpha, float* y, int64_t n, void* stream) {
  if (n <= 0) return CN_OK;
  lrelu_fwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, alpha, y, (size_t)n);
  CN_CHECK_LAUNCH(); return CN_OK;
}
// 16-byte form: four float4 of both operands in flight per thread (the scalar kernel above: 0.5-0.65 of the HBM rate)
__global__ void __launch_bounds__(256)
act_bwd_vec_kernel(const float4* __restrict__ gy, const float4* __restrict__ ref, int act, float alpha,
                   float4* __restrict__ gx, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  auto one = [&](float r, float g) {
    if (act == CN_ACT_LRELU) g *= (r > 0.f ? 1.f : alpha);
    else if (act == CN_ACT_RELU) g = r > 0.f ? g : 0.f;
    else if (act == CN_ACT_TANH) g *= (1.f - r * r);
    return g;
  };
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 r[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      r[u] = cn_ldg4_ordered(reinterpret_cast<const float*>(ref + i + u * stride));
      g[u] = cn_ldg4_ordered(reinterpret_cast<const float*>(gy + i + u * stride));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      gx[i + u * stride] = make_float4(one(r[u].x, g[u].x), one(r[u].y, g[u].y), one(r[u].z, g[u].z), one(r[u].w, g[u].w));
  }
  for (; i < n4; i += stride) {
    const float4 r = ref[i], g = gy[i];
    gx[i] = make_float4(one(r.x, g.x), one(r.y, g.y), one(r.z, g.z), one(r.w, g.w));
  }
}
// Activation backward with a squared-difference loss tapped on the activation (perceptual loss, perceptual_loss.py:61-82:
// loss_l = mean((y - t)^2) on a VGG activation y that also feeds the next layer):
//   gx = (gy + (2 (y - t)) * (gloss[0] * k)) * act'(y)        gy may be NULL (the deepest tapped layer)
// - the loss gradient, its accumulation onto the gradient from the next layer and the activation derivative in ONE pass
// (3 reads + 1 write) instead of reduce_bwd + an add by the autograd engine + act_bwd (9 tensor passes), same arithmetic.
__global__ void __launch_bounds__(256)
act_bwd_sqdiff_kernel(const float4* __restrict__ gy, const float4* __restrict__ y, const float4* __restrict__ t,
                      const float* __restrict__ gloss, float k, int act, float alpha, float4* __restrict__ gx, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float gs = gloss[0] * k;
  auto one = [&](float g, float r, float tv) {
    g = __fadd_rn(g, __fmul_rn(2.f * (r - tv), gs));      // two roundings, as the separate kernels + add do (no FMA contraction)
    if (act == CN_ACT_LRELU) g *= (r > 0.f ? 1.f : alpha);
    else if (act == CN_ACT_RELU) g = r > 0.f ? g : 0.f;
    else if (act == CN_ACT_TANH) g *= (1.f - r * r);
    return g;
  };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 r[4], tv[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      r[u] = cn_ldg4_ordered(reinterpret_cast<const float*>(y + i + u * stride));
      tv[u] = cn_ldg4_ordered(reinterpret_cast<const float*>(t + i + u * stride));
      g[u] = gy ? cn_ldg4_ordered(reinterpret_cast<const float*>(gy + i + u * stride)) : z4;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      gx[i + u * stride] = make_float4(one(g[u].x, r[u].x, tv[u].x), one(g[u].y, r[u].y, tv[u].y), one(g[u].z, r[u].z, tv[u].z),
                                       one(g[u].w, r[u].w, tv[u].w));
  }
  for (; i < n4; i += stride) {
    const float4 r = y[i], tv = t[i], g = gy ? gy[i] : z4;
    gx[i] = make_float4(one(g.x, r.x, tv.x), one(g.y, r.y, tv.y), one(g.z, r.z, tv.z), one(g.w, r.w, tv.w));
  }
}
extern "C" int cn_act_bwd_sqdiff(const float* gy, const float* y, const float* t, const float* gloss, float k, int act,
                                 float alpha, float* gx, int64_t n, void* stream) {
  CN_REQUIRE(y && t && gloss && gx && n > 0, CN_ERR_BAD_SHAPE, "cn_act_bwd_sqdiff: bad arguments");
  CN_REQUIRE((n & 3) == 0 && (((uintptr_t)gy | (uintptr_t)y | (uintptr_t)t | (uintptr_t)gx) & 15) == 0, CN_ERR_BAD_ALIGN,
             "cn_act_bwd_sqdiff: tensors must be 16-byte aligned with a multiple of 4 elements");
  const size_t n4 = (size_t)n / 4;
  size_t blocks = (n4 + 1023) / 1024; if (blocks > 8 * 148) blocks = 8 * 148; if (blocks < 1) blocks = 1;
  act_bwd_sqdiff_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)gy, (const float4*)y, (const float4*)t, gloss, k,
                                                                           act, alpha, (float4*)gx, n4);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_act_bwd(const float* gy, const float* y, int act, float alpha, float* gx, int64_t n, void* stream) {
  if (n <= 0) return CN_OK;
  if ((n & 3) == 0 && n >= 4096 && (((uintptr_t)gy | (uintptr_t)y | (uintptr_t)gx) & 15) == 0) {
    const size_t n4 = (size_t)n / 4;
    size_t blocks = (n4 + 1023) / 1024; if (blocks > 8 * 148) blocks = 8 * 148;
    act_bwd_vec_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)gy, (const float4*)y, act, alpha, (float4*)gx, n4);
    CN_CHECK_LAUNCH(); return CN_OK;
  }
  act_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(gy, y, act, alpha, gx, (size_t)n);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_axpby(const float* x, const float* y, float a, float b, float* out, int64_t n, void* stream) {
  if (n <= 0) return CN_OK;
  axpby_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, a, b, out, (size_t)n);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ max-pool 2x2/s2 (NHWC)
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, int h, int w, int c, float* __restrict__ y, size_t total) {
  const int oh = h / 2, ow = w / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % c); size_t t = i / c;
    int ox = (int)(t % ow); t /= ow;
    int oy = (int)(t % oh); size_t n = t / oh;
    const float* p = x + ((n * h + 2 * oy) * w + 2 * ox) * c + ch;
    float m = fmaxf(fmaxf(p[0], p[c]), fmaxf(p[(size_t)w * c], p[(size_t)w * c + c]));
    y[i] = m;
  }
}
// gradient goes to the first element equal to the maximum in (dy,dx) scan order
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                                    int h, int w, int c, float* __restrict__ gx, size_t total) {
  const int oh = h / 2, ow = w / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % c); size_t t = i / c;
    int ox = (int)(t % ow); t /= ow;
    int oy = (int)(t % oh); size_t n = t / oh;
    size_t base = ((n * h + 2 * oy) * w + 2 * ox) * c + ch;
    float m = y[i], g = gy[i];
    bool done = false;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        size_t j = base + ((size_t)dy * w + dx) * c;
        bool hit = !done && x[j] == m;
        gx[j] = hit ? g : 0.f;
        done = done || hit;
      }
  }
}
// 16-byte forms (c % 4 == 0): thread = (output pixel, channel quad); the four window loads of a thread are issued together,
// consecutive threads cover consecutive 16-byte pieces.  Same arithmetic as the scalar kernels (max is exact, the gradient
// goes to the first window element equal to the maximum in (dy, dx) order).
__global__ void __launch_bounds__(256)
maxpool2_fwd_vec_kernel(const float* __restrict__ x, int h, int w, int c4, float* __restrict__ y, size_t total4) {
  const int oh = h / 2, ow = w / 2;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int q = (int)(i % c4); size_t t = i / c4;
  const int ox = (int)(t % ow); t /= ow;
  const int oy = (int)(t % oh); const size_t n = t / oh;
  const float* p = x + (((n * h + 2 * oy) * w + 2 * ox) * c4 + q) * 4;
  const size_t cs = (size_t)c4 * 4, rs = (size_t)w * cs;
  const float4 a = cn_ldg4_ordered(p), b = cn_ldg4_ordered(p + cs), c = cn_ldg4_ordered(p + rs), d = cn_ldg4_ordered(p + rs + cs);
  float4 m;
  m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)); m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
  m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)); m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
  reinterpret_cast<float4*>(y)[i] = m;
}
__global__ void __launch_bounds__(256)
maxpool2_bwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                        int h, int w, int c4, float* __restrict__ gx, size_t total4) {
  const int oh = h / 2, ow = w / 2;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int q = (int)(i % c4); size_t t = i / c4;
  const int ox = (int)(t % ow); t /= ow;
  const int oy = (int)(t % oh); const size_t n = t / oh;
  const size_t base = (((n * h + 2 * oy) * w + 2 * ox) * c4 + q) * 4;
  const size_t cs = (size_t)c4 * 4, rs = (size_t)w * cs;
  const float4 xv[4] = {cn_ldg4_ordered(x + base), cn_ldg4_ordered(x + base + cs), cn_ldg4_ordered(x + base + rs),
                        cn_ldg4_ordered(x + base + rs + cs)};
  const float4 m4 = cn_ldg4_ordered(y + 4 * i), g4 = cn_ldg4_ordered(gy + 4 * i);
  const float m[4] = {m4.x, m4.y, m4.z, m4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
  float o[4][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bool done = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xe = e == 0 ? xv[k].x : e == 1 ? xv[k].y : e == 2 ? xv[k].z : xv[k].w;
      const bool hit = !done && xe == m[e];
      o[k][e] = hit ? g[e] : 0.f;
      done = done || hit;
    }
  }
  *reinterpret_cast<float4*>(gx + base) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
  *reinterpret_cast<float4*>(gx + base + cs) = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
  *reinterpret_cast<float4*>(gx + base + rs) = make_float4(o[2][0], o[2][1], o[2][2], o[2][3]);
  *reinterpret_cast<float4*>(gx + base + rs + cs) = make_float4(o[3][0], o[3][1], o[3][2], o[3][3]);
}
static inline bool mp_vec_ok(int c, const void* a, const void* b, const void* d = nullptr, const void* e = nullptr) {
  return c % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)d | (uintptr_t)e) & 15) == 0;
}
extern "C" int cn_maxpool2_fwd(const float* x, int n, int h, int w, int c, float* y, void* stream) {
  CN_REQUIRE(h % 2 == 0 && w % 2 == 0, CN_ERR_BAD_SHAPE, "maxpool2: odd spatial size");
  size_t total = (size_t)n * (h / 2) * (w / 2) * c;
  if (mp_vec_ok(c, x, y) && total / 4 < ((size_t)1 << 31) * 256) {
    maxpool2_fwd_vec_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, h, w, c / 4, y, total / 4);
    CN_CHECK_LAUNCH(); return CN_OK;
  }
  maxpool2_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, h, w, c, y, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_maxpool2_bwd(const float* x, const float* y, const float* gy, int n, int h, int w, int c,
                               float* gx, void* stream) {
  CN_REQUIRE(h % 2 == 0 && w % 2 == 0, CN_ERR_BAD_SHAPE, "maxpool2: odd spatial size");
  size_t total = (size_t)n * (h / 2) * (w / 2) * c;
  if (mp_vec_ok(c, x, y, gy, gx) && total / 4 < ((size_t)1 << 31) * 256) {
    maxpool2_bwd_vec_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, gy, h, w, c / 4, gx, total / 4);
    CN_CHECK_LAUNCH(); return CN_OK;
  }
  maxpool2_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, gy, h, w, c, gx, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ rotate3d
// confignet_utils.py:63-120.  One warp per output voxel, lanes over channels.
struct RotCorner { int idx[8]; float d0, d1, d2; };

__device__ __forceinline__ RotCorner rot_corners(const float* __restrict__ R, int s, int x, int y, int z) {
  const float ctr = (s - 1) * 0.5f, hi = (float)(s - 1);
  const float px = x - ctr, py = y - ctr, pz = z - ctr;
  float t0 = R[0] * px + R[1] * py + R[2] * pz + ctr;
  float t1 = R[3] * px + R[4] * py + R[5] * pz + ctr;
  float t2 = R[6] * px + R[7] * py + R[8] * pz + ctr;
  t0 = fminf(fmaxf(t0, 0.f), hi); t1 = fminf(fmaxf(t1, 0.f), hi); t2 = fminf(fmaxf(t2, 0.f), hi);
  const float f0 = floorf(t0), f1 = floorf(t1), f2 = floorf(t2);
  const int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2;
  const int c0 = min(i0 + 1, s - 1), c1 = min(i1 + 1, s - 1), c2 = min(i2 + 1, s - 1);
  RotCorner r;
  r.d0 = t0 - f0; r.d1 = t1 - f1; r.d2 = t2 - f2;
  // order: 000 100 010 110 001 101 011 111  (bit0 = axis0 ceil, bit1 = axis1 ceil, bit2 = axis2 ceil)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int a = (k & 1) ? c0 : i0, b = (k & 2) ? c1 : i1, c = (k & 4) ? c2 : i2;
    r.idx[k] = (a * s + b) * s + c;
  }
  return r;
}

__global__ void rotate3d_fwd_kernel(const float* __restrict__ grid, const float* __restrict__ rot, int s, int c,
                                    float* __restrict__ out, int nvox_total) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nvox_total) return;
  const int s3 = s * s * s;
  const int b = warp / s3, v = warp % s3;
  const int z = v % s, y = (v / s) % s, x = v / (s * s);
  RotCorner rc = rot_corners(rot + b * 9, s, x, y, z);
  const float* g = grid + (size_t)b * s3 * c;
  float* o = out + (size_t)warp * c;
  const float d0 = rc.d0, d1 = rc.d1, d2 = rc.d2;
  for (int ch = lane; ch < c; ch += 32) {
    float v000 = g[(size_t)rc.idx[0] * c + ch], v100 = g[(size_t)rc.idx[1] * c + ch];
    float v010 = g[(size_t)rc.idx[2] * c + ch], v110 = g[(size_t)rc.idx[3] * c + ch];
    float v001 = g[(size_t)rc.idx[4] * c + ch], v101 = g[(size_t)rc.idx[5] * c + ch];
    float v011 = g[(size_t)rc.idx[6] * c + ch], v111 = g[(size_t)rc.idx[7] * c + ch];
    float c00 = v000 * (1.f - d0) + v100 * d0, c01 = v001 * (1.f - d0) + v101 * d0;
    float c10 = v010 * (1.f - d0) + v110 * d0, c11 = v011 * (1.f - d0) + v111 * d0;
    float c0 = c00 * (1.f - d1) + c10 * d1, c1 = c01 * (1.f - d1) + c11 * d1;
    o[ch] = c0 * (1.f - d2) + c1 * d2;
  }
}
// Gradient wrt the volume: a scatter (every output voxel adds into its 8 clamped corners; clamping piles whole rays onto
// the faces, so there is no bounded gather form).  Floating-point atomics would make the result depend on the order
// in which the warps arrive; the contributions are therefore accumulated as 64-bit FIXED-POINT integers (integer
// addition is associative: bit-reproducible) with a scale taken from max|gout|: 2^46 steps below the largest
// gradient, far finer than the fp32 rounding of the contributions themselves.
__global__ void absmax_bits_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bit patterns
}
__device__ __forceinline__ int fx_exponent(unsigned amax_bits) {       // e with max|g| < 2^e
  int e = 0;
  frexpf(__uint_as_float(amax_bits), &e);
  return e;
}
__global__ void rotate3d_bwd_grid_kernel(const float* __restrict__ gout, const float* __restrict__ rot, int s, int c,
                                         unsigned long long* __restrict__ acc, int nvox_total, const unsigned* __restrict__ amax_bits) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nvox_total) return;
  const unsigned ab = *amax_bits;
  if (ab == 0u || ab >= 0x7f800000u) return;             // all-zero (or non-finite) gradient: the accumulator stays zero
  const float scale = ldexpf(1.f, 46 - fx_exponent(ab));
  const int s3 = s * s * s;
  const int b = warp / s3, v = warp % s3;
  const int z = v % s, y = (v / s) % s, x = v / (s * s);
  RotCorner rc = rot_corners(rot + b * 9, s, x, y, z);
  unsigned long long* g = acc + (size_t)b * s3 * c;
  const float* go = gout + (size_t)warp * c;
  float wt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    wt[k] = ((k & 1) ? rc.d0 : 1.f - rc.d0) * ((k & 2) ? rc.d1 : 1.f - rc.d1) * ((k & 4) ? rc.d2 : 1.f - rc.d2);
  for (int ch = lane; ch < c; ch += 32) {
    float gv = go[ch];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long q = __float2ll_rn(gv * wt[k] * scale);
      if (q != 0) atomicAdd(g + (size_t)rc.idx[k] * c + ch, (unsigned long long)q);
    }
  }
}
__global__ void fx_to_float_kernel(const long long* __restrict__ acc, size_t n, const unsigned* __restrict__ amax_bits, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned ab = *amax_bits;
  if (ab == 0u || ab >= 0x7f800000u) { out[i] = 0.f; return; }
  out[i] = (float)((double)acc[i] * ldexp(1.0, fx_exponent(ab) - 46));
}
extern "C" int cn_rotate3d_fwd(const float* grid, const float* rot, int b, int s, int c, float* out, void* stream) {
  int nv = b * s * s * s;
  rotate3d_fwd_kernel<<<(nv * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grid, rot, s, c, out, nv);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_rotate3d_bwd_grid(const float* gout, const float* rot, int b, int s, int c, float* ggrid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = b * s * s * s;
  const size_t n = (size_t)nv * c;
  if (n == 0) return CN_OK;
  static char key;
  float* ws = nullptr;
  int rc = cn_scratch(&key, n * 8 + 16, &ws); if (rc) return rc;        // [amax bits (16 B)] [int64 accumulators]
  unsigned* amax = reinterpret_cast<unsigned*>(ws);
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(ws + 4);
  CN_CHECK_CUDA(cudaMemsetAsync(ws, 0, n * 8 + 16, st));
  int blocks = (int)((n + 255) / 256); if (blocks > 8 * 148) blocks = 8 * 148;
  absmax_bits_kernel<<<blocks, 256, 0, st>>>(gout, n, amax);
  CN_CHECK_LAUNCH();
  rotate3d_bwd_grid_kernel<<<(nv * 32 + 255) / 256, 256, 0, st>>>(gout, rot, s, c, acc, nv, amax);
  CN_CHECK_LAUNCH();
  fx_to_float_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const long long*>(acc), n, amax, ggrid);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ reductions
#define CN_RED_BLOCKS 592
extern "C" int cn_reduce_ws_floats(void) { return CN_RED_BLOCKS; }

__device__ __forceinline__ float red_term(int kind, float x, float y, float sign) {
  if (kind == 0) { float z = sign * x; return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }   // softplus
  if (kind == 1) { float d = x - y; return d * d; }
  if (kind == 2) return x * x;
  return x;
}
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sm[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sm[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;   // valid in thread 0
}
// weighted variant: term *= wgt[i / wdiv] (eye loss mask broadcast over channels)
__global__ void reduce_stage1_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ wgt,
                                     int wdiv, size_t n, int kind, float sign, float* __restrict__ ws) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float t = red_term(kind, x[i], y ? y[i] : 0.f, sign);
    if (wgt) t *= wgt[i / wdiv];
    acc += t;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) ws[blockIdx.x] = acc;
}
// 16-byte form for the large unweighted reductions (the perceptual-loss terms read 2 x 268 MB per call): four independent
// float4 loads per operand in flight per thread; the scalar kernel above ran at 2.2 TB/s (one 4-byte load in flight).
__global__ void __launch_bounds__(256)
reduce_stage1_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ y, size_t n4, int kind, float sign,
                         float* __restrict__ ws) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto term4 = [&](const float4& a, const float4& b) {
    return (red_term(kind, a.x, b.x, sign) + red_term(kind, a.y, b.y, sign)) + (red_term(kind, a.z, b.z, sign) + red_term(kind, a.w, b.w, sign));
  };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (; i + 3 * stride < n4; i += 4 * stride) {
    const float4 a0 = x[i], a1 = x[i + stride], a2 = x[i + 2 * stride], a3 = x[i + 3 * stride];
    float4 b0 = z4, b1 = z4, b2 = z4, b3 = z4;
    if (y) { b0 = y[i]; b1 = y[i + stride]; b2 = y[i + 2 * stride]; b3 = y[i + 3 * stride]; }
    acc += (term4(a0, b0) + term4(a1, b1)) + (term4(a2, b2) + term4(a3, b3));
  }
  for (; i < n4; i += stride) acc += term4(x[i], y ? y[i] : z4);
  acc = block_sum(acc);
  if (threadIdx.x == 0) ws[blockIdx.x] = acc;
}
__global__ void reduce_stage2_kernel(const float* __restrict__ ws, int nb, float scale, float* __restrict__ result) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += ws[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) result[0] = acc * scale;
}
// few elements (discriminator scores, latent losses: 32 .. a few thousand): both stages in ONE single-block launch
__global__ void __launch_bounds__(256)
reduce_small_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ wgt,
                    int wdiv, int n, int kind, float sign, float scale, float* __restrict__ result) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    float t = red_term(kind, x[i], y ? y[i] : 0.f, sign);
    if (wgt) t *= wgt[i / wdiv];
    acc += t;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) result[0] = acc * scale;
}
extern "C" int cn_reduce(const float* x, const float* y, const float* wgt, int wdiv, int64_t n, int kind, float sign,
                         float scale, float* ws, float* result, void* stream) {
  CN_REQUIRE(x && ws && result && n > 0 && kind >= 0 && kind <= 3, CN_ERR_BAD_SHAPE, "cn_reduce: bad arguments");
  if (n <= 8192) {
    reduce_small_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(x, y, wgt, wdiv > 0 ? wdiv : 1, (int)n, kind, sign, scale, result);
    CN_CHECK_LAUNCH(); return CN_OK;
  }
  int nb = grid_for(n, 4); if (nb > CN_RED_BLOCKS) nb = CN_RED_BLOCKS;
  const bool vec = wgt == nullptr && (n & 3) == 0 && n >= 4096 && ((uintptr_t)x & 15) == 0 && (y == nullptr || ((uintptr_t)y & 15) == 0);
  if (vec) reduce_stage1_vec_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)y, (size_t)n / 4, kind, sign, ws);
  else reduce_stage1_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(x, y, wgt, wdiv > 0 ? wdiv : 1, (size_t)n, kind, sign, ws);
  CN_CHECK_LAUNCH();
  reduce_stage2_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ws, nb, scale, result);
  CN_CHECK_LAUNCH(); return CN_OK;
}
__global__ void reduce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ wgt,
                                  int wdiv, size_t n, int kind, float sign, float k, const float* __restrict__ gscale,
                                  float* __restrict__ gx) {
  const float gs = gscale[0] * k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float xv = x[i], g;
    if (kind == 0) { float z = sign * xv; g = sign / (1.f + expf(-z)); }
    else if (kind == 1) g = 2.f * (xv - y[i]);
    else if (kind == 2) g = 2.f * xv;
    else g = 1.f;
    if (wgt) g *= wgt[i / wdiv];
    gx[i] = g * gs;
  }
}
__global__ void __launch_bounds__(256)
reduce_bwd_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ y, size_t n4, int kind, float sign, float k,
                      const float* __restrict__ gscale, float4* __restrict__ gx) {
  const float gs = gscale[0] * k;
  auto g1 = [&](float xv, float yv) {
    float g;
    if (kind == 0) { float z = sign * xv; g = sign / (1.f + expf(-z)); }
    else if (kind == 1) g = 2.f * (xv - yv);
    else if (kind == 2) g = 2.f * xv;
    else g = 1.f;
    return g * gs;
  };
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = x[i];
    const float4 b = y ? y[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    gx[i] = make_float4(g1(a.x, b.x), g1(a.y, b.y), g1(a.z, b.z), g1(a.w, b.w));
  }
}
extern "C" int cn_reduce_bwd(const float* x, const float* y, const float* wgt, int wdiv, int64_t n, int kind, float sign,
                             float k, const float* gscale, float* gx, void* stream) {
  CN_REQUIRE(x && gscale && gx && n > 0, CN_ERR_BAD_SHAPE, "cn_reduce_bwd: bad arguments");
  if (wgt == nullptr && (n & 3) == 0 && n >= 4096 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)gx & 15) == 0 &&
      (y == nullptr || ((uintptr_t)y & 15) == 0)) {
    reduce_bwd_vec_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)y, (size_t)n / 4, kind, sign, k,
                                                                             gscale, (float4*)gx);
    CN_CHECK_LAUNCH(); return CN_OK;
  }
  reduce_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, wgt, wdiv > 0 ? wdiv : 1, (size_t)n, kind, sign, k, gscale, gx);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ images
__global__ void to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = fminf(fmaxf(x[i], -1.f), 1.f);
    v = (v + 1.f) * 127.5f;            // same fp32 operation order as NumPy: (clip(x)+1)*127.5
    out[i] = (uint8_t)v;               // truncation, like ndarray.astype(np.uint8) on [0,255]
  }
}
__global__ void from_uint8_kernel(const uint8_t* __restrict__ x, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (float)x[i] / 127.5f - 1.f;
}
extern "C" int cn_to_uint8(const float* x, uint8_t* out, int64_t n, void* stream) {
  if (n <= 0) return CN_OK;
  to_uint8_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, out, (size_t)n);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_from_uint8(const uint8_t* x, float* out, int64_t n, void* stream) {
  if (n <= 0) return CN_OK;
  from_uint8_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, out, (size_t)n);
  CN_CHECK_LAUNCH(); return CN_OK;
}
// (x+1)*127.5, channel flip, subtract caffe BGR means; backward: gx[..., 2-c] = 127.5 * g[..., c]
__global__ void vgg_preprocess_kernel(const float* __restrict__ x, float* __restrict__ out, size_t npix, int mode) {
  const float mean[3] = {103.939f, 116.779f, 123.68f};
  const float face_mean[3] = {93.5940f, 104.7624f, 129.1863f};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const float* p = x + i * 3; float* o = out + i * 3;
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = (p[2 - c] + 1.f) * 127.5f - mean[c];
    } else if (mode == 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) o[2 - c] = 127.5f * p[c];
    } else if (mode == 2) {
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = (p[c] + 1.f) * 127.5f - face_mean[c];
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = 127.5f * p[c];
    }
  }
}
extern "C" int cn_vgg_preprocess(const float* x, float* out, int64_t npix, int backward, void* stream) {
  if (npix <= 0) return CN_OK;
  CN_REQUIRE(backward >= 0 && backward <= 3, CN_ERR_BAD_SHAPE, "cn_vgg_preprocess: mode must be 0..3");
  vgg_preprocess_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(x, out, (size_t)npix, backward);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ Adam + EMA
__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, float* __restrict__ ema, size_t n, float lr_t, float b1,
                                float b2, float eps, float ema_alpha, float gscale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (ema) ema[i] = ema_alpha * ema[i] + (1.f - ema_alpha) * pi;
  }
}
// same update with lr_t read from device memory: the launch can then live in a captured CUDA graph while the host
// advances Keras' bias-corrected learning rate between replays
__global__ void adam_ema_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, float* __restrict__ ema, size_t n, const float* __restrict__ lr_dev,
                                    float b1, float b2, float eps, float ema_alpha, float gscale) {
  const float lr_t = *lr_dev;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (ema) ema[i] = ema_alpha * ema[i] + (1.f - ema_alpha) * pi;
  }
}
extern "C" int cn_adam_ema_step_dev(float* p, const float* g, float* m, float* v, float* ema, int64_t n, const float* lr_t_dev,
                                    float b1, float b2, float eps, float ema_alpha, float gscale, void* stream) {
  if (n <= 0) return CN_OK;
  CN_REQUIRE(p && g && m && v && lr_t_dev, CN_ERR_BAD_SHAPE, "cn_adam_ema_step_dev: null pointer");
  cn_mark_params_changed(p); if (ema) cn_mark_params_changed(ema);
  adam_ema_dev_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, (size_t)n, lr_t_dev, b1, b2, eps, ema_alpha, gscale);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr_t,
                                float b1, float b2, float eps, float ema_alpha, float gscale, void* stream) {
  if (n <= 0) return CN_OK;
  CN_REQUIRE(p && g && m && v, CN_ERR_BAD_SHAPE, "cn_adam_ema_step: null pointer");
  cn_mark_params_changed(p); if (ema) cn_mark_params_changed(ema);
  adam_ema_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, (size_t)n, lr_t, b1, b2, eps, ema_alpha, gscale);
  CN_CHECK_LAUNCH(); return CN_OK;
}
// EMA alone (generator_smoothed when the optimizer step is not fused)
__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ p, size_t n, float alpha) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    ema[i] = alpha * ema[i] + (1.f - alpha) * p[i];
}
extern "C" int cn_ema(float* ema, const float* p, int64_t n, float alpha, void* stream) {
  cn_mark_params_changed(ema);
  if (n <= 0) return CN_OK;
  ema_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(ema, p, (size_t)n, alpha);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ gradient gather
// Packs up to CN_MULTI_MAX separately allocated gradient tensors into one flat buffer (the buffer the
// NCCL all-reduce and the fused Adam kernel run on).  A NULL source zero-fills its slot (unused variable,
// tape.gradient -> None).  blockIdx.y = tensor.
struct MultiCopyArgs {
  const float* src[CN_MULTI_MAX];
  long long off[CN_MULTI_MAX];
  long long n[CN_MULTI_MAX];
};
__global__ void multi_copy_kernel(MultiCopyArgs a, float* __restrict__ dst) {
  const int t = blockIdx.y;
  const float* s = a.src[t];
  float* d = dst + a.off[t];
  const long long n = a.n[t];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = s ? s[i] : 0.f;
}
extern "C" int cn_multi_copy(int count, const float* const* src, const int64_t* dst_off, const int64_t* n,
                             float* dst, void* stream) {
  CN_REQUIRE(count >= 0 && dst, CN_ERR_BAD_SHAPE, "cn_multi_copy: bad arguments");
  for (int base = 0; base < count; base += CN_MULTI_MAX) {
    MultiCopyArgs a;
    int c = count - base < CN_MULTI_MAX ? count - base : CN_MULTI_MAX;
    long long maxn = 1;
    for (int i = 0; i < c; ++i) {
      a.src[i] = src[base + i]; a.off[i] = dst_off[base + i]; a.n[i] = n[base + i];
      if (a.n[i] > maxn) maxn = a.n[i];
    }
    int bx = (int)((maxn + 1023) / 1024); if (bx > 64) bx = 64; if (bx < 1) bx = 1;
    multi_copy_kernel<<<dim3(bx, c), 256, 0, (cudaStream_t)stream>>>(a, dst);
    CN_CHECK_LAUNCH();
  }
  return CN_OK;
}
