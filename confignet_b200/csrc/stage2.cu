// Second-stage / fine-tune pieces of the ConfigNet hot path:
// the ResNet50 encoder's non-convolution layers (inference-mode BatchNorm with trainable gamma/beta fused with the
// residual add and ReLU, the 3x3/s2 max-pool, global average pooling), the differentiable Euler-angle -> matrix
// map and the rotation gradient of the 3-D resampler (the encoder's rotation output is trained through it),
// the batch-normalised latent regression loss, and a per-column scale.  All HBM- or latency-bound.
#include "common.cuh"

static inline int grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------ BatchNorm (inference statistics)
// keras BatchNormalization called without training=True inside the manual tape (real_encoder.py:27):
// y = gamma * (x - moving_mean) / sqrt(moving_var + eps) + beta, eps = 1.001e-5 [TF-2.1 keras-applications].
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int c, float* __restrict__ scale, float* __restrict__ shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    float r = 1.f / sqrtf(var[i] + eps);
    float s = gamma[i] * r;
    scale[i] = s; shift[i] = beta[i] - mean[i] * s;
  }
}
extern "C" int cn_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int c,
                          float* scale, float* shift, void* stream) {
  bn_fold_kernel<<<(c + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, c, scale, shift);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// out = relu?(x * scale[c] + shift[c] + residual)
__global__ void bn_act_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                  const float4* __restrict__ res, int relu, float4* __restrict__ out, size_t n4, int c4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c4) * 4;
    float4 v = x[i];
    const float4 s = *reinterpret_cast<const float4*>(scale + ch), t = *reinterpret_cast<const float4*>(shift + ch);
    v.x = v.x * s.x + t.x; v.y = v.y * s.y + t.y; v.z = v.z * s.z + t.z; v.w = v.w * s.w + t.w;
    if (res) { const float4 r = res[i]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    out[i] = v;
  }
}
extern "C" int cn_bn_act_fwd(const float* x, const float* scale, const float* shift, const float* residual, int relu,
                             float* out, int64_t npix, int c, void* stream) {
  CN_REQUIRE(c % 4 == 0, CN_ERR_BAD_ALIGN, "cn_bn_act_fwd: channels must be a multiple of 4");
  size_t n4 = (size_t)npix * (c / 4);
  if (n4 == 0) return CN_OK;
  bn_act_fwd_kernel<<<grid_for(n4), 256, 0, (cudaStream_t)stream>>>((const float4*)x, scale, shift, (const float4*)residual, relu,
                                                                    (float4*)out, n4, c / 4);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// g = relu ? gout * (out > 0) : gout;  gx = g * scale[c];  gres = g;  dgamma[c] += sum g * (x - mean) * rstd;  dbeta[c] += sum g
// grid (ceil(C/32), row splits), block (32, 8): lane = channel (coalesced 128-byte rows), 8 pixel rows in flight
__global__ void bn_act_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out, const float* __restrict__ x,
                                  const float* __restrict__ scale, const float* __restrict__ mean, const float* __restrict__ var,
                                  float eps, int relu, float* __restrict__ gx, float* __restrict__ gres,
                                  float* __restrict__ dgamma, float* __restrict__ dbeta, int npix, int c) {
  __shared__ float sg[8][33], sb[8][33];
  const int ch = blockIdx.x * 32 + threadIdx.x;
  float ag = 0.f, ab = 0.f;
  if (ch < c) {
    const float sc = scale[ch], mu = mean[ch], rstd = 1.f / sqrtf(var[ch] + eps);
    for (int r = blockIdx.y * 8 + threadIdx.y; r < npix; r += 8 * gridDim.y) {
      const size_t i = (size_t)r * c + ch;
      float g = gout[i];
      if (relu && !(out[i] > 0.f)) g = 0.f;
      gx[i] = g * sc;
      if (gres) gres[i] = g;
      ag = fmaf(g, (x[i] - mu) * rstd, ag);
      ab += g;
    }
  }
  sg[threadIdx.y][threadIdx.x] = ag; sb[threadIdx.y][threadIdx.x] = ab;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += sg[i][threadIdx.x]; b += sb[i][threadIdx.x]; }
    // slab blockIdx.y of the partial buffer: [dgamma (c) | dbeta (c)]; cn_sum_slabs adds the slabs in order
    dgamma[(size_t)blockIdx.y * 2 * c + ch] = a; dbeta[(size_t)blockIdx.y * 2 * c + ch] = b;
  }
}
// dgamma / dbeta <- the two halves of the summed slab
__global__ void bn_split_sums_kernel(const float* __restrict__ sum2, int c, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) { dgamma[i] = sum2[i]; dbeta[i] = sum2[c + i]; }
}
extern "C" int cn_bn_act_bwd(const float* gout, const float* out, const float* x, const float* scale, const float* mean,
                             const float* var, float eps, int relu, float* gx, float* gres, float* dgamma, float* dbeta,
                             int64_t npix, int c, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (npix == 0) {
    CN_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)c * sizeof(float), st));
    CN_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)c * sizeof(float), st));
    return CN_OK;
  }
  int ysplit = (int)((npix + 63) / 64); if (ysplit > 148 * 4) ysplit = 148 * 4; if (ysplit < 1) ysplit = 1;
  dim3 grid((c + 31) / 32, ysplit), block(32, 8);
  // per-row-split partial sums (no atomics), then a fixed-order sum: the gamma / beta gradients are bit-reproducible
  static char key;
  float* ws = nullptr;
  int rc = cn_scratch(&key, ((size_t)ysplit + 1) * 2 * c * sizeof(float), &ws); if (rc) return rc;
  bn_act_bwd_kernel<<<grid, block, 0, st>>>(gout, out, x, scale, mean, var, eps, relu, gx, gres, ws, ws + c, (int)npix, c);
  CN_CHECK_LAUNCH();
  float* sum2 = ws + (size_t)ysplit * 2 * c;
  rc = cn_sum_slabs(ws, ysplit, 2 * c, sum2, st); if (rc) return rc;
  bn_split_sums_kernel<<<(c + 255) / 256, 256, 0, st>>>(sum2, c, dgamma, dbeta);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ ZeroPadding2D(1) + MaxPool 3x3 / s2 VALID
// keras-applications ResNet50 'pool1_pad' + 'pool1_pool': the padded zeros take part in the maximum.
__global__ void maxpool3s2_fwd_kernel(const float* __restrict__ x, int h, int w, int c, int oh, int ow,
                                      float* __restrict__ y, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % c); size_t t = i / c;
    int ox = (int)(t % ow); t /= ow;
    int oy = (int)(t % oh); size_t n = t / oh;
    float m = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        int iy = 2 * oy - 1 + dy, ix = 2 * ox - 1 + dx;
        float v = ((unsigned)iy < (unsigned)h && (unsigned)ix < (unsigned)w) ? x[((n * h + iy) * w + ix) * c + ch] : 0.f;
        m = fmaxf(m, v);
      }
    y[i] = m;
  }
}
// gather form (deterministic): an input element receives the gradient of every window in which it is the first
// maximum in (dy,dx) scan order, padded zeros included (their share is dropped, as ZeroPadding2D's backward slices)
__global__ void maxpool3s2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                                      int h, int w, int c, int oh, int ow, float* __restrict__ gx, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % c); size_t t = i / c;
    int ix = (int)(t % w); t /= w;
    int iy = (int)(t % h); size_t n = t / h;
    const float xv = x[i];
    float acc = 0.f;
    for (int oy = (iy) / 2; oy <= (iy + 1) / 2; ++oy) {
      if (oy >= oh) continue;
      for (int ox = (ix) / 2; ox <= (ix + 1) / 2; ++ox) {
        if (ox >= ow) continue;
        const size_t o = ((n * oh + oy) * ow + ox) * c + ch;
        const float m = y[o];
        if (xv != m) continue;
        const int my = iy - (2 * oy - 1), mx = ix - (2 * ox - 1);      // position inside the window
        bool first = true;
        for (int q = 0; q < my * 3 + mx && first; ++q) {
          int jy = 2 * oy - 1 + q / 3, jx = 2 * ox - 1 + q % 3;
          float v = ((unsigned)jy < (unsigned)h && (unsigned)jx < (unsigned)w) ? x[((n * h + jy) * w + jx) * c + ch] : 0.f;
          if (v == m) first = false;
        }
        if (first) acc += gy[o];
      }
    }
    gx[i] = acc;
  }
}
extern "C" int cn_maxpool3s2_fwd(const float* x, int n, int h, int w, int c, float* y, void* stream) {
  const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
  size_t total = (size_t)n * oh * ow * c;
  if (total == 0) return CN_OK;
  maxpool3s2_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, h, w, c, oh, ow, y, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_maxpool3s2_bwd(const float* x, const float* y, const float* gy, int n, int h, int w, int c, float* gx, void* stream) {
  const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
  size_t total = (size_t)n * h * w * c;
  if (total == 0) return CN_OK;
  maxpool3s2_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, gy, h, w, c, oh, ow, gx, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ global average pooling
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, int p, int c, float* __restrict__ y, int total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int n = i / c, ch = i % c;
  const float* src = x + (size_t)n * p * c + ch;
  float acc = 0.f;
  for (int q = 0; q < p; ++q) acc += src[(size_t)q * c];
  y[i] = acc / (float)p;
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ gy, int p, int c, float* __restrict__ gx, size_t total) {
  const float inv = 1.f / (float)p;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % c); size_t n = i / ((size_t)p * c);
    gx[i] = gy[n * c + ch] * inv;
  }
}
extern "C" int cn_avgpool_fwd(const float* x, int n, int p, int c, float* y, void* stream) {
  int total = n * c;
  if (total == 0) return CN_OK;
  avgpool_fwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, p, c, y, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_avgpool_bwd(const float* gy, int n, int p, int c, float* gx, void* stream) {
  size_t total = (size_t)n * p * c;
  if (total == 0) return CN_OK;
  avgpool_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(gy, p, c, gx, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ per-column scale
__global__ void col_scale_kernel(const float* __restrict__ x, const float* __restrict__ s, int c, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = x[i] * s[i % c];
}
extern "C" int cn_col_scale(const float* x, const float* scale, int rows, int c, float* out, void* stream) {
  size_t n = (size_t)rows * c;
  if (n == 0) return CN_OK;
  col_scale_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, scale, c, out, n);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ Euler angles -> rotation matrix
// confignet_utils.py:122-145
__global__ void euler_fwd_kernel(const float* __restrict__ a, int b, float* __restrict__ r) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  float s0, c0, s1, c1, s2, c2;
  sincosf(a[3 * i], &s0, &c0); sincosf(a[3 * i + 1], &s1, &c1); sincosf(a[3 * i + 2], &s2, &c2);
  float* m = r + 9 * i;
  m[0] = c2 * c1; m[1] = -s2; m[2] = c2 * s1;
  m[3] = s0 * s1 + c0 * c1 * s2; m[4] = c0 * c2; m[5] = c0 * s2 * s1 - c1 * s0;
  m[6] = c1 * s0 * s2 - c0 * s1; m[7] = c2 * s0; m[8] = c0 * c1 + s0 * s1 * s2;
}
__global__ void euler_bwd_kernel(const float* __restrict__ a, const float* __restrict__ g, int b, float* __restrict__ ga) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  float s0, c0, s1, c1, s2, c2;
  sincosf(a[3 * i], &s0, &c0); sincosf(a[3 * i + 1], &s1, &c1); sincosf(a[3 * i + 2], &s2, &c2);
  const float* q = g + 9 * i;
  ga[3 * i] = q[3] * (c0 * s1 - s0 * c1 * s2) + q[4] * (-s0 * c2) + q[5] * (-s0 * s2 * s1 - c1 * c0) +
              q[6] * (c1 * c0 * s2 + s0 * s1) + q[7] * (c2 * c0) + q[8] * (-s0 * c1 + c0 * s1 * s2);
  ga[3 * i + 1] = q[0] * (-c2 * s1) + q[2] * (c2 * c1) + q[3] * (s0 * c1 - c0 * s1 * s2) + q[5] * (c0 * s2 * c1 + s1 * s0) +
                  q[6] * (-s1 * s0 * s2 - c0 * c1) + q[8] * (-c0 * s1 + s0 * c1 * s2);
  ga[3 * i + 2] = q[0] * (-s2 * c1) + q[1] * (-c2) + q[2] * (-s2 * s1) + q[3] * (c0 * c1 * c2) + q[4] * (-c0 * s2) +
                  q[5] * (c0 * c2 * s1) + q[6] * (c1 * s0 * c2) + q[7] * (-s2 * s0) + q[8] * (s0 * s1 * c2);
}
extern "C" int cn_euler_fwd(const float* angles, int b, float* rot, void* stream) {
  if (b == 0) return CN_OK;
  euler_fwd_kernel<<<(b + 127) / 128, 128, 0, (cudaStream_t)stream>>>(angles, b, rot);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_euler_bwd(const float* angles, const float* grot, int b, float* gangles, void* stream) {
  if (b == 0) return CN_OK;
  euler_bwd_kernel<<<(b + 127) / 128, 128, 0, (cudaStream_t)stream>>>(angles, grot, b, gangles);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ rotate3d: gradient wrt the matrix
// transform_3d_grid_tf (confignet_utils.py:63-120): out = trilinear(grid, clip(R (p - ctr) + ctr)).  d out / d R[i][j] =
// (d out / d diff_i) * [0 <= t_i <= S-1] * (p_j - ctr); floor() carries no gradient.  One warp per output voxel
// (lanes over channels), 8 voxels of one sample per block, 9 atomics per block.
__global__ void rotate3d_bwd_rot_kernel(const float* __restrict__ grid, const float* __restrict__ gout, const float* __restrict__ rot,
                                        int s, int c, float* __restrict__ grot, int nvox_total) {
  __shared__ float red[8][9];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + wl;
  const int s3 = s * s * s;
  float out9[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) out9[k] = 0.f;
  const int b = min(warp, nvox_total - 1) / s3;
  if (warp < nvox_total) {
    const int v = warp % s3;
    const int z = v % s, y = (v / s) % s, x = v / (s * s);
    const float* R = rot + b * 9;
    const float ctr = (s - 1) * 0.5f, hi = (float)(s - 1);
    const float q[3] = {x - ctr, y - ctr, z - ctr};
    float d[3]; int fl[3], ce[3]; bool pass[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float ti = R[3 * i] * q[0] + R[3 * i + 1] * q[1] + R[3 * i + 2] * q[2] + ctr;
      pass[i] = ti >= 0.f && ti <= hi;
      ti = fminf(fmaxf(ti, 0.f), hi);
      const float f = floorf(ti);
      d[i] = ti - f; fl[i] = (int)f; ce[i] = min(fl[i] + 1, s - 1);
    }
    const float* g = grid + (size_t)b * s3 * c;
    const float* go = gout + (size_t)warp * c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
      auto at = [&](int i0, int i1, int i2) { return g[(size_t)((i0 * s + i1) * s + i2) * c + ch]; };
      const float v000 = at(fl[0], fl[1], fl[2]), v100 = at(ce[0], fl[1], fl[2]), v010 = at(fl[0], ce[1], fl[2]), v110 = at(ce[0], ce[1], fl[2]);
      const float v001 = at(fl[0], fl[1], ce[2]), v101 = at(ce[0], fl[1], ce[2]), v011 = at(fl[0], ce[1], ce[2]), v111 = at(ce[0], ce[1], ce[2]);
      const float c00 = v000 * (1.f - d[0]) + v100 * d[0], c01 = v001 * (1.f - d[0]) + v101 * d[0];
      const float c10 = v010 * (1.f - d[0]) + v110 * d[0], c11 = v011 * (1.f - d[0]) + v111 * d[0];
      const float c0 = c00 * (1.f - d[1]) + c10 * d[1], c1 = c01 * (1.f - d[1]) + c11 * d[1];
      const float dd0 = ((v100 - v000) * (1.f - d[1]) + (v110 - v010) * d[1]) * (1.f - d[2]) +
                        ((v101 - v001) * (1.f - d[1]) + (v111 - v011) * d[1]) * d[2];
      const float dd1 = (c10 - c00) * (1.f - d[2]) + (c11 - c01) * d[2];
      const float dd2 = c1 - c0;
      const float gv = go[ch];
      a0 = fmaf(gv, dd0, a0); a1 = fmaf(gv, dd1, a1); a2 = fmaf(gv, dd2, a2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    const float gd[3] = {pass[0] ? a0 : 0.f, pass[1] ? a1 : 0.f, pass[2] ? a2 : 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) out9[3 * i + j] = gd[i] * q[j];
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) red[wl][k] = out9[k];
  }
  __syncthreads();
  // the 8 voxels of a block belong to one sample (s^3 is a multiple of 8): one partial row per block
  if (threadIdx.x < 9 && blockIdx.x * 8 < nvox_total) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    grot[(size_t)blockIdx.x * 9 + threadIdx.x] = v;
  }
}
// grot[b][k] = sum over the sample's `per` block partials, fixed order: thread t adds rows t, t+256, ..., then a fixed tree
__global__ void __launch_bounds__(256)
rot_partials_sum_kernel(const float* __restrict__ part, int per, float* __restrict__ grot) {
  __shared__ float red[256][9];
  const float* p = part + (size_t)blockIdx.x * per * 9;
  float a[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) a[k] = 0.f;
  for (int r = threadIdx.x; r < per; r += 256)
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] += p[(size_t)r * 9 + k];
#pragma unroll
  for (int k = 0; k < 9; ++k) red[threadIdx.x][k] = a[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int k = 0; k < 9; ++k) red[threadIdx.x][k] += red[threadIdx.x + o][k];
    __syncthreads();
  }
  if (threadIdx.x < 9) grot[(size_t)blockIdx.x * 9 + threadIdx.x] = red[0][threadIdx.x];
}
extern "C" int cn_rotate3d_bwd_rot(const float* grid, const float* gout, const float* rot, int b, int s, int c, float* grot, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CN_REQUIRE((s * s * s) % 8 == 0, CN_ERR_UNSUPPORTED, "cn_rotate3d_bwd_rot: grid size^3 must be a multiple of 8");
  int nv = b * s * s * s;
  if (nv == 0) return CN_OK;
  static char key;
  float* ws = nullptr;
  int rc = cn_scratch(&key, (size_t)(nv / 8) * 9 * sizeof(float), &ws); if (rc) return rc;
  rotate3d_bwd_rot_kernel<<<nv / 8, 256, 0, st>>>(grid, gout, rot, s, c, ws, nv);
  CN_CHECK_LAUNCH();
  rot_partials_sum_kernel<<<b, 256, 0, st>>>(ws, s * s * s / 8, grot);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ normalised latent regression loss
// compute_normalized_latent_regression_loss (confignet_second_stage.py:93-107).  Per column j with batch means m_l,
// m_o, delta = m_l - m_o, e_b = (l_b - m_l) - (o_b - m_o), den^2 = var_batch(l) + 1e-3 (1 for the last nrot columns):
//   loss = weight / (B J) * sum_j [ B delta^2 + sum_b e_b^2 / den^2 ].
// One thread per column (J = latent_dim + 3 <= 1024), single block.
__global__ void norm_latent_loss_kernel(const float* __restrict__ o, const float* __restrict__ l, int B, int J, int nrot, float weight,
                                        const float* __restrict__ gscale, float* __restrict__ loss, float* __restrict__ g_o,
                                        float* __restrict__ g_l) {
  __shared__ float sm[32];
  const int j = threadIdx.x;
  float S = 0.f;
  if (j < J) {
    float ml = 0.f, mo = 0.f;
    for (int b = 0; b < B; ++b) { ml += l[(size_t)b * J + j]; mo += o[(size_t)b * J + j]; }
    ml /= B; mo /= B;
    float v = 0.f, E = 0.f;
    for (int b = 0; b < B; ++b) {
      const float lp = l[(size_t)b * J + j] - ml, e = lp - (o[(size_t)b * J + j] - mo);
      v += lp * lp; E += e * e;
    }
    v /= B;
    const bool normed = j < J - nrot;
    const float d2 = normed ? v + 1e-3f : 1.f;
    const float delta = ml - mo;
    S = B * delta * delta + E / d2;
    if (g_o != nullptr) {
      const float k = weight / ((float)B * (float)J) * gscale[0];
      for (int b = 0; b < B; ++b) {
        const float lp = l[(size_t)b * J + j] - ml, e = lp - (o[(size_t)b * J + j] - mo);
        g_o[(size_t)b * J + j] = k * (-2.f * delta - 2.f * e / d2);
        float gl = 2.f * delta + 2.f * e / d2;
        if (normed) gl -= 2.f * E * lp / ((float)B * d2 * d2);
        g_l[(size_t)b * J + j] = k * gl;
      }
    }
  }
  if (loss != nullptr) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) S += __shfl_xor_sync(0xffffffffu, S, q);
    if (lane == 0) sm[w] = S;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += sm[q];
      loss[0] = weight / ((float)B * (float)J) * t;
    }
  }
}
extern "C" int cn_norm_latent_loss_fwd(const float* out, const float* labels, int b, int j, int nrot, float weight, float* loss, void* stream) {
  CN_REQUIRE(j >= 1 && j <= 1024 && b >= 1, CN_ERR_BAD_SHAPE, "cn_norm_latent_loss: 1 <= columns <= 1024");
  norm_latent_loss_kernel<<<1, (j + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(out, labels, b, j, nrot, weight, nullptr, loss, nullptr, nullptr);
  CN_CHECK_LAUNCH(); return CN_OK;
}
extern "C" int cn_norm_latent_loss_bwd(const float* out, const float* labels, int b, int j, int nrot, float weight, const float* gscale,
                                       float* g_out, float* g_labels, void* stream) {
  CN_REQUIRE(j >= 1 && j <= 1024 && b >= 1, CN_ERR_BAD_SHAPE, "cn_norm_latent_loss: 1 <= columns <= 1024");
  norm_latent_loss_kernel<<<1, (j + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(out, labels, b, j, nrot, weight, gscale, nullptr, g_out, g_labels);
  CN_CHECK_LAUNCH(); return CN_OK;
}
