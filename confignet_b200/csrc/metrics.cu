// Training-time metric networks: the layers of keras-applications InceptionV3
// (metrics/inception_distance.py:9-27) and MobileNetV2 (metrics/celeba_attribute_prediction.py:54-62,128-141) that are not
// convolutions of the implicit-GEMM family - window pooling (max, TF-SAME average), the depthwise 3x3 convolution with its
// folded BatchNorm + ReLU6, and cv2.resize's bilinear interpolation.  Forward only (the metric networks are never trained
// on this path), HBM-bound: a thread owns four channels of one output pixel, every access is a 16-byte piece of a
// channels-last pixel record.
#include "common.cuh"

static inline int grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------ window pooling
// y[n, oy, ox, 0:c] (pixel records ldy floats apart: a branch of an Inception block writes its slice of the concatenated
// tensor) = max / mean over the in-bounds part of the kh x kw window at (oy*stride - pad_t, ox*stride - pad_l).
// Mean divides by the number of in-bounds elements: tf.nn.avg_pool with SAME padding leaves the padding out of the count.
template <int VEC>
__global__ void pool2d_fwd_kernel(const float* __restrict__ x, int h, int w, int c, int kh, int kw, int stride, int pad_t,
                                  int pad_l, int oh, int ow, int mode, float* __restrict__ y, int ldy, size_t total) {
  const int cv = c / VEC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cv) * VEC; size_t t = i / cv;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh); const size_t n = t / oh;
    const int y0 = oy * stride - pad_t, x0 = ox * stride - pad_l;
    float acc[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] = mode == 0 ? -3.402823466e38f : 0.f;
    int cnt = 0;
    for (int ky = 0; ky < kh; ++ky) {
      const int iy = y0 + ky;
      if ((unsigned)iy >= (unsigned)h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int ix = x0 + kx;
        if ((unsigned)ix >= (unsigned)w) continue;
        const float* src = x + ((n * h + iy) * w + ix) * c + ch;
        float v[VEC];
        if (VEC == 4) { const float4 f = *reinterpret_cast<const float4*>(src); v[0] = f.x; v[1 % VEC] = f.y; v[2 % VEC] = f.z; v[3 % VEC] = f.w; }
        else v[0] = *src;
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = mode == 0 ? fmaxf(acc[q], v[q]) : acc[q] + v[q];
        ++cnt;
      }
    }
    if (mode != 0) {
      const float d = (float)(cnt > 0 ? cnt : 1);
#pragma unroll
      for (int q = 0; q < VEC; ++q) acc[q] = acc[q] / d;
    }
    float* dst = y + ((n * oh + oy) * ow + ox) * (size_t)ldy + ch;
    if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
    else dst[0] = acc[0];
  }
}

// 3x3 windows, four channels x a strip of TX output pixels per thread: a stride-1 strip of 4 reads 6 columns per row
// instead of 12 (4.5 loads per output instead of 9), every row's loads are issued together.  The one-pixel-per-thread form
// above ran the stride-1 average pool of InceptionV3 at 0.24 of the HBM rate (profiles/r02_m6_metric_kernels_hbm.txt).
template <int STRIDE, int MODE, int TX>
__global__ void __launch_bounds__(256)
pool3x3_strip_kernel(const float* __restrict__ x, int h, int w, int c, int pad_t, int pad_l, int oh, int ow,
                     float* __restrict__ y, int ldy, size_t total) {
  constexpr int NC = (TX - 1) * STRIDE + 3;
  const int c4 = c >> 2, nxs = (ow + TX - 1) / TX;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c4) * 4; size_t t = i / c4;
    const int ox0 = (int)(t % nxs) * TX; t /= nxs;
    const int oy = (int)(t % oh); const size_t n = t / oh;
    const int y0 = oy * STRIDE - pad_t, x0 = ox0 * STRIDE - pad_l;
    float4 acc[TX];
#pragma unroll
    for (int j = 0; j < TX; ++j) acc[j] = MODE == CN_POOL_MAX ? make_float4(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f)
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
    int rows = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y0 + ky;
      if ((unsigned)iy >= (unsigned)h) continue;
      ++rows;
      const float* row = x + ((n * h + iy) * w) * (size_t)c + ch;
      float4 v[NC];
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        const int ix = x0 + q;
        v[q] = (unsigned)ix < (unsigned)w ? cn_ldg4_ordered(row + (size_t)ix * c)
                                          : (MODE == CN_POOL_MAX ? make_float4(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f));
      }
#pragma unroll
      for (int j = 0; j < TX; ++j)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 f = v[j * STRIDE + kx];
          if (MODE == CN_POOL_MAX) { acc[j].x = fmaxf(acc[j].x, f.x); acc[j].y = fmaxf(acc[j].y, f.y); acc[j].z = fmaxf(acc[j].z, f.z); acc[j].w = fmaxf(acc[j].w, f.w); }
          else { acc[j].x += f.x; acc[j].y += f.y; acc[j].z += f.z; acc[j].w += f.w; }
        }
    }
#pragma unroll
    for (int j = 0; j < TX; ++j) {
      const int ox = ox0 + j;
      if (ox >= ow) break;
      float4 o = acc[j];
      if (MODE != CN_POOL_MAX) {
        int cols = 0;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) cols += (unsigned)(ox * STRIDE - pad_l + kx) < (unsigned)w ? 1 : 0;
        const float d = (float)(rows * cols > 0 ? rows * cols : 1);
        o.x = o.x / d; o.y = o.y / d; o.z = o.z / d; o.w = o.w / d;
      }
      *reinterpret_cast<float4*>(y + ((n * oh + oy) * ow + ox) * (size_t)ldy + ch) = o;
    }
  }
}

extern "C" int cn_pool2d_fwd(const float* x, int n, int h, int w, int c, int kh, int kw, int stride, int pad_t, int pad_l,
                             int oh, int ow, int mode, float* y, int ldy, void* stream) {
  CN_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && kh > 0 && kw > 0 && stride > 0 && oh > 0 && ow > 0 && ldy >= c,
             CN_ERR_BAD_SHAPE, "cn_pool2d_fwd: bad shape n=%d h=%d w=%d c=%d k=%dx%d stride=%d out=%dx%d ldy=%d", n, h, w, c, kh,
             kw, stride, oh, ow, ldy);
  CN_REQUIRE(mode == CN_POOL_MAX || mode == CN_POOL_AVG_VALID, CN_ERR_UNSUPPORTED, "cn_pool2d_fwd: mode %d", mode);
  // every window must hold at least one in-bounds element (true for VALID and SAME geometries)
  CN_REQUIRE((oh - 1) * stride - pad_t < h && (ow - 1) * stride - pad_l < w && pad_t < kh && pad_l < kw, CN_ERR_BAD_SHAPE,
             "cn_pool2d_fwd: a window of the %dx%d output lies outside the %dx%d input", oh, ow, h, w);
  if (n == 0) return CN_OK;
  const bool vec = c % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0;
  if (vec && kh == 3 && kw == 3 && (stride == 1 || stride == 2)) {
    constexpr int TX = 4;
    const size_t items = (size_t)n * oh * ((ow + TX - 1) / TX) * (c / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1 && mode == CN_POOL_MAX) pool3x3_strip_kernel<1, CN_POOL_MAX, TX><<<grid_for(items), 256, 0, st>>>(x, h, w, c, pad_t, pad_l, oh, ow, y, ldy, items);
    else if (stride == 1) pool3x3_strip_kernel<1, CN_POOL_AVG_VALID, TX><<<grid_for(items), 256, 0, st>>>(x, h, w, c, pad_t, pad_l, oh, ow, y, ldy, items);
    else if (mode == CN_POOL_MAX) pool3x3_strip_kernel<2, CN_POOL_MAX, TX><<<grid_for(items), 256, 0, st>>>(x, h, w, c, pad_t, pad_l, oh, ow, y, ldy, items);
    else pool3x3_strip_kernel<2, CN_POOL_AVG_VALID, TX><<<grid_for(items), 256, 0, st>>>(x, h, w, c, pad_t, pad_l, oh, ow, y, ldy, items);
    CN_CHECK_LAUNCH(); return CN_OK;
  }
  const size_t total = (size_t)n * oh * ow * (vec ? c / 4 : c);
  if (vec)
    pool2d_fwd_kernel<4><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, h, w, c, kh, kw, stride, pad_t, pad_l, oh, ow, mode, y, ldy, total);
  else
    pool2d_fwd_kernel<1><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, h, w, c, kh, kw, stride, pad_t, pad_l, oh, ow, mode, y, ldy, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3 convolution
// keras DepthwiseConv2D(3, strides=s, use_bias=False) + BatchNormalization + ReLU(6) of a MobileNetV2 inverted residual
// block: y[n,oy,ox,ch] = act(sum_{ky,kx} x[n, oy*s - pt + ky, ox*s - pl + kx, ch] * wk[ky][kx][ch] + bias[ch]), TF SAME padding
// (stride 2 behind keras-applications' correct_pad + VALID is the same geometry, see netspec.mobilenet_v2_spec).  wk is the
// Keras (3,3,C,1) kernel with the BatchNorm scale folded in, bias the folded shift.
template <int VEC>
__global__ void dwconv3x3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ bias,
                                     int h, int w, int c, int stride, int pad_t, int pad_l, int oh, int ow, int act, float alpha,
                                     float* __restrict__ y, size_t total) {
  const int cv = c / VEC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cv) * VEC; size_t t = i / cv;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh); const size_t n = t / oh;
    const int y0 = oy * stride - pad_t, x0 = ox * stride - pad_l;
    float acc[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y0 + ky;
      if ((unsigned)iy >= (unsigned)h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x0 + kx;
        if ((unsigned)ix >= (unsigned)w) continue;
        const float* src = x + ((n * h + iy) * w + ix) * c + ch;
        const float* wp = wk + (ky * 3 + kx) * c + ch;
        if (VEC == 4) {
          const float4 f = *reinterpret_cast<const float4*>(src), g = *reinterpret_cast<const float4*>(wp);
          acc[0] = fmaf(f.x, g.x, acc[0]); acc[1 % VEC] = fmaf(f.y, g.y, acc[1 % VEC]);
          acc[2 % VEC] = fmaf(f.z, g.z, acc[2 % VEC]); acc[3 % VEC] = fmaf(f.w, g.w, acc[3 % VEC]);
        } else {
          acc[0] = fmaf(*src, *wp, acc[0]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] = cn_apply_act_ext(acc[q] + (bias ? bias[ch + q] : 0.f), act, alpha);
    float* dst = y + ((n * oh + oy) * ow + ox) * (size_t)c + ch;
    if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
    else dst[0] = acc[0];
  }
}

// the same strip form for the depthwise convolution: four channels x TX output pixels per thread, the nine kernel quads
// in registers, one row of (TX - 1) * STRIDE + 3 input quads in flight at a time
template <int STRIDE, int TX>
__global__ void __launch_bounds__(256)
dwconv3x3_strip_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ bias, int h, int w, int c,
                       int pad_t, int pad_l, int oh, int ow, int act, float alpha, float* __restrict__ y, size_t total) {
  constexpr int NC = (TX - 1) * STRIDE + 3;
  const int c4 = c >> 2, nxs = (ow + TX - 1) / TX;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c4) * 4; size_t t = i / c4;
    const int ox0 = (int)(t % nxs) * TX; t /= nxs;
    const int oy = (int)(t % oh); const size_t n = t / oh;
    const int y0 = oy * STRIDE - pad_t, x0 = ox0 * STRIDE - pad_l;
    float4 acc[TX];
#pragma unroll
    for (int j = 0; j < TX; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y0 + ky;
      if ((unsigned)iy >= (unsigned)h) continue;
      const float* row = x + ((n * h + iy) * w) * (size_t)c + ch;
      float4 v[NC];
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        const int ix = x0 + q;
        v[q] = (unsigned)ix < (unsigned)w ? cn_ldg4_ordered(row + (size_t)ix * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 g = *reinterpret_cast<const float4*>(wk + (ky * 3 + kx) * c + ch);
#pragma unroll
        for (int j = 0; j < TX; ++j) {
          const float4 f = v[j * STRIDE + kx];
          acc[j].x = fmaf(f.x, g.x, acc[j].x); acc[j].y = fmaf(f.y, g.y, acc[j].y);
          acc[j].z = fmaf(f.z, g.z, acc[j].z); acc[j].w = fmaf(f.w, g.w, acc[j].w);
        }
      }
    }
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b = *reinterpret_cast<const float4*>(bias + ch);
#pragma unroll
    for (int j = 0; j < TX; ++j) {
      const int ox = ox0 + j;
      if (ox >= ow) break;
      float4 o;
      o.x = cn_apply_act_ext(acc[j].x + b.x, act, alpha); o.y = cn_apply_act_ext(acc[j].y + b.y, act, alpha);
      o.z = cn_apply_act_ext(acc[j].z + b.z, act, alpha); o.w = cn_apply_act_ext(acc[j].w + b.w, act, alpha);
      *reinterpret_cast<float4*>(y + ((n * oh + oy) * ow + ox) * (size_t)c + ch) = o;
    }
  }
}

extern "C" int cn_dwconv3x3_fwd(const float* x, const float* wk, const float* bias, int n, int h, int w, int c, int stride,
                                int act, float alpha, float* y, void* stream) {
  CN_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && (stride == 1 || stride == 2), CN_ERR_BAD_SHAPE,
             "cn_dwconv3x3_fwd: bad shape n=%d h=%d w=%d c=%d stride=%d", n, h, w, c, stride);
  if (n == 0) return CN_OK;
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;
  const int tot_h = (oh - 1) * stride + 3 - h, tot_w = (ow - 1) * stride + 3 - w;       // TF SAME: the smaller half in front
  const int pad_t = (tot_h > 0 ? tot_h : 0) / 2, pad_l = (tot_w > 0 ? tot_w : 0) / 2;
  const bool vec = c % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)wk % 16) == 0;
  if (vec && (bias == nullptr || ((uintptr_t)bias % 16) == 0)) {
    constexpr int TX = 4;
    const size_t items = (size_t)n * oh * ((ow + TX - 1) / TX) * (c / 4);
    if (stride == 1) dwconv3x3_strip_kernel<1, TX><<<grid_for(items), 256, 0, (cudaStream_t)stream>>>(x, wk, bias, h, w, c, pad_t, pad_l, oh, ow, act, alpha, y, items);
    else dwconv3x3_strip_kernel<2, TX><<<grid_for(items), 256, 0, (cudaStream_t)stream>>>(x, wk, bias, h, w, c, pad_t, pad_l, oh, ow, act, alpha, y, items);
    CN_CHECK_LAUNCH(); return CN_OK;
  }
  const size_t total = (size_t)n * oh * ow * (vec ? c / 4 : c);
  if (vec)
    dwconv3x3_fwd_kernel<4><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, wk, bias, h, w, c, stride, pad_t, pad_l, oh, ow, act, alpha, y, total);
  else
    dwconv3x3_fwd_kernel<1><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, wk, bias, h, w, c, stride, pad_t, pad_l, oh, ow, act, alpha, y, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// ------------------------------------------------------------------------------------------------ cv2.resize, INTER_LINEAR
// predict_attributes (metrics/celeba_attribute_prediction.py:131-136) resizes every image to the classifier's input
// shape with cv2.resize's default interpolation.  Source position of destination pixel d: f = (d + 0.5) * (src / dst) - 0.5
// in float, s = floor(f), clamped at both borders with the fraction set to 0 (OpenCV resize.cpp, the coordinate tables of
// the INTER_LINEAR branch).  uint8 images go through OpenCV's fixed-point form: weights rounded to 1/2048 (shorts), the
// horizontal pass in int32, the vertical pass ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
struct ResizeAxis { int s0, s1; float f; };
__device__ __forceinline__ ResizeAxis resize_axis(int d, int src, double scale) {
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= src - 1) { f = 0.f; s = src - 1; }
  ResizeAxis a; a.s0 = s; a.s1 = s + 1 < src ? s + 1 : src - 1; a.f = f;
  return a;
}
__device__ __forceinline__ int resize_coef(float v) {      // saturate_cast<short>(v * 2048): round half to even
  int r = __float2int_rn(v * 2048.f);
  return r > 32767 ? 32767 : (r < -32768 ? -32768 : r);
}
__global__ void resize_u8_kernel(const unsigned char* __restrict__ x, int h, int w, int c, int oh, int ow, double sy, double sx,
                                 unsigned char* __restrict__ y, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c); size_t t = i / c;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh); const size_t n = t / oh;
    const ResizeAxis ax = resize_axis(ox, w, sx), ay = resize_axis(oy, h, sy);
    const int a0 = resize_coef(1.f - ax.f), a1 = resize_coef(ax.f), b0 = resize_coef(1.f - ay.f), b1 = resize_coef(ay.f);
    const unsigned char* r0 = x + ((n * h + ay.s0) * w) * (size_t)c + ch;
    const unsigned char* r1 = x + ((n * h + ay.s1) * w) * (size_t)c + ch;
    const int h0 = (int)r0[(size_t)ax.s0 * c] * a0 + (int)r0[(size_t)ax.s1 * c] * a1;
    const int h1 = (int)r1[(size_t)ax.s0 * c] * a0 + (int)r1[(size_t)ax.s1 * c] * a1;
    int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    y[i] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}
__global__ void resize_f32_kernel(const float* __restrict__ x, int h, int w, int c, int oh, int ow, double sy, double sx,
                                  float* __restrict__ y, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c); size_t t = i / c;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh); const size_t n = t / oh;
    const ResizeAxis ax = resize_axis(ox, w, sx), ay = resize_axis(oy, h, sy);
    const float a0 = 1.f - ax.f, a1 = ax.f, b0 = 1.f - ay.f, b1 = ay.f;
    const float* r0 = x + ((n * h + ay.s0) * w) * (size_t)c + ch;
    const float* r1 = x + ((n * h + ay.s1) * w) * (size_t)c + ch;
    // no contraction: OpenCV's two passes round each product and sum separately
    const float h0 = __fadd_rn(__fmul_rn(r0[(size_t)ax.s0 * c], a0), __fmul_rn(r0[(size_t)ax.s1 * c], a1));
    const float h1 = __fadd_rn(__fmul_rn(r1[(size_t)ax.s0 * c], a0), __fmul_rn(r1[(size_t)ax.s1 * c], a1));
    y[i] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
  }
}
extern "C" int cn_resize_bilinear(const void* x, int n, int h, int w, int c, int oh, int ow, int is_u8, void* y, void* stream) {
  CN_REQUIRE(n >= 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, CN_ERR_BAD_SHAPE,
             "cn_resize_bilinear: bad shape n=%d h=%d w=%d c=%d out=%dx%d", n, h, w, c, oh, ow);
  if (n == 0) return CN_OK;
  const size_t total = (size_t)n * oh * ow * c;
  const double sy = 1.0 / ((double)oh / (double)h), sx = 1.0 / ((double)ow / (double)w);   // OpenCV: scale = 1 / inv_scale
  if (is_u8)
    resize_u8_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const unsigned char*)x, h, w, c, oh, ow, sy, sx, (unsigned char*)y, total);
  else
    resize_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const float*)x, h, w, c, oh, ow, sy, sx, (float*)y, total);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// uint8 -> float32 without the /127.5 - 1 map (cn_from_uint8 applies it): predict_attributes feeds uint8 images straight
// to mobilenet_v2.preprocess_input, which casts then divides; the cast alone is needed ahead of the float resize.
__global__ void u8_to_f32_kernel(const unsigned char* __restrict__ x, float* __restrict__ y, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) y[i] = (float)x[i];
}
extern "C" int cn_u8_to_f32(const void* x, float* y, int64_t n, void* stream) {
  if (n <= 0) return CN_OK;
  u8_to_f32_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>((const unsigned char*)x, y, (size_t)n);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// The two float maps around the classifier / Inception inputs, with the reference's operation order (one rounding per
// NumPy operation): mode 0: (x + 1) * 127.5 (celeba_attribute_prediction.py:129-130); mode 1: x / 127.5 - 1
// (keras-applications preprocess_input, mode "tf": inception_distance.py:24, celeba_attribute_prediction.py:138).
__global__ void pixel_map_kernel(const float* __restrict__ x, float* __restrict__ y, size_t total, int mode) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = mode == 0 ? __fmul_rn(__fadd_rn(v, 1.f), 127.5f) : __fadd_rn(__fdiv_rn(v, 127.5f), -1.f);
  }
}
extern "C" int cn_pixel_map(const float* x, float* y, int64_t n, int mode, void* stream) {
  CN_REQUIRE(mode == 0 || mode == 1, CN_ERR_UNSUPPORTED, "cn_pixel_map: mode %d", mode);
  if (n <= 0) return CN_OK;
  pixel_map_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>(x, y, (size_t)n, mode);
  CN_CHECK_LAUNCH(); return CN_OK;
}

// y = act(x) for the two activation codes the conv epilogues do not carry (CN_ACT_RELU6 after a conv that ran with
// CN_ACT_RELU is the clamp at 6; CN_ACT_SIGMOID on the attribute head's logits), in place or out of place.
__global__ void act_ext_kernel(const float* __restrict__ x, float* __restrict__ y, size_t total, int act) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    y[i] = cn_apply_act_ext(x[i], act, 0.f);
}
__global__ void act_ext4_kernel(const float4* __restrict__ x, float4* __restrict__ y, size_t total4, int act) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    v.x = cn_apply_act_ext(v.x, act, 0.f); v.y = cn_apply_act_ext(v.y, act, 0.f);
    v.z = cn_apply_act_ext(v.z, act, 0.f); v.w = cn_apply_act_ext(v.w, act, 0.f);
    y[i] = v;
  }
}
extern "C" int cn_act_ext(const float* x, float* y, int64_t n, int act, void* stream) {
  CN_REQUIRE(act == CN_ACT_RELU6 || act == CN_ACT_SIGMOID, CN_ERR_UNSUPPORTED, "cn_act_ext: activation code %d", act);
  if (n <= 0) return CN_OK;
  if (n % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0)
    act_ext4_kernel<<<grid_for((size_t)n / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)y, (size_t)n / 4, act);
  else
    act_ext_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>(x, y, (size_t)n, act);
  CN_CHECK_LAUNCH(); return CN_OK;
}
