// Shared declarations for libconfignet_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/confignet_b200.h"

void cn_set_error(const char* fmt, ...);
extern unsigned long long g_cn_launches;     // kernels launched by this library (bench.py reports it)

#define CN_CHECK_CUDA(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      cn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CN_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define CN_CHECK_LAUNCH()                                                              \
  do {                                                                                 \
    ++g_cn_launches;                                                                   \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      cn_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CN_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define CN_REQUIRE(cond, code, ...)                                                    \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      cn_set_error(__VA_ARGS__);                                                       \
      return (code);                                                                   \
    }                                                                                  \
  } while (0)

// Geometry of one implicit GEMM (see conv.cu).  Passed to kernels by value.
struct GemmPlan {
  int n_img;        // samples
  int E[3];         // extents of the row decode (m -> n_img, e0, e1, e2)
  int M;            // n_img*E0*E1*E2
  int mstride;      // u_d = e_d*mstride + off_d
  int U[3];         // validity bounds on u_d
  int ushift;       // source coordinate = u_d >> ushift
  int S[3];         // source spatial dims
  int Csrc;         // source channels
  int Q[3];         // destination spatial dims
  int ostride;      // q_d = e_d*ostride + ooff_d
  int ooff[3];
  int Cn;           // destination channels (GEMM N)
  int ntaps;        // generalized taps
  int Ktot;         // ntaps*Csrc
  int wsc, wsn;     // weight element strides for the source-channel and N index
  const int2* taps; // device: {packed (off_d+8) in 10-bit fields, weight base offset}
};

__device__ __forceinline__ float cn_apply_act(float v, int act, float alpha) {
  if (act == CN_ACT_LRELU) return v >= 0.f ? v : v * alpha;
  if (act == CN_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == CN_ACT_TANH) return tanhf(v);
  return v;
}
