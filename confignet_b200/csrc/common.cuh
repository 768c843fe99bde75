// Shared declarations for libconfignet_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/confignet_b200.h"

void cn_set_error(const char* fmt, ...);
extern unsigned long long g_cn_weight_epoch;  // see error.cu: global part of every parameter buffer's epoch (cn_weights_changed)
void cn_mark_params_changed(const void* p);   // conv.cu: the registered buffer containing p was written
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

// Internal helpers shared by the translation units (defined in conv.cu).
//   cn_scratch      grow-only device scratch slot identified by `key` (any stable address); launches on one stream serialise
//   cn_sum_slabs    out[i] = sum_b part[b][i] over `nslabs` slabs of n floats, in slab order (deterministic, no atomics)
int cn_scratch(const void* key, size_t bytes, float** out);
int cn_sum_slabs(const float* part, int nslabs, int n, float* out, cudaStream_t st);

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
  // Phased plans (nphase > 1): several GEMMs that share everything but their tap list and output offset run
  // in ONE launch, blockIdx.y = phase * n_tiles + n_tile.  Used by the sub-pixel phases of the folded
  // upsample+conv forward and by the parity phases of the stride-2 input gradient.
  int nphase;
  int kb_stride;    // k-blocks per (phase, n-tile) slot of the packed-weight buffer (max over phases)
  int taps_total;   // entries of the tap table (all phases)
  int ph_ntaps[8];  // taps of phase ph are taps[ph_tap0[ph] .. ph_tap0[ph] + ph_ntaps[ph])
  int ph_tap0[8];
  int ph_ooff[8];   // ooff bits of the phase: bit d = ooff[d]
};

// Select the phase of a phased plan from blockIdx.y; returns the n-tile index.  (Select chains instead of
// indexed reads: a dynamically indexed kernel parameter would be copied to local memory.)
__device__ __forceinline__ int cn_select_phase(GemmPlan& p, int by, int grid_y) {
  if (p.nphase <= 1) return by;
  const int ntn = grid_y / p.nphase;
  const int ph = by / ntn;
  int tap0 = 0, nt = 0, bits = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i == ph) { tap0 = p.ph_tap0[i]; nt = p.ph_ntaps[i]; bits = p.ph_ooff[i]; }
  p.taps += tap0;
  p.ntaps = nt;
  p.Ktot = nt * p.Csrc;
  p.ooff[0] = bits & 1; p.ooff[1] = (bits >> 1) & 1; p.ooff[2] = (bits >> 2) & 1;
  return by - ph * ntn;
}

// Read-only loads as volatile asm statements.  The compiler keeps volatile asm in program order, so a batch of these is
// issued back to back; plain C++ loads of a streaming loop get sunk next to their first use - ONE load in flight per thread,
// 0.3-0.5 of the HBM rate (SASS of the round-1 statistics / column-sum kernels, profiles/r02_hbm_kernels_batched_loads.txt).
__device__ __forceinline__ float4 cn_ldg4_ordered(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float cn_ldg1_ordered(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float cn_apply_act(float v, int act, float alpha) {
  if (act == CN_ACT_LRELU) return v >= 0.f ? v : v * alpha;
  if (act == CN_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == CN_ACT_TANH) return tanhf(v);
  return v;
}
// The hot kernels take the four codes above only: the tcgen05 conv kernel is ~119 KB of SASS and its epilogue unrolls the
// activation 64 times - one more inlined branch there (ReLU6) cost 5 ms of the 39 ms training step on B200 (measured
// A/B, profiles/r02_m3_activation_code_size_ab.txt).  The metric networks' extra codes exist in the cold kernels only.
__device__ __forceinline__ float cn_apply_act_ext(float v, int act, float alpha) {
  if (act == CN_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  if (act == CN_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return cn_apply_act(v, act, alpha);
}
