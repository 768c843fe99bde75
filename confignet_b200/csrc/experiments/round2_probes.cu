// Round-2 probes for the tcgen05 convolution kernel.  NOT part of
// libconfignet_b200.so: built into confignet_b200/lib/libcn_probes.so by scripts/build_probes.sh and driven by
// scripts/gpu_probe_round2.py.  Three questions that decide the next kernel design, each answered by a tiny
// single-purpose kernel that is checked against a host model:
//
//   1. probe_tf32_operands   what does kind::tf32 do with the 13 low mantissa bits of a raw fp32 operand -
//                            ignore them (truncate) or round?  If it truncates, the two "big" products of the
//                            3xTF32 split can read the RAW activation tile and only a_small has to be built.
//   2. probe_tma_tile        a (C, W, H, N) tiled tensor map with box (32, bw*s, bh*s, 1), element strides
//                            (1, s, s, 1), SWIZZLE_128B and negative / overhanging start coordinates: does the
//                            zero fill give exactly the SAME-padded im2col rows of one (tap, 32-channel) k-block,
//                            in the K-major swizzled layout the UMMA descriptor reads (row = pixel, 128 B per row,
//                            16-byte chunk index XOR (row & 7), 8-row groups 1024 B apart)?
//   3. probe_conv_tma        both together: a 3x3 SAME convolution (stride 1 or 2, Cout = 16) whose A operand is
//                            fetched ONLY by TMA (no per-thread gathers) and fed to the MMA straight from shared
//                            memory (SS form, one tf32 product - integer test data keep it exact).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 4-D tiled TMA load: coordinates are (c, x, y, n) in ELEMENTS of the tensor map's dimensions, signed
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(n), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B operand tile: rows of 128 B (32 tf32 of K), 8-row groups 1024 B apart (layout type 2)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3ffffu) >> 4);
  d |= (uint64_t)((16u >> 4) & 0x3fffu) << 16;      // LBO (unused for swizzled K-major)
  d |= (uint64_t)((1024u >> 4) & 0x3fffu) << 32;    // SBO
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
__host__ __device__ inline uint32_t umma_idesc_tf32(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format f32
  d |= 2u << 7;                      // A format tf32
  d |= 2u << 10;                     // B format tf32 (both K-major)
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}
__host__ __device__ inline uint32_t swz_off(int row, int k) {      // byte offset of element (row, k) in a K-major SWIZZLE_128B tile
  return (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + ((uint32_t)((k >> 2) ^ (row & 7)) << 4) + (uint32_t)(k & 3) * 4u;
}

constexpr int PN = 16;                 // GEMM N of the probes (output channels)
constexpr int A_BYTES = 128 * 128;     // one activation stage: 128 pixels x 32 channels fp32
constexpr int B_BYTES = PN * 128;      // one weight stage: 16 output channels x 32 channels fp32

struct __align__(1024) ProbeSmem {
  unsigned char a[A_BYTES];
  unsigned char b[B_BYTES];
  unsigned long long bar_full, bar_mma;
  uint32_t tmem;
};

__device__ __forceinline__ uint32_t tmem_alloc32(uint32_t holder_addr, const uint32_t* holder, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_addr), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *(volatile const uint32_t*)holder;
}
__device__ __forceinline__ void tmem_free32(uint32_t base, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32) : "memory");
}

// ---- 1. raw fp32 operands through kind::tf32: D[128][16] = A[128][32] * B[16][32]^T, one k-block
__global__ void __launch_bounds__(128) tf32_operand_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ unsigned char raw[];
  ProbeSmem* s = reinterpret_cast<ProbeSmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int k = 0; k < 32; ++k) *reinterpret_cast<float*>(s->a + swz_off(tid, k)) = A[tid * 32 + k];
  if (tid < PN) for (int k = 0; k < 32; ++k) *reinterpret_cast<float*>(s->b + swz_off(tid, k)) = B[tid * 32 + k];
  if (tid == 0) mbar_init(smem_u32(&s->bar_mma), 1);
  fence_proxy_async();                                   // generic-proxy writes -> visible to the tensor core's async proxy
  const uint32_t tmem = tmem_alloc32(smem_u32(&s->tmem), &s->tmem, warp);
  if (tid == 0) {
    const uint64_t ad = umma_desc_k128(smem_u32(s->a)), bd = umma_desc_k128(smem_u32(s->b));
    for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ss(tmem, ad + kk * 2, bd + kk * 2, umma_idesc_tf32(PN), kk > 0);   // +32 B per k-step
    tc_commit(smem_u32(&s->bar_mma));
  }
  mbar_wait(smem_u32(&s->bar_mma), 0);
  tc_fence_after();
  uint32_t v[16];
  tc_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < PN; ++j) D[tid * PN + j] = __uint_as_float(v[j]);
  tmem_free32(tmem, warp);
}

// ---- 2. one TMA tile, dumped raw (16 KB, swizzled as it lies in shared memory)
__global__ void __launch_bounds__(128) tma_tile_kernel(const __grid_constant__ CUtensorMap map, int c, int x, int y, int n, float* __restrict__ out) {
  extern __shared__ unsigned char raw[];
  ProbeSmem* s = reinterpret_cast<ProbeSmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x;
  for (int i = tid; i < A_BYTES / 4; i += 128) reinterpret_cast<float*>(s->a)[i] = -12345.0f;     // so untouched bytes show
  if (tid == 0) mbar_init(smem_u32(&s->bar_full), 1);
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(smem_u32(&s->bar_full), A_BYTES);
    tma_load_4d(smem_u32(s->a), &map, c, x, y, n, smem_u32(&s->bar_full));
  }
  mbar_wait(smem_u32(&s->bar_full), 0);
  for (int i = tid; i < A_BYTES / 4; i += 128) out[i] = reinterpret_cast<float*>(s->a)[i];
}

// ---- 3. 3x3 SAME convolution, A by TMA only, SS-form MMA.  One CTA per tile of 128 output pixels (bw x bh),
//         k-blocks = 9 taps x C/32, strictly sequential (load -> MMA -> next): a correctness probe, not a fast kernel.
//         wp: per k-block one 2 KB weight stage already in the swizzled K-major layout (host-packed).
__global__ void __launch_bounds__(128) conv_tma_kernel(const __grid_constant__ CUtensorMap map, const float* __restrict__ wp,
                                                       float* __restrict__ y, int Ho, int Wo, int C, int stride, int pad, int bw, int bh) {
  extern __shared__ unsigned char raw[];
  ProbeSmem* s = reinterpret_cast<ProbeSmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tiles_x = Wo / bw, tiles_y = Ho / bh;
  const int tile = blockIdx.x, n = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
  const int x0 = tx * bw, y0 = ty * bh;
  if (tid == 0) { mbar_init(smem_u32(&s->bar_full), 1); mbar_init(smem_u32(&s->bar_mma), 1); }
  fence_proxy_async();
  const uint32_t tmem = tmem_alloc32(smem_u32(&s->tmem), &s->tmem, warp);
  const int cblocks = C / 32, num_kb = 9 * cblocks;
  for (int kb = 0; kb < num_kb; ++kb) {
    const int tap = kb / cblocks, cb = kb % cblocks, dy = tap / 3, dx = tap % 3;
    if (tid == 0) {
      mbar_arrive_expect_tx(smem_u32(&s->bar_full), A_BYTES + B_BYTES);
      tma_load_4d(smem_u32(s->a), &map, cb * 32, x0 * stride + dx - pad, y0 * stride + dy - pad, n, smem_u32(&s->bar_full));
      bulk_g2s(smem_u32(s->b), wp + (size_t)kb * (B_BYTES / 4), B_BYTES, smem_u32(&s->bar_full));
    }
    mbar_wait(smem_u32(&s->bar_full), kb & 1);
    tc_fence_after();
    if (tid == 0) {
      const uint64_t ad = umma_desc_k128(smem_u32(s->a)), bd = umma_desc_k128(smem_u32(s->b));
      for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ss(tmem, ad + kk * 2, bd + kk * 2, umma_idesc_tf32(PN), (kb | kk) != 0);
      tc_commit(smem_u32(&s->bar_mma));
    }
    mbar_wait(smem_u32(&s->bar_mma), kb & 1);          // the stage is free again once the MMAs that read it have retired
    tc_fence_after();
  }
  uint32_t v[16];
  tc_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  const int px = x0 + tid % bw, py = y0 + tid / bw;      // GEMM row = pixel of the box, x fastest (the TMA's order)
  float* dst = y + (((size_t)n * Ho + py) * Wo + px) * PN;
  for (int j = 0; j < PN; ++j) dst[j] = __uint_as_float(v[j]);
  tmem_free32(tmem, warp);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* map, const float* x, int N, int H, int W, int C, int bw, int bh, int stride) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return -1;
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), 1};
  const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 100;
}

template <typename K>
int prep(K kernel) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProbeSmem) + 1024) == cudaSuccess ? 0 : -2;
}
int finish() {
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "probe: %s\n", cudaGetErrorString(e)); return -3; }
  return 0;
}

}  // namespace

// All pointers are DEVICE pointers.  Return 0 or a negative code (-1 no driver entry point, -2 launch set-up, -3 kernel
// error, <= -100: -(CUresult) - 100 from cuTensorMapEncodeTiled).
extern "C" int probe_tf32_operands(const float* A, const float* B, float* D) {
  if (prep(tf32_operand_kernel)) return -2;
  tf32_operand_kernel<<<1, 128, sizeof(ProbeSmem) + 1024>>>(A, B, D);
  return finish();
}

extern "C" int probe_tma_tile(const float* x, int N, int H, int W, int C, int bw, int bh, int stride, int c, int xs, int ys, int n, float* out) {
  if (bw * bh != 128) return -2;
  CUtensorMap map;
  const int r = make_map(&map, x, N, H, W, C, bw, bh, stride);
  if (r) return r;
  if (prep(tma_tile_kernel)) return -2;
  tma_tile_kernel<<<1, 128, sizeof(ProbeSmem) + 1024>>>(map, c, xs, ys, n, out);
  return finish();
}

extern "C" int probe_conv_tma(const float* x, const float* wp, float* y, int N, int H, int W, int C, int stride) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const int bw = Wo < 128 ? Wo : 128, bh = 128 / bw;
  if (C % 32 || 128 % bw || Wo % bw || Ho % bh) return -2;
  const int total = (Ho - 1) * stride + 3 - H, pad = (total > 0 ? total : 0) / 2;     // TF SAME: the smaller half in front
  CUtensorMap map;
  const int r = make_map(&map, x, N, H, W, C, bw, bh, stride);
  if (r) return r;
  if (prep(conv_tma_kernel)) return -2;
  conv_tma_kernel<<<N * (Ho / bh) * (Wo / bw), 128, sizeof(ProbeSmem) + 1024>>>(map, wp, y, Ho, Wo, C, stride, pad, bw, bh);
  return finish();
}
