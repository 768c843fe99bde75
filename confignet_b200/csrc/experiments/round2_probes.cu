// Round-2 probes for the tcgen05 convolution kernel.  NOT part of
// libconfignet_b200.so: built into confignet_b200/lib/libcn_probes.so by __graft_entry__.build() and driven by
// scripts/gpu_probe_round2.py.  Three questions that decide the next kernel design, each answered by a tiny
// single-purpose kernel that is checked against a host model, and the candidate kernel they lead to:
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
//   4. probe_conv_tma_fast   the candidate: persistent, pipelined, 3xTF32, TMA-fed A with only a_small built by threads,
//                            coalesced channel-major epilogue; timed against the production kernel on the same layer.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// makes the initialised barriers visible to the async proxy (TMA completions, tcgen05.commit arrivals)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
// 5-D tiled TMA load, coordinates (c, x, y, z, n): volumes, and images as volumes of depth 1
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int z, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(z), "r"(n), "r"(bar) : "memory");
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
  if (tid == 0) { mbar_init(smem_u32(&s->bar_mma), 1); mbar_init_fence(); }
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
  if (tid == 0) { mbar_init(smem_u32(&s->bar_full), 1); mbar_init_fence(); }
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
  if (tid == 0) { mbar_init(smem_u32(&s->bar_full), 1); mbar_init(smem_u32(&s->bar_mma), 1); mbar_init_fence(); }
  fence_proxy_async();
  const uint32_t tmem = tmem_alloc32(smem_u32(&s->tmem), &s->tmem, warp);
  const int cblocks = (C + 31) / 32, num_kb = 9 * cblocks;   // a partial last block: the TMA zero-fills the missing channels
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

// ---- 4. the candidate: pipelined, persistent, 3xTF32 forward 3x3 SAME convolution (stride 1 / 2, Cout = BN <= 128)
//         * A: raw fp32 tile by TMA (no per-thread gathers).  The two products that use a_big read the RAW tile from shared
//           memory (SS form; kind::tf32 is assumed to ignore the 13 low mantissa bits - probe 1 decides), only a_small is
//           built (4 warps: ld.shared of the swizzled rows, 3 ALU ops per element, tcgen05.st into a tensor-memory stage).
//         * B: pre-split, pre-swizzled stage images (big plane, small plane) by one bulk copy per k-block, as in production.
//         * two ping-pong accumulators, a chunk of CH k-blocks each, promoted into an fp32 running total in shared memory
//           ([column][129] floats: lane = row for the promotion, lane = column for the final pass - both conflict-free).
//         * epilogue: bias + LeakyReLU, lanes along the CHANNELS so that every store instruction writes one 128-byte line
//           (production stores one row per thread: 32 lines per instruction).
//         warps: 0 TMA producer, 1 MMA issue, 2-5 a_small builders, 6-9 promotion / epilogue.
constexpr int F_THREADS = 320;
constexpr int F_CH = 8;                    // k-blocks per accumulation chunk (production: TC_CHUNK_KB)
constexpr int F_TOT_LD = 129;              // leading dimension of the running total

// Phases of one launch: GEMMs over the same pixel grid that differ in tap list, weights and output offset (the parity phases
// of a stride-2 input gradient, the sub-pixel phases of a folded upsample + conv; a plain convolution is one phase).
// Taps of phase p: [tap0[p], tap0[p] + ntaps[p]); dx/dy/dz = source offset in source pixels / voxels (SAME padding folded in);
// (oz, oy, ox)[p] = offset of the phase's outputs on the ostride-spaced output grid.
struct PhaseList { int nphase; int ntaps[8], tap0[8]; short dx[64], dy[64], dz[64]; short oz[8], oy[8], ox[8]; };

struct FastCfg { int stages, stage_bytes, b_plane, tot_off, bar_off, tmem_off, smem_bytes, a_col0; };
__host__ __device__ inline FastCfg fast_cfg(int bn) {
  FastCfg c;
  c.b_plane = bn * 128;
  c.stage_bytes = A_BYTES + 2 * c.b_plane;
  const int tot_bytes = ((bn * F_TOT_LD * 4) + 1023) & ~1023;
  int st = (226 * 1024 - tot_bytes) / c.stage_bytes;      // 227 KB per CTA minus the alignment slack and the barriers
  if (st > 6) st = 6;
  while (2 * bn + st * 32 > 512) --st;
  c.stages = st;
  c.tot_off = st * c.stage_bytes;
  c.bar_off = c.tot_off + tot_bytes;
  c.tmem_off = c.bar_off + 8 * (3 * 6 + 4);
  c.smem_bytes = c.tmem_off + 16;
  c.a_col0 = 2 * bn;
  return c;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
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

__global__ void __launch_bounds__(F_THREADS) conv_tma_fast_kernel(const __grid_constant__ CUtensorMap map, const float* __restrict__ wp,
                                                                  const float* __restrict__ bias, float* __restrict__ y, int N, int Do, int Ho, int Wo,
                                                                  int C, int bn, int cout, int stride, const __grid_constant__ PhaseList P, int bw, int bh, int bd,
                                                                  float alpha, int od, int oh, int ow, int ostride) {
  // `map` is always 5-D (C, W, H, D, N); images are volumes of depth 1 (Do = bd = od = 1, dz = oz = 0).
  // GEMM rows = the (Do x) Ho x Wo grid of this launch; row (py, px) reads source pixel (py * stride + dy, px * stride + dx)
  // for tap (dx, dy) and writes output pixel (py * ostride + oy, px * ostride + ox) of an oh x ow image: a forward
  // convolution has ostride 1, a parity phase of a stride-2 input gradient / a sub-pixel phase of a folded upsample has 2.
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const FastCfg L = fast_cfg(bn);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = L.stages;
  const uint32_t bar_full = sbase + L.bar_off, bar_small = bar_full + 8 * 6, bar_empty = bar_small + 8 * 6,
                 bar_accfull = bar_empty + 8 * 6, bar_accempty = bar_accfull + 8 * 2;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_small + 8 * s, 128); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_accfull + 8 * b, 1); mbar_init(bar_accempty + 8 * b, 128); }
    mbar_init_fence();
  }
  fence_proxy_async();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + L.tmem_off), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.tmem_off);
  const int tiles_x = Wo / bw, tiles_y = Ho / bh, tiles_z = Do / bd, tiles_img = tiles_x * tiles_y * tiles_z;
  const int n_nt = cout / bn, nph = P.nphase, n_tiles = N * tiles_img * nph * n_nt;   // item = (pixel tile, phase, channel tile), channel tile innermost
  const int cblocks = (C + 31) / 32;                              // a partial last block: the TMA zero-fills the missing channels
  const uint32_t stage_tx = (uint32_t)L.stage_bytes;

  if (warp == 0) {
    // ===== producer: one TMA tile (A) + one bulk copy (B big + small planes) per k-block =====
    if (lane == 0) {
      int s = 0, ph = 0; long git = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
        const int nt = w % n_nt, ph_i = (w / n_nt) % nph, t = w / (n_nt * nph);
        const int n = t / tiles_img, tz = (t / (tiles_x * tiles_y)) % tiles_z, ty = (t / tiles_x) % tiles_y, tx = t % tiles_x;
        const int num_kb = P.ntaps[ph_i] * cblocks, tap_base = P.tap0[ph_i];
        // weight stages: phase after phase; inside a phase channel tile after channel tile, k-blocks tap-major
        const float* wsrc = wp + ((size_t)tap_base * cblocks * n_nt + (size_t)nt * num_kb) * (2 * L.b_plane / 4);
        for (int kb = 0; kb < num_kb; ++kb, ++git) {
          if (git >= S) mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const int tap = tap_base + kb / cblocks, cb = kb % cblocks;
          const uint32_t dst = sbase + s * L.stage_bytes;
          mbar_arrive_expect_tx(bar_full + 8 * s, stage_tx);
          tma_load_5d(dst, &map, cb * 32, tx * bw * stride + P.dx[tap], ty * bh * stride + P.dy[tap], tz * bd * stride + P.dz[tap], n,
                      bar_full + 8 * s);
          bulk_g2s(dst + A_BYTES, wsrc + (size_t)kb * (2 * L.b_plane / 4), 2 * L.b_plane, bar_full + 8 * s);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issue: per k-step (a_small [tmem] x b_big), (a_big x b_small), (a_big x b_big) =====
    const uint32_t idesc = umma_idesc_tf32(bn);
    int s = 0, ph = 0, b = 0, inchunk = 0, c = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int num_kb = P.ntaps[(t / n_nt) % nph] * cblocks;
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool chunk_first = inchunk == 0, chunk_last = inchunk == F_CH - 1 || kb == num_kb - 1;
        if (chunk_first && c >= 2) mbar_wait(bar_accempty + 8 * b, ((c >> 1) - 1) & 1);
        mbar_wait(bar_full + 8 * s, ph);
        mbar_wait(bar_small + 8 * s, ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = sbase + s * L.stage_bytes;
          const uint64_t ad = umma_desc_k128(a_addr), bb = umma_desc_k128(a_addr + A_BYTES), bs = umma_desc_k128(a_addr + A_BYTES + L.b_plane);
          const uint32_t d_t = tmem + b * bn, a_small = tmem + L.a_col0 + s * 32;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            tc_mma_tf32_ts(d_t, a_small + kk * 8, bb + kk * 2, idesc, !(chunk_first && kk == 0));
            tc_mma_tf32_ss(d_t, ad + kk * 2, bs + kk * 2, idesc, 1);
            tc_mma_tf32_ss(d_t, ad + kk * 2, bb + kk * 2, idesc, 1);
          }
          tc_commit(bar_empty + 8 * s);
          if (chunk_last) tc_commit(bar_accfull + 8 * b);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
        if (++inchunk == F_CH || kb == num_kb - 1) { inchunk = 0; ++c; b ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===== a_small builders: row = tensor-memory lane; the raw row is read from the swizzled stage =====
    const int row = (warp & 3) * 32 + lane;
    const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u, x7 = (uint32_t)(row & 7);
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    int s = 0, ph = 0; long git = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int num_kb = P.ntaps[(t / n_nt) % nph] * cblocks;
      for (int kb = 0; kb < num_kb; ++kb, ++git) {
        mbar_wait(bar_full + 8 * s, ph);
        if (git >= S) mbar_wait(bar_empty + 8 * s, ph ^ 1);          // the tensor-memory stage was read by the MMAs of its last use
        tc_fence_after();
        uint32_t v[32];
        const uint32_t a_addr = sbase + s * L.stage_bytes + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t q0, q1, q2, q3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(a_addr + ((j ^ x7) << 4)));
          const uint32_t q[4] = {q0, q1, q2, q3};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float xv = __uint_as_float(q[i]);
            v[4 * j + i] = __float_as_uint(xv - __uint_as_float(q[i] & 0xffffe000u)) & 0xffffe000u;
          }
        }
        tc_st32(tmem + lane_addr + L.a_col0 + s * 32, v);
        tc_fence_before();
        mbar_arrive(bar_small + 8 * s);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== promotion + epilogue =====
    const int q = warp & 3, row = q * 32 + lane;
    float* tot = reinterpret_cast<float*>(sm + L.tot_off);
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int b = 0, c = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int ph_i = (t / n_nt) % nph, num_kb = P.ntaps[ph_i] * cblocks;
      const int nchunks = (num_kb + F_CH - 1) / F_CH;
      for (int ch = 0; ch < nchunks; ++ch, ++c) {
        mbar_wait(bar_accfull + 8 * b, (c >> 1) & 1);
        tc_fence_after();
        for (int c0 = 0; c0 < bn; c0 += 32) {
          float* dst = tot + c0 * F_TOT_LD + row;
          if (bn - c0 >= 32) {
            uint32_t v[32];
            tc_ld32(tmem + lane_addr + b * bn + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * F_TOT_LD] = ch == 0 ? __uint_as_float(v[j]) : dst[j * F_TOT_LD] + __uint_as_float(v[j]);
          } else {
            uint32_t v[16];
            tc_ld16(tmem + lane_addr + b * bn + c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[j * F_TOT_LD] = ch == 0 ? __uint_as_float(v[j]) : dst[j * F_TOT_LD] + __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(bar_accempty + 8 * b);
        b ^= 1;
      }
      __syncwarp();                               // a warp finishes the 32 rows it promoted itself: no cross-warp dependency
      const int pt = t / (n_nt * nph), n0 = (t % n_nt) * bn;
      const int ox = P.ox[ph_i], oy = P.oy[ph_i], oz = P.oz[ph_i];
      const int n = pt / tiles_img, tz = (pt / (tiles_x * tiles_y)) % tiles_z, ty = (pt / tiles_x) % tiles_y, tx = pt % tiles_x;
      float bz[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) bz[i] = (bias != nullptr && lane + 32 * i < bn) ? bias[n0 + lane + 32 * i] : 0.f;
      for (int r = q * 32; r < q * 32 + 32; ++r) {
        const int px = (tx * bw + r % bw) * ostride + ox, py = (ty * bh + (r / bw) % bh) * ostride + oy, pz = (tz * bd + r / (bw * bh)) * ostride + oz;
        float* dst = y + ((((size_t)n * od + pz) * oh + py) * ow + px) * cout + n0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = lane + 32 * i;
          if (c0 < bn) {
            const float v = tot[c0 * F_TOT_LD + r] + bz[i];
            dst[c0] = v > 0.f ? v : alpha * v;
          }
        }
      }
      __syncwarp();                               // the rows may be overwritten by the next tile's first chunk
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
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

int make_map5(CUtensorMap* map, const float* x, int N, int D, int H, int W, int C, int bw, int bh, int bd, int stride) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return -1;
  const EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)D * H * W * C * 4};
  const cuuint32_t box[5] = {32, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)(bd * stride), 1};
  const cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), dims, strides, box, estr,
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

// The general entry.  x: source (N, D, H, W, C) (D = 1 for images); the launch covers a gd x gh x gw grid of GEMM pixels and
// nphase phases; tap t of phase p (taps [tap0, tap0 + ntaps[p]) in dx / dy / dz, phases one after the other) reads source voxel
// (pz * stride + dz[t], py * stride + dy[t], px * stride + dx[t]) (zero outside) and the phase's result goes to output voxel
// (pz * ostride + oz[p], py * ostride + oy[p], px * ostride + ox[p]) of y (N, od, oh, ow, cout).  alpha: LeakyReLU slope
// (1 = no activation); bias may be null.  geom = {D, H, W, gd, gh, gw, od, oh, ow}.
// wp: phase after phase; inside a phase per channel tile (bn = min(cout, 128) output channels), per k-block (tap-major, then
// 32-channel block) the big plane then the small plane, each bn rows x 128 B in the swizzled K-major layout.  Launches `iters`
// times (after one warm-up) on the default stream; *avg_us receives the mean launch time (CUDA events).
extern "C" int probe_conv_tma_phases(const float* x, int N, int C, const int* geom, const float* wp, const float* bias, float* y,
                                     int cout, int stride, int nphase, const int* ntaps, const int* dx, const int* dy, const int* dz,
                                     const int* oz, const int* oy, const int* ox, int ostride, float alpha, int iters, float* avg_us) {
  const int D = geom[0], H = geom[1], W = geom[2], gd = geom[3], gh = geom[4], gw = geom[5], od = geom[6], oh = geom[7], ow = geom[8];
  const int bn = cout <= 128 ? cout : 128;
  if (cout % bn || nphase < 1 || nphase > 8) return -2;
  const int bw = gw < 128 ? gw : 128, bh = gh < 128 / bw ? gh : 128 / bw, bd = 128 / (bw * bh);
  if (C % 4 || bn % 16 || 128 % bw || (128 / bw) % bh || gw % bw || gh % bh || gd % bd) return -2;      // C * 4 bytes: the tensor map's 16-byte stride rule
  PhaseList P;
  memset(&P, 0, sizeof(P));
  P.nphase = nphase;
  int total = 0;
  for (int p = 0; p < nphase; ++p) {
    if (ntaps[p] < 1) return -2;
    P.ntaps[p] = ntaps[p]; P.tap0[p] = total; total += ntaps[p];
    P.oz[p] = (short)(oz ? oz[p] : 0); P.oy[p] = (short)(oy ? oy[p] : 0); P.ox[p] = (short)(ox ? ox[p] : 0);
  }
  if (total > 64) return -2;
  for (int t = 0; t < total; ++t) { P.dx[t] = (short)dx[t]; P.dy[t] = (short)dy[t]; P.dz[t] = (short)(dz ? dz[t] : 0); }
  CUtensorMap map;
  const int r = make_map5(&map, x, N, D, H, W, C, bw, bh, bd, stride);
  if (r) return r;
  const FastCfg L = fast_cfg(bn);
  if (L.stages < 2) return -2;
  if (cudaFuncSetAttribute(conv_tma_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem_bytes + 1024) != cudaSuccess) return -2;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = N * (gd / bd) * (gh / bh) * (gw / bw) * nphase * (cout / bn), grid = n_tiles < sms ? n_tiles : sms;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < iters + 1; ++it) {
    if (it == 1) cudaEventRecord(e0);
    conv_tma_fast_kernel<<<grid, F_THREADS, L.smem_bytes + 1024>>>(map, wp, bias, y, N, gd, gh, gw, C, bn, cout, stride, P, bw, bh, bd, alpha,
                                                                   od, oh, ow, ostride);
  }
  cudaEventRecord(e1);
  const int f = finish();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (avg_us) *avg_us = iters > 0 ? ms * 1000.f / iters : 0.f;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return f;
}

// the candidate's shared / tensor memory plan for a channel tile of bn columns (host only; checked by the CPU suite):
// out = {stages, stage_bytes, b_plane, tot_off, bar_off, tmem_off, smem_bytes, a_col0}
extern "C" int probe_fast_cfg(int bn, int* out) {
  const FastCfg c = fast_cfg(bn);
  const int v[8] = {c.stages, c.stage_bytes, c.b_plane, c.tot_off, c.bar_off, c.tmem_off, c.smem_bytes, c.a_col0};
  for (int i = 0; i < 8; ++i) out[i] = v[i];
  return 0;
}

// forward 3x3 SAME convolution, stride 1 or 2, + bias + LeakyReLU(alpha): one phase of 9 taps
extern "C" int probe_conv_tma_fast(const float* x, const float* wp, const float* bias, float* y, int N, int H, int W, int C, int cout,
                                   int stride, float alpha, int iters, float* avg_us) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const int total = (Ho - 1) * stride + 3 - H, pad = (total > 0 ? total : 0) / 2;       // TF SAME: the smaller half in front
  int dx[9], dy[9];
  for (int t = 0; t < 9; ++t) { dx[t] = t % 3 - pad; dy[t] = t / 3 - pad; }
  const int geom[9] = {1, H, W, 1, Ho, Wo, 1, Ho, Wo}, ntaps = 9;
  return probe_conv_tma_phases(x, N, C, geom, wp, bias, y, cout, stride, 1, &ntaps, dx, dy, nullptr, nullptr, nullptr, nullptr, 1, alpha,
                               iters, avg_us);
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

// ---- 5. kind::tf32 MMA peak: an MMA-only loop (no loads) on every SM.  One thread per CTA issues `iters` k-blocks of
//         4 k-steps x `per_step` MMAs (M = 128, N = n, K = 8 each) into one accumulator; operands are pseudo-random
//         fp32 bit patterns already in shared / tensor memory.  mode 0: A from tensor memory (TS form, production's a_small
//         product), mode 1: A from shared memory (SS form), mode 2: the production mix (1 TS + 2 TS = all TS), mode 3: the
//         candidate mix (1 TS + 2 SS).  -> per-CTA clocks; the host times the launch with CUDA events.
namespace {
__global__ void __launch_bounds__(128) mma_peak_kernel(int n, int iters, int mode, long long* __restrict__ clk) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  // [A tile 16 KB][B big n*128][B small n*128][barrier][tmem ptr]
  const uint32_t sbase = smem_u32(sm);
  const uint32_t a_off = 0, b_off = 16384, b2_off = b_off + n * 128, bar_off = b2_off + n * 128, tm_off = bar_off + 8;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t x = 0x9e3779b9u * (blockIdx.x * 128 + tid + 1);
  for (int i = tid; i < (int)(bar_off / 4); i += 128) {
    x = x * 1664525u + 1013904223u;
    reinterpret_cast<uint32_t*>(sm)[i] = 0x3f000000u | (x >> 9);            // floats in [0.5, 1)
  }
  if (tid == 0) { mbar_init(sbase + bar_off, 1); mbar_init_fence(); }
  fence_proxy_async();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + tm_off), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + tm_off);
  {   // A operand in tensor memory: 32 columns of pseudo-random floats per lane (columns 256..287)
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) { x = x * 1664525u + 1013904223u; v[j] = 0x3f000000u | (x >> 9); }
    tc_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256, v);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(n);
    const uint64_t ad = umma_desc_k128(sbase + a_off), bb = umma_desc_k128(sbase + b_off), bs = umma_desc_k128(sbase + b2_off);
    const uint32_t a_t = tmem + 256;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t acc = (it | kk) != 0;
        if (mode == 0) {
          tc_mma_tf32_ts(tmem, a_t + kk * 8, bb + kk * 2, idesc, acc);
        } else if (mode == 1) {
          tc_mma_tf32_ss(tmem, ad + kk * 2, bb + kk * 2, idesc, acc);
        } else if (mode == 2) {
          tc_mma_tf32_ts(tmem, a_t + kk * 8, bb + kk * 2, idesc, acc);
          tc_mma_tf32_ts(tmem, a_t + kk * 8, bs + kk * 2, idesc, 1);
          tc_mma_tf32_ts(tmem, a_t + kk * 8, bb + kk * 2, idesc, 1);
        } else {
          tc_mma_tf32_ts(tmem, a_t + kk * 8, bb + kk * 2, idesc, acc);
          tc_mma_tf32_ss(tmem, ad + kk * 2, bs + kk * 2, idesc, 1);
          tc_mma_tf32_ss(tmem, ad + kk * 2, bb + kk * 2, idesc, 1);
        }
      }
    }
    tc_commit(sbase + bar_off);
  }
  mbar_wait(sbase + bar_off, 0);
  tc_fence_after();
  if (tid == 0 && clk) clk[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
}  // namespace

// n: MMA N (16..256, multiple of 16); iters k-blocks per CTA; grid CTAs (one per SM); -> *ms = launch time (CUDA events, after one
// warm-up launch), clk (device, grid entries or null) = per-CTA clocks of the issue loop incl. completion
extern "C" int probe_mma_peak(int n, int iters, int mode, int grid, float* ms, long long* clk) {
  if (n < 16 || n > 256 || n % 16) return -2;
  const int smem = 16384 + 2 * n * 128 + 64 + 1024;
  if (cudaFuncSetAttribute(mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma_peak_kernel<<<grid, 128, 200 * 1024>>>(n, iters, mode, clk);     // 200 KB: one CTA per SM (each allocates all 512 TMEM columns)
  cudaEventRecord(e0);
  mma_peak_kernel<<<grid, 128, 200 * 1024>>>(n, iters, mode, clk);
  cudaEventRecord(e1);
  const int f = finish();
  float t = 0.f;
  cudaEventElapsedTime(&t, e0, e1);
  if (ms) *ms = t;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  (void)smem;
  return f;
}
