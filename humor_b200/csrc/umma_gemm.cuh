// tcgen05 (5th-gen tensor core) GEMM with fp32-level accuracy:  C[M,N] = A[M,K] * B[N,K]^T
//
// The 1e-5 parity bound on HuMoR's decoder states / prior log-prob rules out a single TF32 pass, so each
// operand is split once into  x = hi + lo  (hi = top 11 mantissa bits, lo = the exact remainder) and three
// MMAs accumulate  hi*hi + lo*hi + hi*lo  into one TMEM accumulator ("3xTF32", ~2^-21 relative).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes {32 floats = 128 B, rows}, SWIZZLE_128B,
//               4 tiles per stage (A_hi, A_lo, B_hi, B_lo) into a 3-stage mbarrier ring
//   warp 1      allocates TMEM, issues tcgen05.mma.cta_group::1.kind::tf32 (one elected lane),
//               tcgen05.commit releases smem stages / signals the epilogue
//   warps 2..5  epilogue: tcgen05.ld (each warp owns its 32-lane TMEM quadrant = 32 output rows),
//               one thread per output row, so a whole GroupNorm group (64 or 32 columns) is private to a
//               thread: bias + two-pass GroupNorm + ReLU (or its reverse) without any shuffle,
//               then writes the result as hi/lo planes for the next layer's TMA loads.
#pragma once
#include <cuda.h>
#include "gemm.cuh"

// tests/host: the same kernel sources run on the CPU against a functional emulation of shared-memory addresses, mbarriers,
// TMA tile loads (SWIZZLE_128B), tcgen05.mma kind::tf32 and TMEM (tests/host/shim/tc_emul.h); every PTX primitive below has
// an emulated twin selected by HB_HOST_SHIM.  Device builds are unaffected (SASS identical before/after this split).
#ifdef HB_HOST_SHIM
#include "tc_emul.h"
#define HB_DYN_SMEM(name) uint8_t* name = tcemu::dyn_smem()
#else
#define HB_DYN_SMEM(name) extern __shared__ uint8_t name[]
#endif

namespace hb {

constexpr int UM_BM = 128;
constexpr int UM_BK = 32;          // tf32 elements per stage row = 128 bytes = one swizzle span

#ifdef HB_HOST_SHIM
using tcemu::smem_u32; using tcemu::mbar_init; using tcemu::mbar_expect_tx; using tcemu::mbar_wait; using tcemu::mbar_arrive;
using tcemu::tma_load_2d; using tcemu::umma_tf32; using tcemu::umma_f16; using tcemu::umma_commit; using tcemu::tmem_ld32;
using tcemu::tmem_alloc; using tcemu::tmem_relinquish; using tcemu::tmem_dealloc; using tcemu::tc_fence_before;
using tcemu::tc_fence_after; using tcemu::mbar_fence_init; using tcemu::ld_shared_u32; using tcemu::l2_policy_evict_last;
using tcemu::tma_load_2d_hint;
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// TMEM allocation (one warp, .sync.aligned): the base address lands in shared memory at `dst`
__device__ __forceinline__ void tmem_alloc(uint32_t dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    // suspend-time hint: the thread sleeps in hardware until the phase completes (event-driven wake-up) instead of spinning through
    // the loop - the spinning lanes of the round-1 kernels issued 20-60 % of all instructions (profiles/r02a, r02b)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y) : "memory");
}
// the same load with an L2 eviction-priority hint (createpolicy): operands every CTA re-reads stay resident while a kernel streams
// gigabytes of output through the L2
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y), "l"(policy) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// fp16 x fp16 -> fp32 (K = 16 per instruction); idesc a/b format 0
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
#endif  // HB_HOST_SHIM (primitives)
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (1024 B between 8-row groups)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// GroupNorm(+ReLU) forward / reverse on GS columns held privately by one thread (fully unrolled: v stays in registers)
template <int GS>
__device__ __forceinline__ void gn_relu_fwd_group(float* v, int col, int row, const GemmEpi& ep) {
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < GS; ++j) { v[j] += ep.bias[col + j]; sum += v[j]; }
  const float mean = sum / (float)GS;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < GS; ++j) { const float d = v[j] - mean; sq += d * d; }
  const float rs = rsqrtf(sq / (float)GS + 1e-5f);
  ep.rstd[(size_t)row * 16 + col / GS] = rs;
#pragma unroll
  for (int j = 0; j < GS; j += 4) {
    float4 xh;
    float* xp = &xh.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      xp[e] = (v[j + e] - mean) * rs;
      v[j + e] = fmaxf(fmaf(ep.gamma[col + j + e], xp[e], ep.beta[col + j + e]), 0.f);
    }
    *reinterpret_cast<float4*>(ep.xhat + (size_t)row * ep.ldxh + col + j) = xh;
  }
}
template <int GS>
__device__ __forceinline__ void gn_relu_bwd_group(float* v, int col, int row, const GemmEpi& ep) {
  float xh[GS];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < GS; j += 4) {
    const float4 x4 = *reinterpret_cast<const float4*>(ep.xhat + (size_t)row * ep.ldxh + col + j);
    xh[j] = x4.x; xh[j + 1] = x4.y; xh[j + 2] = x4.z; xh[j + 3] = x4.w;
  }
#pragma unroll
  for (int j = 0; j < GS; ++j) {
    const float g = ep.gamma[col + j];
    const bool on = fmaf(g, xh[j], ep.beta[col + j]) > 0.f;
    const float u = on ? g * v[j] : 0.f;
    v[j] = u;
    s1 += u;
    s2 += u * xh[j];
  }
  const float rs = ep.rstd[(size_t)row * 16 + col / GS];
  const float inv = 1.f / (float)GS;
#pragma unroll
  for (int j = 0; j < GS; ++j) v[j] = rs * (v[j] - s1 * inv - xh[j] * s2 * inv);
}

template <int BN>
struct UmmaSmem {
  static constexpr int A_TILE = UM_BM * 128;                  // bytes
  static constexpr int B_TILE = BN * 128;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = (BN <= 64) ? 4 : 3;           // 4 x 48 KB or 3 x 64 KB
  static constexpr int TOTAL = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
};

constexpr int UM_CHUNK = 4;        // k-blocks (4 x 32 = K 128) accumulated in TMEM before promotion to fp32 registers

#ifndef HB_HOST_SHIM
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#endif

// The tensor-core accumulator truncates on every add (measured on B200: -2e-5 relative bias over K = 1024 with
// positive operands, error growing ~K), which alone would break the 1e-5 parity bound.  Each K-chunk of 128 is
// therefore accumulated from zero in one of two TMEM buffers and then PROMOTED: the epilogue warps add it to fp32
// register accumulators (round-to-nearest) while the MMA warp already fills the other buffer.
// KS > 1: split-K over a thread-block cluster of KS CTAs (same output tile, disjoint K ranges).  The sequential decoder
// steps have only M = sub-sequences-per-GPU rows, i.e. ~32 output tiles: splitting K four ways puts 128 CTAs on the
// chip and shortens each CTA's dependent TMA->MMA chain 4x.  Partials meet in the leader's (rank 0) shared memory
// through DSMEM stores between two cluster barriers; only the leader runs the epilogue.
#ifdef HB_HOST_SHIM     // tests/host: single-CTA emulation; cluster instantiations (KS > 1) compile but must not run
using tcemu::cluster_ctarank; using tcemu::cluster_sync_all; using tcemu::map_to_cta; using tcemu::st_cluster_v4;
using tcemu::ld_shared_v4; using tcemu::ld_shared_v2; using tcemu::st_shared_v4; using tcemu::bulk_g2s;
#else
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 ld_shared_v2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completing `bytes` on an mbarrier (addresses and size multiples of 16)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif
constexpr int UM_RED_LD = 68;      // floats per row of a partial tile in the leader's smem (64 + pad, float4-aligned)

template <int BN, int EPI, int KS>
__global__ void __launch_bounds__(192, 1)
umma_gemm3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  int M, int N, int K, float* __restrict__ C, float* __restrict__ C_hi, float* __restrict__ C_lo, int ldc,
                  GemmEpi ep) {
  static_assert(BN == 128 || BN == 64, "epilogue is written for 64- or 128-column tiles");
  using SM = UmmaSmem<BN>;
  constexpr int UM_STAGES = SM::STAGES;
  HB_DYN_SMEM(smem_raw);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + UM_STAGES * SM::STAGE;          // full[3] | empty[3] | tfull[2] | tempty[2] | tmem_ptr
  const uint32_t full0 = bars, empty0 = bars + 8 * UM_STAGES, tfull0 = bars + 16 * UM_STAGES, tempty0 = tfull0 + 16,
                 tptr = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(KS == 1 || BN == 64, "split-K partials are laid out for 64-column tiles");
  const uint32_t krank = (KS > 1) ? cluster_ctarank() : 0u;
  const int m0 = blockIdx.y * UM_BM, n0 = (blockIdx.x / KS) * BN;
  const int nkb_all = K / UM_BK;
  const int nkb_per = (nkb_all + KS - 1) / KS;
  const int kb0 = (int)krank * nkb_per;                         // this CTA's K range, in 32-element blocks
  const int nkb = max(0, min(nkb_all, kb0 + nkb_per) - kb0);
  const int nchunk = (nkb + UM_CHUNK - 1) / UM_CHUNK;

  // programmatic dependent launch: the next kernel of the chain may be scheduled now and run ITS prologue (barrier
  // init, TMEM allocation) under this kernel's tail; ours ran under the predecessor's and waits below before reading.
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < UM_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, (uint32_t)(2 * BN));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_u32(tptr);
  // Weight tiles do not depend on the predecessor grid: with ep.b_const the producer fills the ring's B planes while this CTA
  // would otherwise sit in the wait below (in the decoder chain 96 of 128 CTAs of a split-K kernel retire before the leaders'
  // epilogues, so the successor's CTAs are resident for those microseconds).  The stage's expect_tx covers both operands.
  int pre = 0;
  if (ep.b_const && warp == 0 && lane == 0) {
    pre = min(nkb, UM_STAGES);
    for (int kb = 0; kb < pre; ++kb) {                          // first use of every stage: nothing to wait for
      const uint32_t st = base + kb * SM::STAGE;
      mbar_expect_tx(full0 + 8 * kb, SM::STAGE);
      tma_load_2d(st + 2 * SM::A_TILE, &tmB_hi, full0 + 8 * kb, (kb0 + kb) * UM_BK, n0);
      tma_load_2d(st + 2 * SM::A_TILE + SM::B_TILE, &tmB_lo, full0 + 8 * kb, (kb0 + kb) * UM_BK, n0);
    }
  }
  pdl_wait();                                                   // predecessor grid complete, its writes visible

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % UM_STAGES;
        const uint32_t ph = (kb / UM_STAGES) & 1;
        const uint32_t st = base + s * SM::STAGE;
        if (kb < pre) {                                         // B planes are on their way already
          tma_load_2d(st, &tmA_hi, full0 + 8 * s, (kb0 + kb) * UM_BK, m0);
          tma_load_2d(st + SM::A_TILE, &tmA_lo, full0 + 8 * s, (kb0 + kb) * UM_BK, m0);
          continue;
        }
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, SM::STAGE);
        tma_load_2d(st, &tmA_hi, full0 + 8 * s, (kb0 + kb) * UM_BK, m0);
        tma_load_2d(st + SM::A_TILE, &tmA_lo, full0 + 8 * s, (kb0 + kb) * UM_BK, m0);
        tma_load_2d(st + 2 * SM::A_TILE, &tmB_hi, full0 + 8 * s, (kb0 + kb) * UM_BK, n0);
        tma_load_2d(st + 2 * SM::A_TILE + SM::B_TILE, &tmB_lo, full0 + 8 * s, (kb0 + kb) * UM_BK, n0);
      }
    }
    if (KS > 1) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32 (bit 4), A=B=tf32 (2 at bits 7 and 10), K-major both, N>>3 at 17, M>>4 at 24
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        mbar_wait(tempty0 + 8 * buf, ((c >> 1) & 1) ^ 1);      // epilogue has drained this TMEM buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        const int kb_end = min(nkb, (c + 1) * UM_CHUNK);
        for (int kb = c * UM_CHUNK; kb < kb_end; ++kb) {
          const int s = kb % UM_STAGES;
          const uint32_t ph = (kb / UM_STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t st = base + s * SM::STAGE;
#pragma unroll
          for (int k = 0; k < UM_BK / 8; ++k) {                // one UMMA consumes K = 8 tf32 = 32 bytes
            const uint64_t a_hi = umma_desc_sw128(st + k * 32);
            const uint64_t a_lo = umma_desc_sw128(st + SM::A_TILE + k * 32);
            const uint64_t b_hi = umma_desc_sw128(st + 2 * SM::A_TILE + k * 32);
            const uint64_t b_lo = umma_desc_sw128(st + 2 * SM::A_TILE + SM::B_TILE + k * 32);
            umma_tf32(tacc, a_hi, b_hi, idesc, (kb != c * UM_CHUNK) || (k != 0));
            umma_tf32(tacc, a_lo, b_hi, idesc, 1);
            umma_tf32(tacc, a_hi, b_lo, idesc, 1);
          }
          umma_commit(empty0 + 8 * s);                         // frees the stage once these MMAs retire
        }
        umma_commit(tfull0 + 8 * buf);                         // chunk complete -> promote
      }
    }
    if (KS > 1) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }
  } else {
    // ------------------------------------------------------------------ epilogue warps: one thread per output row
    const int q = warp & 3;                                    // TMEM lane quadrant this warp may access
    const int row = m0 + q * 32 + lane;
    const bool rok = row < M;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      const int buf = c & 1;
      mbar_wait(tfull0 + 8 * buf, (c >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float t[32];
        tmem_ld32(trow + buf * BN + c0, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += t[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    if (KS > 1) {
      cluster_sync_all();                                      // #1: every CTA's stage memory is idle now
      if (krank != 0) {
        const uint32_t dst = map_to_cta(base, 0) + (uint32_t)(((krank - 1) * UM_BM + q * 32 + lane) * UM_RED_LD) * 4u;
#pragma unroll
        for (int j = 0; j < BN; j += 4) st_cluster_v4(dst + j * 4, acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
      cluster_sync_all();                                      // #2: partials have landed in the leader
      if (krank == 0) {
#pragma unroll
        for (int r = 0; r < KS - 1; ++r) {
          const uint32_t src = base + (uint32_t)((r * UM_BM + q * 32 + lane) * UM_RED_LD) * 4u;
#pragma unroll
          for (int j = 0; j < BN; j += 4) {
            const float4 v = ld_shared_v4(src + j * 4);
            acc[j] += v.x; acc[j + 1] += v.y; acc[j + 2] += v.z; acc[j + 3] += v.w;
          }
        }
      }
    }
    auto store4 = [&](int col, float a, float b, float c_, float d) {
      if (col + 3 < N) {
        if (C) *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = make_float4(a, b, c_, d);
        if (C_hi) {
          const float4 h = make_float4(tf32_hi(a), tf32_hi(b), tf32_hi(c_), tf32_hi(d));
          *reinterpret_cast<float4*>(C_hi + (size_t)row * ldc + col) = h;
          *reinterpret_cast<float4*>(C_lo + (size_t)row * ldc + col) = make_float4(a - h.x, b - h.y, c_ - h.z, d - h.w);
        }
      } else {
        const float o[4] = {a, b, c_, d};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (col + jj < N) {
            if (C) C[(size_t)row * ldc + col + jj] = o[jj];
            if (C_hi) { const float h = tf32_hi(o[jj]); C_hi[(size_t)row * ldc + col + jj] = h; C_lo[(size_t)row * ldc + col + jj] = o[jj] - h; }
          }
      }
    };
    if (rok && krank == 0) {
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 64) {
        const int col = n0 + c0;
        if (col < N) {
          if (EPI == EPI_BIAS) {
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[c0 + j] += (ep.bias && col + j < N) ? ep.bias[col + j] : 0.f;
          } else if (EPI == EPI_GN_RELU) {
            if (ep.gsize == 64) gn_relu_fwd_group<64>(acc + c0, col, row, ep);
            else { gn_relu_fwd_group<32>(acc + c0, col, row, ep); gn_relu_fwd_group<32>(acc + c0 + 32, col + 32, row, ep); }
          } else if (col < ep.Cch) {
            if (ep.gsize == 64) gn_relu_bwd_group<64>(acc + c0, col, row, ep);
            else { gn_relu_bwd_group<32>(acc + c0, col, row, ep); gn_relu_bwd_group<32>(acc + c0 + 32, col + 32, row, ep); }
          }
#pragma unroll
          for (int j = 0; j < 64; j += 4) store4(col + j, acc[c0 + j], acc[c0 + j + 1], acc[c0 + j + 2], acc[c0 + j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, (uint32_t)(2 * BN));
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Persistent form of the 128x128-tile kernel for the BATCHED products (the prior network over all S*B rows, 118 x 8 tiles at the
// benchmark size; the blend GEMM of dense-LBS form 1).  Measured on the B200 (profiles/r02m_prior_gemm_*): one tile per CTA spent
// 18.6 us of its 37.6 us in prologue + epilogue with the tensor pipe idle (the K = 96 reverse GEMM is 119 us of pure epilogue:
// issue slots 9 % busy, 1.5 warps per scheduler).  Here a CTA per SM walks tiles blockIdx.x, + gridDim.x, ...; TMEM holds FOUR
// 128-column buffers = two ping-pong pairs, one pair per EPILOGUE GROUP of four warps: group g promotes the chunks of tiles
// g, g+2, ... and runs their epilogue while the MMA lane is already accumulating the next tile into the other pair and group 1-g
// promotes it.  The operand ring runs across tiles, so the next tile's loads are in flight during the last chunk as well.
// ------------------------------------------------------------------------------------------------------------------------------
// Epilogue I/O: a thread owns a ROW of the tile (the TMEM lane), so stores straight from registers touch 32 different 128-byte
// lines per instruction with 16 of 32 bytes per sector used - with both groups' epilogues running the K = 96 reverse GEMM stayed at
// 114 us (profiles/r02n_prior_gemm_persistent_set_full.txt).  Every global access of the epilogue therefore goes through a per-warp
// staging tile of 32 rows x 16 columns (row stride 20 floats: 128-bit accesses conflict-free in both directions): registers ->
// tile -> 8 rows x 64 contiguous bytes per instruction, and the reverse for the saved x-hat of the GroupNorm reverse.
constexpr int UMP_THREADS = 64 + 2 * 128;
constexpr int UMP_SLD = 20;                                    // floats per staging row
constexpr int UMP_STAGE_W = 32 * UMP_SLD * 4;                  // bytes per epilogue warp
constexpr int UMP_STAGING = 8 * UMP_STAGE_W;
constexpr int UMP_SMEM = UmmaSmem<128>::TOTAL + UMP_STAGING;

template <int EPI>
__global__ void __launch_bounds__(UMP_THREADS, 1)
umma_gemm3p_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   int M, int N, int K, float* __restrict__ C, float* __restrict__ C_hi, float* __restrict__ C_lo, int ldc,
                   GemmEpi ep) {
  constexpr int BN = 128;
  using SM = UmmaSmem<BN>;
  constexpr int UM_STAGES = SM::STAGES;
  HB_DYN_SMEM(smem_raw);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bars = base + UM_STAGES * SM::STAGE + UMP_STAGING;   // full[3] | empty[3] | tfull[4] | tempty[4] | tmem_ptr
  const uint32_t full0 = bars, empty0 = bars + 8 * UM_STAGES, tfull0 = bars + 16 * UM_STAGES, tempty0 = tfull0 + 32,
                 tptr = tempty0 + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntn = (N + BN - 1) / BN, ntm = (M + UM_BM - 1) / UM_BM;
  const int ntiles = ntn * ntm;
  const int nkb = K / UM_BK;
  const int nchunk = (nkb + UM_CHUNK - 1) / UM_CHUNK;
  const int nuse0 = (nchunk + 1) >> 1, nuse1 = nchunk >> 1;   // uses of a pair's even / odd buffer per tile

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < UM_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 4; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_u32(tptr);
  pdl_wait();                                                   // predecessor grid complete, its writes visible

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;                                               // k-blocks issued by this CTA (ring position)
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile / ntn) * UM_BM, n0 = (tile % ntn) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % UM_STAGES;
          const uint32_t st = base + s * SM::STAGE;
          mbar_wait(empty0 + 8 * s, ((it / UM_STAGES) & 1) ^ 1);
          mbar_expect_tx(full0 + 8 * s, SM::STAGE);
          tma_load_2d(st, &tmA_hi, full0 + 8 * s, kb * UM_BK, m0);
          tma_load_2d(st + SM::A_TILE, &tmA_lo, full0 + 8 * s, kb * UM_BK, m0);
          tma_load_2d(st + 2 * SM::A_TILE, &tmB_hi, full0 + 8 * s, kb * UM_BK, n0);
          tma_load_2d(st + 2 * SM::A_TILE + SM::B_TILE, &tmB_lo, full0 + 8 * s, kb * UM_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      int it = 0, i = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int g = i & 1, j = i >> 1;                        // epilogue group of this tile, its tile count
        for (int c = 0; c < nchunk; ++c) {
          const int b = 2 * g + (c & 1);
          const int u = j * ((c & 1) ? nuse1 : nuse0) + (c >> 1);   // uses of buffer b before this one
          mbar_wait(tempty0 + 8 * b, (u & 1) ^ 1);              // group g has promoted that chunk
          tc_fence_after();
          const uint32_t tacc = tmem_base + b * BN;
          const int kb_end = min(nkb, (c + 1) * UM_CHUNK);
          for (int kb = c * UM_CHUNK; kb < kb_end; ++kb, ++it) {
            const int s = it % UM_STAGES;
            mbar_wait(full0 + 8 * s, (it / UM_STAGES) & 1);
            tc_fence_after();
            const uint32_t st = base + s * SM::STAGE;
#pragma unroll
            for (int k = 0; k < UM_BK / 8; ++k) {
              const uint64_t a_hi = umma_desc_sw128(st + k * 32);
              const uint64_t a_lo = umma_desc_sw128(st + SM::A_TILE + k * 32);
              const uint64_t b_hi = umma_desc_sw128(st + 2 * SM::A_TILE + k * 32);
              const uint64_t b_lo = umma_desc_sw128(st + 2 * SM::A_TILE + SM::B_TILE + k * 32);
              umma_tf32(tacc, a_hi, b_hi, idesc, (kb != c * UM_CHUNK) || (k != 0));
              umma_tf32(tacc, a_lo, b_hi, idesc, 1);
              umma_tf32(tacc, a_hi, b_lo, idesc, 1);
            }
            umma_commit(empty0 + 8 * s);
          }
          umma_commit(tfull0 + 8 * b);
        }
      }
    }
  } else {
    const int g = (warp - 2) >> 2;                              // epilogue group
    const int q = warp & 3;                                     // TMEM lane quadrant this warp may access
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    int i = g;
    for (int tile = blockIdx.x + g * gridDim.x; tile < ntiles; tile += 2 * gridDim.x, i += 2) {
      const int j = i >> 1;
      const int m0 = (tile / ntn) * UM_BM, n0 = (tile % ntn) * BN;
      const int row = m0 + q * 32 + lane;
      const bool rok = row < M;
      float acc[BN];
#pragma unroll
      for (int jj = 0; jj < BN; ++jj) acc[jj] = 0.f;
      for (int c = 0; c < nchunk; ++c) {
        const int b = 2 * g + (c & 1);
        const int u = j * ((c & 1) ? nuse1 : nuse0) + (c >> 1);
        mbar_wait(tfull0 + 8 * b, u & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float t[32];
          tmem_ld32(trow + b * BN + c0, t);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) acc[c0 + jj] += t[jj];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * b);
      }
      float* S = reinterpret_cast<float*>(gbase + UM_STAGES * SM::STAGE + (warp - 2) * UMP_STAGE_W);
      const int r8 = lane >> 2, cg = (lane & 3) * 4;            // transfer phase: 8 rows x 4 column quads per instruction
      const int row0 = m0 + q * 32;
      // 16 values of this thread's row (columns col0 .. col0+15) -> dst[row][col0 ..], coalesced
      auto put16 = [&](float* dst, int ld, int col0, const float* v) {
        float4* sr = reinterpret_cast<float4*>(S + lane * UMP_SLD);
#pragma unroll
        for (int k = 0; k < 4; ++k) sr[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        __syncwarp();
#pragma unroll
        for (int p_ = 0; p_ < 4; ++p_) {
          const int r = p_ * 8 + r8, grow = row0 + r, col = col0 + cg;
          const float4 x = *reinterpret_cast<const float4*>(S + r * UMP_SLD + cg);
          if (grow < M) {
            float* o = dst + (size_t)grow * ld + col;
            if (col + 3 < N) *reinterpret_cast<float4*>(o) = x;
            else { if (col < N) o[0] = x.x; if (col + 1 < N) o[1] = x.y; if (col + 2 < N) o[2] = x.z; }
          }
        }
        __syncwarp();
      };
      // the reverse: src[row][col0 .. col0+15] of the warp's 32 rows -> 16 values of this thread's row (columns < width, else 0)
      auto get16 = [&](const float* src, int ld, int width, int col0, float* v) {
#pragma unroll
        for (int p_ = 0; p_ < 4; ++p_) {
          const int r = p_ * 8 + r8, grow = row0 + r, col = col0 + cg;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < M && col + 3 < width) x = *reinterpret_cast<const float4*>(src + (size_t)grow * ld + col);
          *reinterpret_cast<float4*>(S + r * UMP_SLD + cg) = x;
        }
        __syncwarp();
        const float4* sr = reinterpret_cast<const float4*>(S + lane * UMP_SLD);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 x = sr[k]; v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w; }
        __syncwarp();
      };
      // result columns col0 .. col0+15 of the row: exact value and / or hi/lo planes
      auto put_result16 = [&](int col0, const float* v) {
        if (C) put16(C, ldc, col0, v);
        if (C_hi) {
          float h[16], l[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { h[e] = tf32_hi(v[e]); l[e] = v[e] - h[e]; }
          put16(C_hi, ldc, col0, h);
          put16(C_lo, ldc, col0, l);
        }
      };
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 64) {
        const int col = n0 + c0;
        if (col >= N) break;                                    // warp-uniform
        float* v = acc + c0;
        if (EPI == EPI_BIAS) {
#pragma unroll
          for (int k = 0; k < 64; k += 16) {
            if (col + k >= N) break;
#pragma unroll
            for (int e = 0; e < 16; ++e) v[k + e] += (ep.bias && col + k + e < N) ? ep.bias[col + k + e] : 0.f;
            put_result16(col + k, v + k);
          }
        } else if (EPI == EPI_GN_RELU) {
          // GroupNorm over the 64 columns (one group: the launcher sends other group sizes to the one-tile kernel) + ReLU;
          // x-hat and 1/sigma go to the tape for the reverse pass
          float sum = 0.f, sq = 0.f;
#pragma unroll
          for (int e = 0; e < 64; ++e) { v[e] += ep.bias[col + e]; sum += v[e]; }
          const float mean = sum * (1.f / 64.f);
#pragma unroll
          for (int e = 0; e < 64; ++e) { const float d = v[e] - mean; sq += d * d; }
          const float rs = rsqrtf(sq * (1.f / 64.f) + 1e-5f);
          if (rok) ep.rstd[(size_t)row * 16 + col / 64] = rs;
#pragma unroll
          for (int k = 0; k < 64; k += 16) {
            float xh[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              xh[e] = (v[k + e] - mean) * rs;
              v[k + e] = fmaxf(fmaf(ep.gamma[col + k + e], xh[e], ep.beta[col + k + e]), 0.f);
            }
            put16(ep.xhat, ep.ldxh, col + k, xh);
            put_result16(col + k, v + k);
          }
        } else {
          if (col < ep.Cch) {
            float xh[64];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 64; k += 16) get16(ep.xhat, ep.ldxh, ep.ldxh, col + k, xh + k);
#pragma unroll
            for (int e = 0; e < 64; ++e) {
              const float gm = ep.gamma[col + e];
              const bool on = fmaf(gm, xh[e], ep.beta[col + e]) > 0.f;
              const float u = on ? gm * v[e] : 0.f;
              v[e] = u;
              s1 += u;
              s2 += u * xh[e];
            }
            const float rs = ep.rstd[(size_t)(rok ? row : 0) * 16 + col / 64];
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] = rs * (v[e] - s1 * (1.f / 64.f) - xh[e] * s2 * (1.f / 64.f));
          }
#pragma unroll
          for (int k = 0; k < 64; k += 16) {
            if (col + k >= N) break;
            put_result16(col + k, v + k);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

}  // namespace hb
