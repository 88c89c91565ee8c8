// C[M,N] = A[M,K] . B[N,K]^T (+ bias) at fp32-level accuracy from FOUR-byte operand elements: fp16 hi + scaled fp16 lo planes.
//
// Why: every tcgen05 GEMM of the step is bound by operand INGEST (~32 B/clk per SM), and the 3xTF32 scheme of
// umma_gemm.cuh ships 8 bytes per element (two fp32 planes of which the tensor core reads 19 bits each).  Here
//     x = h + l * 2^-11,   h = fp16(x),   l = fp16((x - h) * 2^11)            (x - h is exact in fp32; 22-bit significand)
//     A.B = h_a.h_b  +  2^-11 (l_a.h_b + h_a.l_b)           (the l.l term is 2^-22 of the product: dropped, as 3xTF32 drops lo.lo)
// so a k-block of 64 elements has the same 128-byte rows as a 32-float one (same TMA boxes, same SWIZZLE_128B descriptors, same
// stage layout as UmmaSmem), is consumed by 12 tcgen05.mma kind::f16 (K = 16) instead of 24 kind::tf32 (K = 8), and moves half
// the bytes.  The two sums have different scales, so they accumulate in TWO TMEM ranges (D1: h.h, D2: the cross terms); the
// epilogue warps promote acc += D1 + 2^-11 D2 into fp32 registers per K-chunk of 128 - the same ping-pong promotion that keeps
// the tensor core's truncating accumulator out of the result in umma_gemm3_kernel.  TMEM: 2 buffers x 2 ranges x BN columns.
// fp16's range: |x| < 65504 and a 6e-8 absolute floor on h (the lo plane recovers it to ~1e-11): activations, weights, blend
// planes - NOT the reverse chain's gradients (those would need bf16x3).
// Split-K over a 4-CTA cluster as in umma_gemm3_kernel (partials through DSMEM, leader epilogue).
// Status: utility entry point (humor_umma_gemm16) + CPU emulation (tests/test_emul_product.py); not yet executed on hardware;
// nothing on the product path uses it.
#pragma once
#include "umma_gemm.cuh"
#include "umma_split16.cuh"

namespace hb {

constexpr int U16_BK = 64;         // halves per k-block row = 128 bytes
constexpr int U16_CHUNK = 2;       // k-blocks (K = 128) per TMEM accumulation before promotion

__global__ void split16_kernel(const float* __restrict__ x, unsigned short* __restrict__ h, unsigned short* __restrict__ l, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) split16(x[i], h[i], l[i]);
}

// Outputs of one launch: C (fp32, nullable) and/or the fp16 hi/lo planes of the result (C16_h / C16_l, ld16 halves per row, nullable):
// the operand planes of the next layer, written by the producing epilogue as umma_gemm3_kernel does for its fp32 planes.
// EPI_BIAS: + ep.bias (nullable).  EPI_GN_RELU: + bias, GroupNorm over ep.gsize channels, ReLU; x-hat and 1/sigma go to the tape
// (ep.xhat, ep.rstd) exactly as in umma_gemm3_kernel, so the reverse pass is unchanged.
template <int BN, int EPI, int KS>
__global__ void __launch_bounds__(192, 1)
umma_gemm16_kernel(const __grid_constant__ CUtensorMap tmA_h, const __grid_constant__ CUtensorMap tmA_l,
                   const __grid_constant__ CUtensorMap tmB_h, const __grid_constant__ CUtensorMap tmB_l, int M, int N, int K,
                   float* __restrict__ C, int ldc, unsigned short* __restrict__ C16_h, unsigned short* __restrict__ C16_l, int ld16,
                   GemmEpi ep) {
  static_assert(EPI == EPI_BIAS || EPI == EPI_GN_RELU, "forward epilogues only (fp16's range does not hold the reverse chain)");
  static_assert(BN == 128 || BN == 64, "64- or 128-column tiles");
  static_assert(KS == 1 || BN == 64, "split-K partials are laid out for 64-column tiles");
  using SM = UmmaSmem<BN>;                                      // same stage: A_h | A_l | B_h | B_l, rows of 128 bytes
  constexpr int STAGES = SM::STAGES;
  HB_DYN_SMEM(smem_raw);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * SM::STAGE;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16, tptr = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t krank = (KS > 1) ? cluster_ctarank() : 0u;
  const int m0 = blockIdx.y * UM_BM, n0 = (blockIdx.x / KS) * BN;
  const int nkb_all = K / U16_BK;
  const int nkb_per = (nkb_all + KS - 1) / KS;
  const int kb0 = (int)krank * nkb_per;
  const int nkb = max(0, min(nkb_all, kb0 + nkb_per) - kb0);
  const int nchunk = (nkb + U16_CHUNK - 1) / U16_CHUNK;

  pdl_launch_dependents();             // programmatic dependent launch, as umma_gemm3_kernel: prologue under the predecessor's tail
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, (uint32_t)(4 * BN));                       // 2 buffers x (D1 | D2)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_u32(tptr);
  pdl_wait();                          // predecessor grid complete, its writes (our A planes) visible

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
        const uint32_t st = base + s * SM::STAGE;
        mbar_expect_tx(full0 + 8 * s, SM::STAGE);
        tma_load_2d(st, &tmA_h, full0 + 8 * s, (kb0 + kb) * U16_BK, m0);
        tma_load_2d(st + SM::A_TILE, &tmA_l, full0 + 8 * s, (kb0 + kb) * U16_BK, m0);
        tma_load_2d(st + 2 * SM::A_TILE, &tmB_h, full0 + 8 * s, (kb0 + kb) * U16_BK, n0);
        tma_load_2d(st + 2 * SM::A_TILE + SM::B_TILE, &tmB_l, full0 + 8 * s, (kb0 + kb) * U16_BK, n0);
      }
    }
    if (KS > 1) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = f32 (1 at bit 4), A = B = f16 (0 at bits 7 and 10), K-major both, N >> 3 at 17, M >> 4 at 24
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        mbar_wait(tempty0 + 8 * buf, ((c >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d1 = tmem_base + buf * 2 * BN, d2 = d1 + BN;
        const int kb_end = min(nkb, (c + 1) * U16_CHUNK);
        for (int kb = c * U16_CHUNK; kb < kb_end; ++kb) {
          const int s = kb % STAGES;
          mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
          tc_fence_after();
          const uint32_t st = base + s * SM::STAGE;
#pragma unroll
          for (int k = 0; k < U16_BK / 16; ++k) {               // one UMMA consumes K = 16 halves = 32 bytes
            const uint64_t a_h = umma_desc_sw128(st + k * 32);
            const uint64_t a_l = umma_desc_sw128(st + SM::A_TILE + k * 32);
            const uint64_t b_h = umma_desc_sw128(st + 2 * SM::A_TILE + k * 32);
            const uint64_t b_l = umma_desc_sw128(st + 2 * SM::A_TILE + SM::B_TILE + k * 32);
            const uint32_t first = (kb != c * U16_CHUNK) || (k != 0);
            umma_f16(d1, a_h, b_h, idesc, first);
            umma_f16(d2, a_l, b_h, idesc, first);
            umma_f16(d2, a_h, b_l, idesc, 1);
          }
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull0 + 8 * buf);
      }
    }
    if (KS > 1) { __syncwarp(); cluster_sync_all(); cluster_sync_all(); }
  } else {
    const int q = warp & 3;
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
        float t1[32], t2[32];
        tmem_ld32(trow + buf * 2 * BN + c0, t1);
        tmem_ld32(trow + buf * 2 * BN + BN + c0, t2);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += fmaf(t2[j], 0.00048828125f, t1[j]);       // D1 + 2^-11 D2
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
    if (rok && krank == 0) {
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 64) {
        const int col0 = n0 + c0;
        if (col0 < N) {
          if (EPI == EPI_BIAS) {
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[c0 + j] += (ep.bias && col0 + j < N) ? ep.bias[col0 + j] : 0.f;
          } else {
            if (ep.gsize == 64) gn_relu_fwd_group<64>(acc + c0, col0, row, ep);
            else { gn_relu_fwd_group<32>(acc + c0, col0, row, ep); gn_relu_fwd_group<32>(acc + c0 + 32, col0 + 32, row, ep); }
          }
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const int col = col0 + j;
            const float o[4] = {acc[c0 + j], acc[c0 + j + 1], acc[c0 + j + 2], acc[c0 + j + 3]};
            unsigned short h[4], l[4];
            if (C16_h) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) split16(o[jj], h[jj], l[jj]);
            }
            if (col + 3 < N) {
              if (C) *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = make_float4(o[0], o[1], o[2], o[3]);
              if (C16_h) {
                *reinterpret_cast<uint2*>(C16_h + (size_t)row * ld16 + col) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
                *reinterpret_cast<uint2*>(C16_l + (size_t)row * ld16 + col) = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (col + jj < N) {
                  if (C) C[(size_t)row * ldc + col + jj] = o[jj];
                  if (C16_h) { C16_h[(size_t)row * ld16 + col + jj] = h[jj]; C16_l[(size_t)row * ld16 + col + jj] = l[jj]; }
                }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, (uint32_t)(4 * BN));
  }
}

}  // namespace hb
