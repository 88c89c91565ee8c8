// Blend-shape contraction of the dense SMPL+H forward as a PERSISTENT tcgen05 kernel:
//
//   v_posed[frame, 3v+d] = v_template[3v+d] + feat[frame, :224] . blend_t[3v+d, :224]       (3xTF32: hi*hi + lo*hi + hi*lo)
//
// Why not umma_gemm3_kernel<128,BIAS> (one 128x128 tile per CTA): with K = 224 a tile is 7 k-blocks, so prologue, TMA->MMA
// chain and epilogue run back to back (ncu: tensor pipe 22 % active, 9 % occupancy, 13 us per tile of which ~3 us MMA), and
// every tile pulls 448 KB of operand planes from L2 (298 MB per 512-frame slab).  Here
//   * one CTA per SM walks 128 x 256 tiles (operand traffic 218 MB per 512 frames, half as many tiles),
//   * warp 0 = TMA producer running ahead across tile boundaries (2-stage ring of 96 KB: A hi/lo 2 x 16 KB, B hi/lo 2 x 32 KB),
//   * warp 1 = MMA issuer; a tile's 7 k-blocks accumulate into ONE of two 256-column TMEM buffers (K = 224 is short enough
//     that the tensor core's truncating accumulator stays at ~2e-6 relative: no promotion chunks, unlike the K = 1024 MLPs),
//     so tile i+1's MMAs run under tile i's epilogue,
//   * warps 2..5 = epilogue, one thread per frame row: tcgen05.ld 32 columns at a time, + template, 16-byte row stores
//     (columns >= ncols are not written).
//   * `fast` (blend form 3): k-block 0 (the 16 betas + the first 16 pose-feature columns: shape offsets of up to 0.3 m) keeps
//     the three passes; the other 6 k-blocks (pose offsets, a few cm) run ONE pass on the hi planes, which hold the operands
//     ROUNDED to tf32: 36 instead of 84 MMAs and 384 instead of 672 KB of operand planes per tile; vertex error <= 7e-5 m at
//     extreme poses (2e-5 m typical) against the 1e-4 m bound, where the three-pass form is at 1e-6 m.
// Same barrier / TMEM protocol as lbs_fused_kernel (lbs_fused.cuh), which is verified on the B200.
#pragma once
#include "umma_gemm.cuh"

namespace hb {

constexpr int LB_BN = 256;
constexpr int LB_STAGES = 2;
constexpr int LB_A_TILE = UM_BM * 128;              // bytes: 128 rows x 128 B
constexpr int LB_B_TILE = LB_BN * 128;
constexpr int LB_STAGE = 2 * LB_A_TILE + 2 * LB_B_TILE;     // 96 KB
constexpr int LB_SMEM = LB_STAGES * LB_STAGE + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
lbs_blend_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, int M, int ncols, int K,
                 const float* __restrict__ bias, float* __restrict__ C, int ldc, int fast) {
  HB_DYN_SMEM(smem_raw);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + LB_STAGES * LB_STAGE;
  const uint32_t full0 = bars, empty0 = bars + 8 * LB_STAGES, tfull0 = bars + 16 * LB_STAGES, tempty0 = tfull0 + 16,
                 tptr = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrt = (M + UM_BM - 1) / UM_BM, nct = (ncols + LB_BN - 1) / LB_BN;
  const int ntiles = nrt * nct;
  const int nkb = K / UM_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < LB_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
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

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;                                                // k-blocks issued so far (ring position)
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        // column-major tile order: the CTAs of a wave share the same few B tiles (L2 reuse of the 64 KB-per-k-block planes)
        const int m0 = (t % nrt) * UM_BM, n0 = (t / nrt) * LB_BN;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % LB_STAGES;
          mbar_wait(empty0 + 8 * s, ((g / LB_STAGES) & 1) ^ 1);
          const uint32_t st = base + s * LB_STAGE;
          const bool one_pass = fast && kb > 0;                 // only the hi planes travel for single-pass k-blocks
          mbar_expect_tx(full0 + 8 * s, one_pass ? LB_A_TILE + LB_B_TILE : LB_STAGE);
          tma_load_2d(st, &tmA_hi, full0 + 8 * s, kb * UM_BK, m0);
          tma_load_2d(st + 2 * LB_A_TILE, &tmB_hi, full0 + 8 * s, kb * UM_BK, n0);
          if (!one_pass) {
            tma_load_2d(st + LB_A_TILE, &tmA_lo, full0 + 8 * s, kb * UM_BK, m0);
            tma_load_2d(st + 2 * LB_A_TILE + LB_B_TILE, &tmB_lo, full0 + 8 * s, kb * UM_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(LB_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      int g = 0, tc = 0;                                        // k-blocks / tiles consumed so far
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
        const int buf = tc & 1;
        mbar_wait(tempty0 + 8 * buf, ((tc >> 1) & 1) ^ 1);      // epilogue has drained this TMEM buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * LB_BN;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % LB_STAGES;
          mbar_wait(full0 + 8 * s, (g / LB_STAGES) & 1);
          tc_fence_after();
          const uint32_t st = base + s * LB_STAGE;
#pragma unroll
          for (int k = 0; k < UM_BK / 8; ++k) {
            const uint64_t a_hi = umma_desc_sw128(st + k * 32);
            const uint64_t a_lo = umma_desc_sw128(st + LB_A_TILE + k * 32);
            const uint64_t b_hi = umma_desc_sw128(st + 2 * LB_A_TILE + k * 32);
            const uint64_t b_lo = umma_desc_sw128(st + 2 * LB_A_TILE + LB_B_TILE + k * 32);
            umma_tf32(tacc, a_hi, b_hi, idesc, (kb != 0) || (k != 0));
            if (!(fast && kb > 0)) {
              umma_tf32(tacc, a_lo, b_hi, idesc, 1);
              umma_tf32(tacc, a_hi, b_lo, idesc, 1);
            }
          }
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull0 + 8 * buf);
      }
    }
  } else {
    const int q = warp & 3;                                     // TMEM lane quadrant of this warp
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    int tc = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
      const int buf = tc & 1;
      const int m0 = (t % nrt) * UM_BM, n0 = (t / nrt) * LB_BN;
      const int row = m0 + q * 32 + lane;
      const bool rok = row < M;
      float* crow = C + (size_t)(rok ? row : 0) * ldc;
      mbar_wait(tfull0 + 8 * buf, (tc >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < LB_BN; c0 += 32) {
        float v[32];
        tmem_ld32(trow + buf * LB_BN + c0, v);                  // all lanes take part (sync.aligned) whatever rok
        const int col = n0 + c0;
        if (rok && col < ncols) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col + j + 3 < ncols) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + j));
              *reinterpret_cast<float4*>(crow + col + j) = make_float4(v[j] + b.x, v[j + 1] + b.y, v[j + 2] + b.z, v[j + 3] + b.w);
            } else {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (col + j + jj < ncols) crow[col + j + jj] = v[j + jj] + bias[col + j + jj];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace hb
