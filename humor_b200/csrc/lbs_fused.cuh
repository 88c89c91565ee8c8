// Dense SMPL+H forward in ONE persistent tcgen05 kernel: blend GEMM (3xTF32) + skinning + coalesced vertex store.
//
//   v_posed[frame, 3v+d] = feat[frame, :224] . blendF[3v+d, :224]     (betas | R-I | 1 -> shape + pose offsets + template)
//   out[frame, v, :]     = sum_k w_k (A[frame, j_k] . [v_posed; 1]) + trans[frame]
//
// The unfused path wrote v_posed (82.7 KB/frame) and read it back in a second kernel; here the accumulator never
// leaves the SM.  Output tiles are 128 frames x 128 columns = 42 vertices (two halves of 21 vertices * 3 + 1 pad
// column; the host packs blendF in this column order), walked by a persistent grid of one CTA per SM:
//   warp 0      TMA producer (A/B hi+lo boxes, 3-stage ring that runs ahead across tile boundaries)
//   warp 1      tcgen05.mma issuer; the two K-chunks of a tile alternate between two TMEM buffers, so the MMAs of
//               tile i+1 overlap the skinning arithmetic of tile i
//   warps 2..5  epilogue, one thread per frame row: promote the chunks to fp32 registers (same numerics as
//               umma_gemm3_kernel), park the row in shared memory, skin 21 vertices with the frame's 3x4 transforms
//               (kept in registers while consecutive vertices share joints), then the warp writes its 32 rows
//               coalesced (252 contiguous bytes per row and half) with streaming stores.
#pragma once
#include "umma_gemm.cuh"

namespace hb {

struct LbsFusedArgs {
  int N;                 // frames
  int num_verts;
  int nrt, nct;          // row tiles (128 frames), column tiles (42 vertices)
  const int* fw_idx;     // [num_verts][WK]  joint * 12, slots sorted by joint id, zero-weight slots last
  const float* fw_val;   // [num_verts][WK]
  const float* A;        // [N][52][12] skinning transforms
  const float* trans;    // [N][3]
  float* out;            // [N][num_verts][3]
};

constexpr int LF_STAGES = 3;
constexpr int LF_STAGE = 4 * UM_BM * 128;          // A_hi, A_lo, B_hi, B_lo boxes of 128 rows x 128 bytes
constexpr int LF_SLD = 65;                          // staging row stride in floats (odd: conflict-free row-per-thread access)
constexpr int LF_STAGING = UM_BM * LF_SLD * 4;
constexpr int LF_SMEM = LF_STAGES * LF_STAGE + LF_STAGING + 1024 + 256;
constexpr int LF_VH = 21;                           // vertices per 64-column half tile
constexpr int LF_VT = 42;                           // vertices per tile

template <int WK>
__global__ void __launch_bounds__(192, 1)
lbs_fused_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, int K, LbsFusedArgs a) {
  HB_DYN_SMEM(smem_raw);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  float* S = reinterpret_cast<float*>(gbase + LF_STAGES * LF_STAGE);
  const uint32_t bars = base + LF_STAGES * LF_STAGE + LF_STAGING;
  const uint32_t full0 = bars, empty0 = bars + 8 * LF_STAGES, tfull0 = bars + 16 * LF_STAGES, tempty0 = tfull0 + 16,
                 tptr = tempty0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = a.nrt * a.nct;
  const int nkb = K / UM_BK;
  const int nchunk = (nkb + UM_CHUNK - 1) / UM_CHUNK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < LF_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 256u);
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
        const int m0 = (t / a.nct) * UM_BM, n0 = (t % a.nct) * 128;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % LF_STAGES;
          mbar_wait(empty0 + 8 * s, ((g / LF_STAGES) & 1) ^ 1);
          const uint32_t st = base + s * LF_STAGE;
          mbar_expect_tx(full0 + 8 * s, LF_STAGE);
          tma_load_2d(st, &tmA_hi, full0 + 8 * s, kb * UM_BK, m0);
          tma_load_2d(st + 16384, &tmA_lo, full0 + 8 * s, kb * UM_BK, m0);
          tma_load_2d(st + 32768, &tmB_hi, full0 + 8 * s, kb * UM_BK, n0);
          tma_load_2d(st + 49152, &tmB_lo, full0 + 8 * s, kb * UM_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      int g = 0, cc = 0;                                        // k-blocks / chunks consumed so far
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int c = 0; c < nchunk; ++c, ++cc) {
          const int buf = cc & 1;
          mbar_wait(tempty0 + 8 * buf, ((cc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * 128;
          const int kb_end = min(nkb, (c + 1) * UM_CHUNK);
          for (int kb = c * UM_CHUNK; kb < kb_end; ++kb, ++g) {
            const int s = g % LF_STAGES;
            mbar_wait(full0 + 8 * s, (g / LF_STAGES) & 1);
            tc_fence_after();
            const uint32_t st = base + s * LF_STAGE;
#pragma unroll
            for (int k = 0; k < UM_BK / 8; ++k) {
              const uint64_t a_hi = umma_desc_sw128(st + k * 32);
              const uint64_t a_lo = umma_desc_sw128(st + 16384 + k * 32);
              const uint64_t b_hi = umma_desc_sw128(st + 32768 + k * 32);
              const uint64_t b_lo = umma_desc_sw128(st + 49152 + k * 32);
              umma_tf32(tacc, a_hi, b_hi, idesc, (kb != c * UM_CHUNK) || (k != 0));
              umma_tf32(tacc, a_lo, b_hi, idesc, 1);
              umma_tf32(tacc, a_hi, b_lo, idesc, 1);
            }
            umma_commit(empty0 + 8 * s);
          }
          umma_commit(tfull0 + 8 * buf);
        }
      }
    }
  } else {
    const int q = warp & 3;                                     // TMEM lane quadrant of this warp
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    float* Srow = S + (q * 32 + lane) * LF_SLD;
    int cc = 0;
    int cj[4] = {-1, -1, -1, -1};                                // joint held by each transform slot
    int crow = -1;                                               // ... for this frame row
    float4 T0[4], T1[4], T2[4];
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int rt = t / a.nct, ct = t % a.nct;
      const int row = rt * UM_BM + q * 32 + lane;
      const bool rok = row < a.N;
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int c = 0; c < nchunk; ++c, ++cc) {
        const int buf = cc & 1;
        mbar_wait(tfull0 + 8 * buf, (cc >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          float tv[32];
          tmem_ld32(trow + buf * 128 + c0, tv);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += tv[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
      if (row != crow) { crow = row; cj[0] = cj[1] = cj[2] = cj[3] = -1; }
      const float* Arow = a.A + (size_t)(rok ? row : a.N - 1) * 624;
      const float* tr = a.trans + (size_t)(rok ? row : a.N - 1) * 3;
      const float t0 = tr[0], t1 = tr[1], t2 = tr[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 64; ++j) Srow[j] = acc[64 * h + j];
        const int v0 = ct * LF_VT + h * LF_VH;
        const int nvt = max(0, min(LF_VH, a.num_verts - v0));
        for (int i = 0; i < nvt; ++i) {
          const float px = Srow[3 * i], py = Srow[3 * i + 1], pz = Srow[3 * i + 2];
          const int* wi = a.fw_idx + (size_t)(v0 + i) * WK;
          const float* wv = a.fw_val + (size_t)(v0 + i) * WK;
          float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
          for (int k = 0; k < WK; ++k) {
            const float w = __ldg(wv + k);
            if (w != 0.f) {                                      // warp-uniform (weights depend on the vertex only)
              const int j12 = __ldg(wi + k);
              if (k < 4) {
                if (j12 != cj[k]) {
                  cj[k] = j12;
                  const float4* p = reinterpret_cast<const float4*>(Arow + j12);
                  T0[k] = p[0]; T1[k] = p[1]; T2[k] = p[2];
                }
                ox = fmaf(w, fmaf(T0[k].x, px, fmaf(T0[k].y, py, fmaf(T0[k].z, pz, T0[k].w))), ox);
                oy = fmaf(w, fmaf(T1[k].x, px, fmaf(T1[k].y, py, fmaf(T1[k].z, pz, T1[k].w))), oy);
                oz = fmaf(w, fmaf(T2[k].x, px, fmaf(T2[k].y, py, fmaf(T2[k].z, pz, T2[k].w))), oz);
              } else {
                const float4* p = reinterpret_cast<const float4*>(Arow + j12);
                const float4 r0 = p[0], r1 = p[1], r2 = p[2];
                ox = fmaf(w, fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))), ox);
                oy = fmaf(w, fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))), oy);
                oz = fmaf(w, fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))), oz);
              }
            }
          }
          Srow[3 * i] = ox + t0; Srow[3 * i + 1] = oy + t1; Srow[3 * i + 2] = oz + t2;
        }
        __syncwarp();
        const int nfl = nvt * 3;
        for (int rr = 0; rr < 32; ++rr) {
          const int r = rt * UM_BM + q * 32 + rr;
          if (r < a.N) {
            const float* src = S + (q * 32 + rr) * LF_SLD;
            float* dst = a.out + ((size_t)r * a.num_verts + v0) * 3;
            if (lane < nfl) __stcs(dst + lane, src[lane]);
            if (lane + 32 < nfl) __stcs(dst + lane + 32, src[lane + 32]);
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, 256u);
  }
}

}  // namespace hb
