// One row of one rollout step handled by ONE WARP (lane = joint): the step's glue between the decoder MLP output and the
// next step's network input (reference: models/humor_model.py:445-498 decode delta composition, :961-1001 canonicalisation
// + world transform inside roll_out, :696-772 apply_world2local_trans), forward and hand-derived reverse.  Same arithmetic
// as glue_step_fwd / glue_step_bwd in rollout_glue.cuh (the scalar statement tests/host validates against autograd); rows are
// staged through shared memory with coalesced loads/stores, the root-level reverse-mode sums are warp-shuffle reductions.
//
// Callers: glue_fwd_kernel / glue_bwd_kernel (one launch per step, rollout.cu) and the persistent decoder-chain kernel
// (chain_persist.cuh), which runs them from its epilogue warps between GEMM phases.  CG = true: every step-varying input is
// read with ld.global.cg (L2) because another CTA of the SAME launch produced it.
#pragma once
#include "rollout_glue.cuh"
#include "umma_split16.cuh"

namespace hb {

#ifdef HB_HOST_SHIM
template <bool CG> static inline float glue_ld(const float* p) { return *p; }
#else
template <bool CG> __device__ __forceinline__ float glue_ld(const float* p) { return CG ? __ldcg(p) : *p; }
#endif

__device__ __forceinline__ float glue_hi11(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// v as (hi, lo) planes when a lo plane exists, else exactly
__device__ __forceinline__ void glue_put(float* hi, float* lo, int i, float v) {
  if (lo) { const float h = glue_hi11(v); hi[i] = h; lo[i] = v - h; } else { hi[i] = v; }
}
__device__ __forceinline__ float glue_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int GLUE_FWD_SMEM = 340 + 216 + 340 + 348;              // floats per warp
constexpr int GLUE_BWD_SMEM = 340 + 216 + 340 + 348 + 340 + 224 + 16;   // + the step's G when the tape rows are prefetched

struct GlueFwdRow {                 // every pointer addresses THIS row
  const float* xr;                  // xin [339]
  const float* rr;                  // raw [216]
  const float* G;                   // [12] world->local transform of the step
  const float* t2j;                 // [3]
  const float* zt;                  // z of step t+1 [48], nullptr on the last step
  float* xn;                        // next xin row [XIN_LD] (fp32 tape)
  float* xn_hi; float* xn_lo;       // its hi/lo planes (nullable)
  float* wo;                        // world row [WORLD_LD]
  float* gn;                        // next G [12]
  float* h1; float* h1_lo; float* h2; float* h2_lo; float* h3; float* h3_lo;   // z skip columns of the hidden-activation rows (48 each)
  // forward chain on fp16 hi / scaled-lo planes: next step's input row [>= 387 halves] and the z skip columns (nullable: fp32 planes above)
  unsigned short *xn16_h, *xn16_l, *h1_16h, *h1_16l, *h2_16h, *h2_16l, *h3_16h, *h3_16l;
};
__device__ __forceinline__ void glue_put16(unsigned short* h, unsigned short* l, int i, float v) { split16(v, h[i], l[i]); }

template <bool CG>
__device__ __forceinline__ void glue_fwd_warp(const GlueFwdRow& io, int lane, float* sx, float* sr, float* sn, float* sw) {
  for (int i = lane; i < STATE_D; i += 32) sx[i] = glue_ld<CG>(io.xr + i);
  for (int i = lane; i < RAW_D; i += 32) sr[i] = glue_ld<CG>(io.rr + i);
  __syncwarp();
  float Gr[9], Gt[3], t2j[3], tr[3], R0[9], D[9], Ra[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) Gr[e] = glue_ld<CG>(io.G + e);
#pragma unroll
  for (int i = 0; i < 3; ++i) { Gt[i] = glue_ld<CG>(io.G + 9 + i); t2j[i] = io.t2j[i]; tr[i] = sx[i] + sr[i]; }
  rodrigues_fwd(sr + 6, D);
  mat3_mul(D, sx + 6, R0);
  w2a_fwd(R0, Ra);
  const float ta[3] = {-tr[0], -tr[1], 0.f};
  if (lane < NJ) {
    const int k = lane;
    float p[3], v[3], a[3], o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p[i] = sx[207 + 3 * k + i] + sr[75 + 3 * k + i];
      v[i] = sx[273 + 3 * k + i] + sr[141 + 3 * k + i];
      a[i] = p[i] + ta[i] + t2j[i];
    }
    mat3_vec(Ra, a, o);
#pragma unroll
    for (int i = 0; i < 3; ++i) { sn[207 + 3 * k + i] = o[i] - t2j[i]; a[i] = p[i] + t2j[i]; }
    mat3_tvec(Gr, a, o);
#pragma unroll
    for (int i = 0; i < 3; ++i) sw[207 + 3 * k + i] = o[i] - t2j[i] - Gt[i];
    mat3_vec(Ra, v, sn + 273 + 3 * k);
    mat3_tvec(Gr, v, sw + 273 + 3 * k);
    if (k == 0) {
      mat3_mul(Ra, R0, sn + 6);
      mat3_mul_tn(Gr, R0, sw + 6);
    } else {
      const int j = k - 1;
      float Dj[9], Rj[9];
      rodrigues_fwd(sr + 12 + 3 * j, Dj);
      mat3_mul(Dj, sx + 18 + 9 * j, Rj);
#pragma unroll
      for (int e = 0; e < 9; ++e) { sn[18 + 9 * j + e] = Rj[e]; sw[18 + 9 * j + e] = Rj[e]; }
    }
  } else if (lane == 22) {
    float u[3] = {tr[0] + ta[0], tr[1] + ta[1], tr[2] + ta[2]};
    mat3_vec(Ra, u, sn + 0);
    float wt[3];
    mat3_tvec(Gr, tr, wt);
    float gn[12];
    mat3_mul(Gr, Ra, gn);
#pragma unroll
    for (int i = 0; i < 3; ++i) { wt[i] -= Gt[i]; sw[i] = wt[i]; }
    gn[9] = -wt[0]; gn[10] = -wt[1]; gn[11] = 0.f;
#pragma unroll
    for (int e = 0; e < 12; ++e) io.gn[e] = gn[e];
  } else if (lane == 23) {
    float tv[3] = {sx[3] + sr[3], sx[4] + sr[4], sx[5] + sr[5]};
    mat3_vec(Ra, tv, sn + 3);
    mat3_tvec(Gr, tv, sw + 3);
  } else if (lane == 24) {
    float rv[3] = {sx[15] + sr[9], sx[16] + sr[10], sx[17] + sr[11]};
    mat3_vec(Ra, rv, sn + 15);
    mat3_tvec(Gr, rv, sw + 15);
  } else if (lane == 25) {
#pragma unroll
    for (int c = 0; c < 9; ++c) sw[339 + c] = sr[207 + c];
  }
  __syncwarp();
  for (int i = lane; i < STATE_D; i += 32) {
    io.xn[i] = sn[i];
    if (io.xn_lo) glue_put(io.xn_hi, io.xn_lo, i, sn[i]);
    if (io.xn16_h && io.zt) glue_put16(io.xn16_h, io.xn16_l, i, sn[i]);
  }
  for (int i = lane; i < WORLD_LD; i += 32) io.wo[i] = sw[i];
  if (io.zt) {
    for (int i = lane; i < 48; i += 32) {
      const float v = io.zt[i];
      io.xn[STATE_D + i] = v;
      if (io.xn_lo) glue_put(io.xn_hi, io.xn_lo, STATE_D + i, v);
      if (io.xn16_h) {
        glue_put16(io.xn16_h, io.xn16_l, STATE_D + i, v);
        glue_put16(io.h1_16h, io.h1_16l, i, v); glue_put16(io.h2_16h, io.h2_16l, i, v); glue_put16(io.h3_16h, io.h3_16l, i, v);
      } else {
        glue_put(io.h1, io.h1_lo, i, v);
        glue_put(io.h2, io.h2_lo, i, v);
        glue_put(io.h3, io.h3_lo, i, v);
      }
    }
  }
}

struct GlueBwdRow {                 // every pointer addresses THIS row
  const float* xr;                  // xin [339] of the step
  const float* rr;                  // raw [216]
  const float* wr;                  // d world [WORLD_LD]
  const float* G;                   // [12]
  const float* t2j;                 // [3]
  int have_next;                    // gradients from step t+1 exist
  int staged;                       // 1: xin / raw / d world rows are in the staging arrays already (prefetched), do not load them
  const float* a0;                  // d xin of step t+1 from the decoder's first layer [>= 339 (+48 z when dzt)]
  const float* px;                  // d xin of step t+1 from the prior [352]
  float* xs;                        // in: d xin residual path of step t+1 [340]; out: that of this step
  const float* dGn;                 // [12] d G of step t+1
  float* dG;                        // [12] out
  float* dt2j;                      // [3] accumulated
  float* dzt;                       // d z of step t+1 [48] (nullable: the persistent chain computes d z as one batched GEMM)
  const float* dh1; const float* dh1_lo; const float* dh2; const float* dh2_lo; const float* dh3; const float* dh3_lo;   // z skip columns (dzt only)
  float* draw;                      // fp32 d raw row [RAW_LD] (nullable)
  float* draw_hi; float* draw_lo;   // its hi/lo planes (nullable)
};

template <bool CG>
__device__ __forceinline__ void glue_bwd_warp(const GlueBwdRow& io, int lane, float* sx, float* sr, float* dn, float* dw, float* dx,
                                              float* dr) {
  if (!io.staged) {
    for (int i = lane; i < STATE_D; i += 32) sx[i] = io.xr[i];          // forward tape: written by an earlier launch
    for (int i = lane; i < RAW_D; i += 32) sr[i] = io.rr[i];
    for (int i = lane; i < WORLD_LD; i += 32) dw[i] = io.wr[i];
  }
  if (io.have_next) {
    for (int i = lane; i < STATE_D; i += 32) dn[i] = glue_ld<CG>(io.xs + i) + glue_ld<CG>(io.a0 + i) + io.px[i];
    if (io.dzt)
      for (int i = lane; i < 48; i += 32) {
        auto tail = [&](const float* hi, const float* lo) { return lo ? hi[i] + lo[i] : hi[i]; };
        io.dzt[i] = io.a0[STATE_D + i] + tail(io.dh1, io.dh1_lo) + tail(io.dh2, io.dh2_lo) + tail(io.dh3, io.dh3_lo);
      }
  } else {
    for (int i = lane; i < STATE_D; i += 32) dn[i] = 0.f;
  }
  __syncwarp();
  float Gr[9], Gt[3], t2j[3], tr[3], R0[9], D[9], Ra[9], dGn[12];
#pragma unroll
  for (int e = 0; e < 9; ++e) Gr[e] = io.G[e];
#pragma unroll
  for (int i = 0; i < 3; ++i) { Gt[i] = io.G[9 + i]; t2j[i] = io.t2j[i]; tr[i] = sx[i] + sr[i]; }
#pragma unroll
  for (int i = 0; i < 12; ++i) dGn[i] = io.have_next ? glue_ld<CG>(io.dGn + i) : 0.f;
  rodrigues_fwd(sr + 6, D);
  mat3_mul(D, sx + 6, R0);
  w2a_fwd(R0, Ra);
  const float ta[3] = {-tr[0], -tr[1], 0.f};
  // per-lane partial sums of the root-level adjoints
  float dRa[9], dGr[9], dGt[3] = {0.f, 0.f, 0.f}, dta[3] = {0.f, 0.f, 0.f}, d2j[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 9; ++e) { dRa[e] = 0.f; dGr[e] = 0.f; }
  float dR0[9], dtr[3] = {0.f, 0.f, 0.f};            // lane 25 / lane 22 private
#pragma unroll
  for (int e = 0; e < 9; ++e) dR0[e] = 0.f;

  if (lane < NJ) {
    const int k = lane;
    float p[3], v[3], a[3], dp[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f}, du[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p[i] = sx[207 + 3 * k + i] + sr[75 + 3 * k + i];
      v[i] = sx[273 + 3 * k + i] + sr[141 + 3 * k + i];
      a[i] = p[i] + ta[i] + t2j[i];
    }
    const float* dnj = dn + 207 + 3 * k;
    const float* dwj = dw + 207 + 3 * k;
    mv_bwd(Ra, a, dnj, dRa, du);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dp[i] += du[i]; dta[i] += du[i]; d2j[i] += du[i] - dnj[i]; du[i] = 0.f; a[i] = p[i] + t2j[i]; }
    mtv_bwd(Gr, a, dwj, dGr, du);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dp[i] += du[i]; d2j[i] += du[i] - dwj[i]; dGt[i] -= dwj[i]; }
    mv_bwd(Ra, v, dn + 273 + 3 * k, dRa, dv);
    mtv_bwd(Gr, v, dw + 273 + 3 * k, dGr, dv);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dx[207 + 3 * k + i] = dp[i]; dr[75 + 3 * k + i] = dp[i];
      dx[273 + 3 * k + i] = dv[i]; dr[141 + 3 * k + i] = dv[i];
    }
    if (k > 0) {
      const int j = k - 1;
      float Dj[9], dRj[9], dD[9], dRin[9];
      rodrigues_fwd(sr + 12 + 3 * j, Dj);
#pragma unroll
      for (int e = 0; e < 9; ++e) { dRj[e] = dn[18 + 9 * j + e] + dw[18 + 9 * j + e]; dD[e] = 0.f; dRin[e] = 0.f; }
      mat3_mul_bwd(Dj, sx + 18 + 9 * j, dRj, dD, dRin);
      float daa[3] = {0.f, 0.f, 0.f};
      rodrigues_bwd(sr + 12 + 3 * j, dD, daa);
#pragma unroll
      for (int e = 0; e < 9; ++e) dx[18 + 9 * j + e] = dRin[e];
#pragma unroll
      for (int i = 0; i < 3; ++i) dr[12 + 3 * j + i] = daa[i];
    }
  } else if (lane == 22) {
    float dwt[3] = {dw[0] - dGn[9], dw[1] - dGn[10], dw[2]};
    mtv_bwd(Gr, tr, dwt, dGr, dtr);
#pragma unroll
    for (int i = 0; i < 3; ++i) dGt[i] -= dwt[i];
    float u[3] = {tr[0] + ta[0], tr[1] + ta[1], tr[2] + ta[2]};
    float du[3] = {0.f, 0.f, 0.f};
    mv_bwd(Ra, u, dn + 0, dRa, du);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dtr[i] += du[i]; dta[i] += du[i]; }
  } else if (lane == 23) {
    float tv[3] = {sx[3] + sr[3], sx[4] + sr[4], sx[5] + sr[5]}, dtv[3] = {0.f, 0.f, 0.f};
    mv_bwd(Ra, tv, dn + 3, dRa, dtv);
    mtv_bwd(Gr, tv, dw + 3, dGr, dtv);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dx[3 + i] = dtv[i]; dr[3 + i] = dtv[i]; }
  } else if (lane == 24) {
    float rv[3] = {sx[15] + sr[9], sx[16] + sr[10], sx[17] + sr[11]}, drv[3] = {0.f, 0.f, 0.f};
    mv_bwd(Ra, rv, dn + 15, dRa, drv);
    mtv_bwd(Gr, rv, dw + 15, dGr, drv);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dx[15 + i] = drv[i]; dr[9 + i] = drv[i]; }
  } else if (lane == 25) {
    mat3_mul_bwd(Gr, Ra, dGn, dGr, dRa);                 // Gnext = Gr Ra
    const float* dW = dw + 6;                            // world.R0 = Gr^T R0
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float a = 0.f, bb = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { a += R0[i * 3 + k] * dW[j * 3 + k]; bb += Gr[i * 3 + k] * dW[k * 3 + j]; }
        dGr[i * 3 + j] += a;
        dR0[i * 3 + j] += bb;
      }
    mat3_mul_bwd(Ra, R0, dn + 6, dRa, dR0);              // next.R0 = Ra R0
#pragma unroll
    for (int c = 0; c < 9; ++c) dr[207 + c] = dw[339 + c];
    for (int c = RAW_D; c < RAW_LD; ++c) dr[c] = 0.f;
  }
  // ---- warp totals (every lane receives them)
#pragma unroll
  for (int e = 0; e < 9; ++e) { dRa[e] = glue_warp_sum(dRa[e]); dGr[e] = glue_warp_sum(dGr[e]); }
#pragma unroll
  for (int i = 0; i < 3; ++i) { dGt[i] = glue_warp_sum(dGt[i]); dta[i] = glue_warp_sum(dta[i]); d2j[i] = glue_warp_sum(d2j[i]); }
  if (lane == 25) {
    w2a_bwd(R0, dRa, dR0);                               // Ra = w2a(R0)
    float dD[9], dRin[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) { dD[e] = 0.f; dRin[e] = 0.f; }
    mat3_mul_bwd(D, sx + 6, dR0, dD, dRin);              // R0 = D xin.R0
    float daa[3] = {0.f, 0.f, 0.f};
    rodrigues_bwd(sr + 6, dD, daa);
#pragma unroll
    for (int e = 0; e < 9; ++e) dx[6 + e] = dRin[e];
#pragma unroll
    for (int i = 0; i < 3; ++i) dr[6 + i] = daa[i];
  } else if (lane == 22) {
    dtr[0] -= dta[0];                                    // ta = (-tr.x, -tr.y, 0)
    dtr[1] -= dta[1];
#pragma unroll
    for (int i = 0; i < 3; ++i) { dx[i] = dtr[i]; dr[i] = dtr[i]; }
  } else if (lane == 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) io.dG[e] = dGr[e];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      io.dG[9 + i] = dGt[i];
      io.dt2j[i] = (io.have_next ? glue_ld<CG>(io.dt2j + i) : 0.f) + d2j[i];
    }
  }
  __syncwarp();
  for (int i = lane; i < STATE_D; i += 32) io.xs[i] = dx[i];
  for (int i = lane; i < RAW_LD; i += 32) {
    if (io.draw) io.draw[i] = dr[i];
    if (io.draw_lo) glue_put(io.draw_hi, io.draw_lo, i, dr[i]);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// TWO warps per row (persistent chain, when the grid has at least two epilogue warps per sub-sequence): the serial instruction
// stream of one warp is the cost of the glue (3 300 instructions in reverse, ~8.6 clk each with one warp per scheduler), so the
// row is split by data dependence, not by lanes:
//   role 1 ("J")  the 21 body-joint rotations R_j = D_j(raw) xin.R_j - Rodrigues + 3x3 products and their reverse - touch no
//                 root-level quantity: independent of everything role 0 does
//   role 0 ("R")  root frame (D, R0, Ra), joint positions / velocities (lanes 0..21), root velocities as two more "velocity items"
//                 (lanes 22, 23: same code path as the joint velocities, no extra divergent branch), translation + running
//                 transform (lane 24), root-orientation chain (lane 25), the root-level reductions and their serial tail
// Both roles load the row into the shared staging arrays together (64 lanes) and store the results together; `pair_sync` is a
// named barrier of the two warps.  Same arithmetic per element as the one-warp functions above; only the order of the
// root-level warp sums over lanes is unchanged too (role 0 holds every summand).
// ------------------------------------------------------------------------------------------------------------------------
#ifdef HB_HOST_SHIM
static inline void glue_pair_sync(int pi) { shim::sync_pair(pi); }
#else
__device__ __forceinline__ void glue_pair_sync(int pi) { asm volatile("bar.sync %0, 64;" ::"r"(2 + pi) : "memory"); }
#endif

template <bool CG>
__device__ __forceinline__ void glue_fwd_pair(const GlueFwdRow& io, int role, int lane, int pi, float* sx, float* sr, float* sn, float* sw) {
  const int l64 = role * 32 + lane;
  const float zv = (io.zt && l64 < 48) ? io.zt[l64] : 0.f;      // z of the next step: an HBM first touch, requested before anything else
  for (int i = l64; i < STATE_D; i += 64) sx[i] = glue_ld<CG>(io.xr + i);
  for (int i = l64; i < RAW_D; i += 64) sr[i] = glue_ld<CG>(io.rr + i);
  glue_pair_sync(pi);
  if (role == 1) {
    if (lane >= 1 && lane < NJ) {
      const int j = lane - 1;
      float Dj[9], Rj[9];
      rodrigues_fwd(sr + 12 + 3 * j, Dj);
      mat3_mul(Dj, sx + 18 + 9 * j, Rj);
#pragma unroll
      for (int e = 0; e < 9; ++e) { sn[18 + 9 * j + e] = Rj[e]; sw[18 + 9 * j + e] = Rj[e]; }
    }
  } else {
    float Gr[9], Gt[3], t2j[3], tr[3], R0[9], D[9], Ra[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) Gr[e] = glue_ld<CG>(io.G + e);
#pragma unroll
    for (int i = 0; i < 3; ++i) { Gt[i] = glue_ld<CG>(io.G + 9 + i); t2j[i] = io.t2j[i]; tr[i] = sx[i] + sr[i]; }
    rodrigues_fwd(sr + 6, D);
    mat3_mul(D, sx + 6, R0);
    w2a_fwd(R0, Ra);
    const float ta[3] = {-tr[0], -tr[1], 0.f};
    if (lane < 24) {
      // velocity items: joint velocities (lanes 0..21), root translation velocity (22), root angular velocity (23)
      const int xo = lane < NJ ? 273 + 3 * lane : (lane == 22 ? 3 : 15);
      const int ro = lane < NJ ? 141 + 3 * lane : (lane == 22 ? 3 : 9);
      float v[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = sx[xo + i] + sr[ro + i];
      mat3_vec(Ra, v, sn + xo);
      mat3_tvec(Gr, v, sw + xo);
      if (lane < NJ) {
        const int k = lane;
        float p[3], a[3], o[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { p[i] = sx[207 + 3 * k + i] + sr[75 + 3 * k + i]; a[i] = p[i] + ta[i] + t2j[i]; }
        mat3_vec(Ra, a, o);
#pragma unroll
        for (int i = 0; i < 3; ++i) { sn[207 + 3 * k + i] = o[i] - t2j[i]; a[i] = p[i] + t2j[i]; }
        mat3_tvec(Gr, a, o);
#pragma unroll
        for (int i = 0; i < 3; ++i) sw[207 + 3 * k + i] = o[i] - t2j[i] - Gt[i];
        if (k == 0) {
          mat3_mul(Ra, R0, sn + 6);
          mat3_mul_tn(Gr, R0, sw + 6);
        }
      }
    } else if (lane == 24) {
      float u[3] = {tr[0] + ta[0], tr[1] + ta[1], tr[2] + ta[2]};
      mat3_vec(Ra, u, sn + 0);
      float wt[3];
      mat3_tvec(Gr, tr, wt);
      float gn[12];
      mat3_mul(Gr, Ra, gn);
#pragma unroll
      for (int i = 0; i < 3; ++i) { wt[i] -= Gt[i]; sw[i] = wt[i]; }
      gn[9] = -wt[0]; gn[10] = -wt[1]; gn[11] = 0.f;
#pragma unroll
      for (int e = 0; e < 12; ++e) io.gn[e] = gn[e];
    } else if (lane == 25) {
#pragma unroll
      for (int c = 0; c < 9; ++c) sw[339 + c] = sr[207 + c];
    }
  }
  glue_pair_sync(pi);
  for (int i = l64; i < STATE_D; i += 64) {
    io.xn[i] = sn[i];
    if (io.xn_lo) glue_put(io.xn_hi, io.xn_lo, i, sn[i]);
    if (io.xn16_h && io.zt) glue_put16(io.xn16_h, io.xn16_l, i, sn[i]);
  }
  for (int i = l64; i < WORLD_LD; i += 64) io.wo[i] = sw[i];
  if (io.zt) {
    for (int i = l64; i < 48; i += 64) {
      const float v = zv;
      io.xn[STATE_D + i] = v;
      if (io.xn_lo) glue_put(io.xn_hi, io.xn_lo, STATE_D + i, v);
      if (io.xn16_h) {
        glue_put16(io.xn16_h, io.xn16_l, STATE_D + i, v);
        glue_put16(io.h1_16h, io.h1_16l, i, v); glue_put16(io.h2_16h, io.h2_16l, i, v); glue_put16(io.h3_16h, io.h3_16l, i, v);
      } else {
        glue_put(io.h1, io.h1_lo, i, v);
        glue_put(io.h2, io.h2_lo, i, v);
        glue_put(io.h3, io.h3_lo, i, v);
      }
    }
  }
}

template <bool CG>
__device__ __forceinline__ void glue_bwd_pair(const GlueBwdRow& io, int role, int lane, int pi, float* sx, float* sr, float* dn, float* dw,
                                              float* dx, float* dr) {
  const int l64 = role * 32 + lane;
  if (!io.staged) {
    for (int i = l64; i < STATE_D; i += 64) sx[i] = io.xr[i];
    for (int i = l64; i < RAW_D; i += 64) sr[i] = io.rr[i];
    for (int i = l64; i < WORLD_LD; i += 64) dw[i] = io.wr[i];
  }
  if (io.have_next) {
    for (int i = l64; i < STATE_D; i += 64) dn[i] = glue_ld<CG>(io.xs + i) + glue_ld<CG>(io.a0 + i) + io.px[i];
  } else {
    for (int i = l64; i < STATE_D; i += 64) dn[i] = 0.f;
  }
  glue_pair_sync(pi);
  if (role == 1) {
    if (lane >= 1 && lane < NJ) {
      const int j = lane - 1;
      float Dj[9], dRj[9], dD[9], dRin[9];
      rodrigues_fwd(sr + 12 + 3 * j, Dj);
#pragma unroll
      for (int e = 0; e < 9; ++e) { dRj[e] = dn[18 + 9 * j + e] + dw[18 + 9 * j + e]; dD[e] = 0.f; dRin[e] = 0.f; }
      mat3_mul_bwd(Dj, sx + 18 + 9 * j, dRj, dD, dRin);
      float daa[3] = {0.f, 0.f, 0.f};
      rodrigues_bwd(sr + 12 + 3 * j, dD, daa);
#pragma unroll
      for (int e = 0; e < 9; ++e) dx[18 + 9 * j + e] = dRin[e];
#pragma unroll
      for (int i = 0; i < 3; ++i) dr[12 + 3 * j + i] = daa[i];
    }
  } else {
    float Gr[9], Gt[3], t2j[3], tr[3], R0[9], D[9], Ra[9], dGn[12];
#pragma unroll
    for (int e = 0; e < 9; ++e) Gr[e] = io.G[e];
#pragma unroll
    for (int i = 0; i < 3; ++i) { Gt[i] = io.G[9 + i]; t2j[i] = io.t2j[i]; tr[i] = sx[i] + sr[i]; }
#pragma unroll
    for (int i = 0; i < 12; ++i) dGn[i] = io.have_next ? glue_ld<CG>(io.dGn + i) : 0.f;
    rodrigues_fwd(sr + 6, D);
    mat3_mul(D, sx + 6, R0);
    w2a_fwd(R0, Ra);
    const float ta[3] = {-tr[0], -tr[1], 0.f};
    float dRa[9], dGr[9], dGt[3] = {0.f, 0.f, 0.f}, dta[3] = {0.f, 0.f, 0.f}, d2j[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 9; ++e) { dRa[e] = 0.f; dGr[e] = 0.f; }
    float dR0[9], dtr[3] = {0.f, 0.f, 0.f};            // lane 25 / lane 24 private
#pragma unroll
    for (int e = 0; e < 9; ++e) dR0[e] = 0.f;
    (void)Gt;
    if (lane < 24) {
      // velocity items (joint velocities 0..21, root translation velocity 22, root angular velocity 23)
      const int xo = lane < NJ ? 273 + 3 * lane : (lane == 22 ? 3 : 15);
      const int ro = lane < NJ ? 141 + 3 * lane : (lane == 22 ? 3 : 9);
      float v[3], dv[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = sx[xo + i] + sr[ro + i];
      if (lane < NJ) {
        const int k = lane;
        float p[3], a[3], dp[3] = {0.f, 0.f, 0.f}, du[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 3; ++i) { p[i] = sx[207 + 3 * k + i] + sr[75 + 3 * k + i]; a[i] = p[i] + ta[i] + t2j[i]; }
        const float* dnj = dn + 207 + 3 * k;
        const float* dwj = dw + 207 + 3 * k;
        mv_bwd(Ra, a, dnj, dRa, du);
#pragma unroll
        for (int i = 0; i < 3; ++i) { dp[i] += du[i]; dta[i] += du[i]; d2j[i] += du[i] - dnj[i]; du[i] = 0.f; a[i] = p[i] + t2j[i]; }
        mtv_bwd(Gr, a, dwj, dGr, du);
#pragma unroll
        for (int i = 0; i < 3; ++i) { dp[i] += du[i]; d2j[i] += du[i] - dwj[i]; dGt[i] -= dwj[i]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { dx[207 + 3 * k + i] = dp[i]; dr[75 + 3 * k + i] = dp[i]; }
      }
      mv_bwd(Ra, v, dn + xo, dRa, dv);
      mtv_bwd(Gr, v, dw + xo, dGr, dv);
#pragma unroll
      for (int i = 0; i < 3; ++i) { dx[xo + i] = dv[i]; dr[ro + i] = dv[i]; }
    } else if (lane == 24) {
      float dwt[3] = {dw[0] - dGn[9], dw[1] - dGn[10], dw[2]};
      mtv_bwd(Gr, tr, dwt, dGr, dtr);
#pragma unroll
      for (int i = 0; i < 3; ++i) dGt[i] -= dwt[i];
      float u[3] = {tr[0] + ta[0], tr[1] + ta[1], tr[2] + ta[2]};
      float du[3] = {0.f, 0.f, 0.f};
      mv_bwd(Ra, u, dn + 0, dRa, du);
#pragma unroll
      for (int i = 0; i < 3; ++i) { dtr[i] += du[i]; dta[i] += du[i]; }
    } else if (lane == 25) {
      mat3_mul_bwd(Gr, Ra, dGn, dGr, dRa);                 // Gnext = Gr Ra
      const float* dW = dw + 6;                            // world.R0 = Gr^T R0
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float a = 0.f, bb = 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k) { a += R0[i * 3 + k] * dW[j * 3 + k]; bb += Gr[i * 3 + k] * dW[k * 3 + j]; }
          dGr[i * 3 + j] += a;
          dR0[i * 3 + j] += bb;
        }
      mat3_mul_bwd(Ra, R0, dn + 6, dRa, dR0);              // next.R0 = Ra R0
#pragma unroll
      for (int c = 0; c < 9; ++c) dr[207 + c] = dw[339 + c];
      for (int c = RAW_D; c < RAW_LD; ++c) dr[c] = 0.f;
    }
    // ---- warp totals (every lane receives them)
#pragma unroll
    for (int e = 0; e < 9; ++e) { dRa[e] = glue_warp_sum(dRa[e]); dGr[e] = glue_warp_sum(dGr[e]); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { dGt[i] = glue_warp_sum(dGt[i]); dta[i] = glue_warp_sum(dta[i]); d2j[i] = glue_warp_sum(d2j[i]); }
    if (lane == 25) {
      w2a_bwd(R0, dRa, dR0);                               // Ra = w2a(R0)
      float dD[9], dRin[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) { dD[e] = 0.f; dRin[e] = 0.f; }
      mat3_mul_bwd(D, sx + 6, dR0, dD, dRin);              // R0 = D xin.R0
      float daa[3] = {0.f, 0.f, 0.f};
      rodrigues_bwd(sr + 6, dD, daa);
#pragma unroll
      for (int e = 0; e < 9; ++e) dx[6 + e] = dRin[e];
#pragma unroll
      for (int i = 0; i < 3; ++i) dr[6 + i] = daa[i];
    } else if (lane == 24) {
      dtr[0] -= dta[0];                                    // ta = (-tr.x, -tr.y, 0)
      dtr[1] -= dta[1];
#pragma unroll
      for (int i = 0; i < 3; ++i) { dx[i] = dtr[i]; dr[i] = dtr[i]; }
    } else if (lane == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) io.dG[e] = dGr[e];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        io.dG[9 + i] = dGt[i];
        io.dt2j[i] = (io.have_next ? glue_ld<CG>(io.dt2j + i) : 0.f) + d2j[i];
      }
    }
  }
  glue_pair_sync(pi);
  for (int i = l64; i < STATE_D; i += 64) io.xs[i] = dx[i];
  for (int i = l64; i < RAW_LD; i += 64) {
    if (io.draw) io.draw[i] = dr[i];
    if (io.draw_lo) glue_put(io.draw_hi, io.draw_lo, i, dr[i]);
  }
}

}  // namespace hb

