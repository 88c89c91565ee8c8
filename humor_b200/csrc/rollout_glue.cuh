// One row of one HuMoR rollout step: everything between the decoder MLP output and the next
// step's network input (reference: models/humor_model.py:445-498 decode delta composition,
// :961-1001 canonicalisation + world transform inside roll_out, :696-772 apply_world2local_trans).
// Forward and hand-derived reverse mode; host/device so tests/host can check it against autograd.
//
// Row layouts (floats):
//   xin  [XIN_LD=416] : trans 0:3 | trans_vel 3:6 | R0 6:15 | ro_vel 15:18 | pose 18:207 (21x9)
//                       | joints 207:273 | joints_vel 273:339 | z 339:387 | zero pad
//   raw  [RAW_LD=224] : d_trans 0:3 | d_tvel 3:6 | d_ro(aa) 6:9 | d_rovel 9:12 | d_pose(aa) 12:75
//                       | d_joints 75:141 | d_jvel 141:207 | contact logits 207:216 | pad
//   world[WORLD_LD=348]: trans 3 | tvel 3 | R0 9 | rovel 3 | pose 189 | joints 66 | jvel 66 | contacts 9
//   G    [12]         : Gr (3x3 row-major, world->local rotation) | Gt (3)
#pragma once
#include "geom.cuh"

namespace hb {

constexpr int XIN_LD = 416;
constexpr int RAW_LD = 224;
constexpr int WORLD_LD = 348;
constexpr int STATE_D = 339;
constexpr int RAW_D = 216;
constexpr int NJ = 22;

HD void glue_step_fwd(const float* xin, const float* raw, const float* G, const float* t2j,
                      float* xnext, float* world, float* Gnext) {
  const float* Gr = G;
  const float* Gt = G + 9;
  float tr[3], tv[3], rv[3], R0[9], D[9], Ra[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    tr[i] = xin[i] + raw[i];
    tv[i] = xin[3 + i] + raw[3 + i];
    rv[i] = xin[15 + i] + raw[9 + i];
  }
  rodrigues_fwd(raw + 6, D);
  mat3_mul(D, xin + 6, R0);
  w2a_fwd(R0, Ra);
  const float ta[3] = {-tr[0], -tr[1], 0.f};
  // root orientation
  mat3_mul(Ra, R0, xnext + 6);
  mat3_mul_tn(Gr, R0, world + 6);
  // translation
  float u[3] = {tr[0] + ta[0], tr[1] + ta[1], tr[2] + ta[2]};
  mat3_vec(Ra, u, xnext + 0);
  float wt[3];
  mat3_tvec(Gr, tr, wt);
#pragma unroll
  for (int i = 0; i < 3; ++i) { wt[i] -= Gt[i]; world[i] = wt[i]; }
  // velocities
  mat3_vec(Ra, tv, xnext + 3);
  mat3_tvec(Gr, tv, world + 3);
  mat3_vec(Ra, rv, xnext + 15);
  mat3_tvec(Gr, rv, world + 15);
  // body joint rotations (frame independent)
  for (int j = 0; j < 21; ++j) {
    float Dj[9], Rj[9];
    rodrigues_fwd(raw + 12 + 3 * j, Dj);
    mat3_mul(Dj, xin + 18 + 9 * j, Rj);
#pragma unroll
    for (int e = 0; e < 9; ++e) { xnext[18 + 9 * j + e] = Rj[e]; world[18 + 9 * j + e] = Rj[e]; }
  }
  // joints and joint velocities
  for (int k = 0; k < NJ; ++k) {
    float p[3], v[3], a[3], o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p[i] = xin[207 + 3 * k + i] + raw[75 + 3 * k + i];
      v[i] = xin[273 + 3 * k + i] + raw[141 + 3 * k + i];
      a[i] = p[i] + ta[i] + t2j[i];
    }
    mat3_vec(Ra, a, o);
#pragma unroll
    for (int i = 0; i < 3; ++i) { xnext[207 + 3 * k + i] = o[i] - t2j[i]; a[i] = p[i] + t2j[i]; }
    mat3_tvec(Gr, a, o);
#pragma unroll
    for (int i = 0; i < 3; ++i) world[207 + 3 * k + i] = o[i] - t2j[i] - Gt[i];
    mat3_vec(Ra, v, xnext + 273 + 3 * k);
    mat3_tvec(Gr, v, world + 273 + 3 * k);
  }
#pragma unroll
  for (int c = 0; c < 9; ++c) world[339 + c] = raw[207 + c];
  // running world->local transform
  mat3_mul(Gr, Ra, Gnext);
  Gnext[9] = -wt[0]; Gnext[10] = -wt[1]; Gnext[11] = 0.f;
}

// helper: y = M x (or M^T x) backward.  dM += outer, dx += ...
HD void mv_bwd(const float* M, const float* x, const float* dy, float* dM, float* dx) {       // y = M x
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { dM[i * 3 + j] += dy[i] * x[j]; dx[j] += M[i * 3 + j] * dy[i]; }
}
HD void mtv_bwd(const float* M, const float* x, const float* dy, float* dM, float* dx) {      // y = M^T x
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { dM[j * 3 + i] += dy[i] * x[j]; dx[j] += M[j * 3 + i] * dy[i]; }
}

// Reverse of glue_step_fwd.
//   dn  [339]  grad wrt xnext[0:339]      dw [348] grad wrt world      dGn [12] grad wrt Gnext
// Outputs (overwritten): dxin[339] (direct/residual path into this step's past_in), draw[216], dG[12];
// dt2j[3] is accumulated (+=).
HD void glue_step_bwd(const float* xin, const float* raw, const float* G, const float* t2j,
                      const float* dn, const float* dw, const float* dGn,
                      float* dxin, float* draw, float* dG, float* dt2j) {
  const float* Gr = G;
  const float* Gt = G + 9;
  // ---- recompute the forward intermediates that are needed
  float tr[3], tv[3], rv[3], R0[9], D[9], Ra[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    tr[i] = xin[i] + raw[i];
    tv[i] = xin[3 + i] + raw[3 + i];
    rv[i] = xin[15 + i] + raw[9 + i];
  }
  rodrigues_fwd(raw + 6, D);
  mat3_mul(D, xin + 6, R0);
  w2a_fwd(R0, Ra);
  const float ta[3] = {-tr[0], -tr[1], 0.f};

  float dRa[9], dGr[9], dGt[3], dR0[9], dtr[3], dtv[3], drv[3], dta[3], d2j[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) { dRa[e] = 0.f; dGr[e] = 0.f; dR0[e] = 0.f; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { dGt[i] = 0.f; dtr[i] = 0.f; dtv[i] = 0.f; drv[i] = 0.f; dta[i] = 0.f; d2j[i] = 0.f; }

  // Gnext = [Gr Ra | -wt.xy, 0]
  mat3_mul_bwd(Gr, Ra, dGn, dGr, dRa);
  float dwt[3] = {dw[0] - dGn[9], dw[1] - dGn[10], dw[2]};
  // world.trans = Gr^T tr - Gt
  mtv_bwd(Gr, tr, dwt, dGr, dtr);
#pragma unroll
  for (int i = 0; i < 3; ++i) dGt[i] -= dwt[i];
  // world.R0 = Gr^T R0 :  dGr += R0 dW^T ; dR0 += Gr dW
  {
    const float* dW = dw + 6;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { a += R0[i * 3 + k] * dW[j * 3 + k]; b += Gr[i * 3 + k] * dW[k * 3 + j]; }
        dGr[i * 3 + j] += a;
        dR0[i * 3 + j] += b;
      }
  }
  // next.R0 = Ra R0
  mat3_mul_bwd(Ra, R0, dn + 6, dRa, dR0);
  // next.trans = Ra (tr + ta)
  {
    float u[3] = {tr[0] + ta[0], tr[1] + ta[1], tr[2] + ta[2]};
    float du[3] = {0.f, 0.f, 0.f};
    mv_bwd(Ra, u, dn + 0, dRa, du);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dtr[i] += du[i]; dta[i] += du[i]; }
  }
  // velocities
  mv_bwd(Ra, tv, dn + 3, dRa, dtv);
  mtv_bwd(Gr, tv, dw + 3, dGr, dtv);
  mv_bwd(Ra, rv, dn + 15, dRa, drv);
  mtv_bwd(Gr, rv, dw + 15, dGr, drv);
  // joints / joint velocities
  for (int k = 0; k < NJ; ++k) {
    float p[3], v[3], a[3], dp[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f}, du[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p[i] = xin[207 + 3 * k + i] + raw[75 + 3 * k + i];
      v[i] = xin[273 + 3 * k + i] + raw[141 + 3 * k + i];
      a[i] = p[i] + ta[i] + t2j[i];
    }
    const float* dnj = dn + 207 + 3 * k;
    const float* dwj = dw + 207 + 3 * k;
    mv_bwd(Ra, a, dnj, dRa, du);                       // next.joints = Ra a - t2j
#pragma unroll
    for (int i = 0; i < 3; ++i) { dp[i] += du[i]; dta[i] += du[i]; d2j[i] += du[i] - dnj[i]; du[i] = 0.f; a[i] = p[i] + t2j[i]; }
    mtv_bwd(Gr, a, dwj, dGr, du);                      // world.joints = Gr^T a - t2j - Gt
#pragma unroll
    for (int i = 0; i < 3; ++i) { dp[i] += du[i]; d2j[i] += du[i] - dwj[i]; dGt[i] -= dwj[i]; }
    mv_bwd(Ra, v, dn + 273 + 3 * k, dRa, dv);
    mtv_bwd(Gr, v, dw + 273 + 3 * k, dGr, dv);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dxin[207 + 3 * k + i] = dp[i]; draw[75 + 3 * k + i] = dp[i];
      dxin[273 + 3 * k + i] = dv[i]; draw[141 + 3 * k + i] = dv[i];
    }
  }
  // ta = (-tr.x, -tr.y, 0)
  dtr[0] -= dta[0];
  dtr[1] -= dta[1];
  // Ra = w2a(R0)
  w2a_bwd(R0, dRa, dR0);
  // R0 = D xin.R0 ; D = rodrigues(raw.d_ro)
  {
    float dD[9], dRin[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) { dD[e] = 0.f; dRin[e] = 0.f; }
    mat3_mul_bwd(D, xin + 6, dR0, dD, dRin);
    float daa[3] = {0.f, 0.f, 0.f};
    rodrigues_bwd(raw + 6, dD, daa);
#pragma unroll
    for (int e = 0; e < 9; ++e) dxin[6 + e] = dRin[e];
#pragma unroll
    for (int i = 0; i < 3; ++i) draw[6 + i] = daa[i];
  }
  // body rotations: Rj = Dj xin.Rj, appears in next.pose and world.pose
  for (int j = 0; j < 21; ++j) {
    float Dj[9], dRj[9], dD[9], dRin[9];
    rodrigues_fwd(raw + 12 + 3 * j, Dj);
#pragma unroll
    for (int e = 0; e < 9; ++e) { dRj[e] = dn[18 + 9 * j + e] + dw[18 + 9 * j + e]; dD[e] = 0.f; dRin[e] = 0.f; }
    mat3_mul_bwd(Dj, xin + 18 + 9 * j, dRj, dD, dRin);
    float daa[3] = {0.f, 0.f, 0.f};
    rodrigues_bwd(raw + 12 + 3 * j, dD, daa);
#pragma unroll
    for (int e = 0; e < 9; ++e) dxin[18 + 9 * j + e] = dRin[e];
#pragma unroll
    for (int i = 0; i < 3; ++i) draw[12 + 3 * j + i] = daa[i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dxin[i] = dtr[i];  draw[i] = dtr[i];
    dxin[3 + i] = dtv[i];  draw[3 + i] = dtv[i];
    dxin[15 + i] = drv[i]; draw[9 + i] = drv[i];
    dt2j[i] += d2j[i];
    dG[9 + i] = dGt[i];
  }
#pragma unroll
  for (int c = 0; c < 9; ++c) { draw[207 + c] = dw[339 + c]; dG[c] = dGr[c]; }
}

}  // namespace hb
