// Per-frame part of SMPL+H linear blend skinning: Rodrigues, pose feature, kinematic chain,
// skinning transforms.  Restates smplx==0.1.28 lbs()/batch_rigid_transform() as invoked by the
// reference (humor/body_model/body_model.py:78-91; algorithm in SURVEY.md Appendix A.1).
// The 30 hand joints are identity rotations on this path (pose_hand=None, flat_hand_mean=True,
// body_model.py:56-57,82-83): their world rotation is their wrist's, and pose-feature entries
// 189..458 are exactly zero, so only the 22 body rotations are evaluated.
// Host/device so tests/host can check forward and reverse mode against the torch oracle.
#pragma once
#include "geom.cuh"

namespace hb {

constexpr int LBS_J = 52;        // SMPL+H joints
constexpr int LBS_JB = 22;       // body joints (root + 21)
constexpr int LBS_KF = 208;      // feature row: betas 0:16 | pose feature 16:205 | pad
constexpr int LBS_NB = 16;

HD int lbs_rot_owner(int j) { return j < LBS_JB ? j : (j < 37 ? 20 : 21); }

// pose: [66] = root_orient(3) | pose_body(63).  Jrest: [52*3] rest joints of this frame's shape.
// Outputs: feat[16:205] (pose feature), A [52*12] (row-major 3x4 [R|t] mapping rest-space points),
//          Jp [52*3] posed joints (without translation).
HD void lbs_chain_fwd(const float* pose, const float* Jrest, const int* parents, float* featpose,
                      float* A, float* Jp) {
  float rot[LBS_JB * 9];
  float tw[LBS_J * 3];
  for (int j = 0; j < LBS_JB; ++j) {
    float R[9];
    rodrigues_fwd(pose + 3 * j, R);
    if (j > 0 && featpose) {
#pragma unroll
      for (int e = 0; e < 9; ++e) featpose[(j - 1) * 9 + e] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    }
    if (j == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) rot[e] = R[e];
      tw[0] = Jrest[0]; tw[1] = Jrest[1]; tw[2] = Jrest[2];
    } else {
      int p = parents[j];
      mat3_mul(rot + 9 * p, R, rot + 9 * j);
      float rel[3] = {Jrest[3 * j] - Jrest[3 * p], Jrest[3 * j + 1] - Jrest[3 * p + 1], Jrest[3 * j + 2] - Jrest[3 * p + 2]};
      float o[3];
      mat3_vec(rot + 9 * p, rel, o);
      tw[3 * j] = o[0] + tw[3 * p]; tw[3 * j + 1] = o[1] + tw[3 * p + 1]; tw[3 * j + 2] = o[2] + tw[3 * p + 2];
    }
  }
  for (int j = LBS_JB; j < LBS_J; ++j) {
    int p = parents[j];
    const float* rp = rot + 9 * lbs_rot_owner(p);
    float rel[3] = {Jrest[3 * j] - Jrest[3 * p], Jrest[3 * j + 1] - Jrest[3 * p + 1], Jrest[3 * j + 2] - Jrest[3 * p + 2]};
    float o[3];
    mat3_vec(rp, rel, o);
    tw[3 * j] = o[0] + tw[3 * p]; tw[3 * j + 1] = o[1] + tw[3 * p + 1]; tw[3 * j + 2] = o[2] + tw[3 * p + 2];
  }
  for (int j = 0; j < LBS_J; ++j) {
    const float* r = rot + 9 * lbs_rot_owner(j);
    float c[3];
    mat3_vec(r, Jrest + 3 * j, c);
    if (A) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        A[j * 12 + i * 4 + 0] = r[i * 3 + 0];
        A[j * 12 + i * 4 + 1] = r[i * 3 + 1];
        A[j * 12 + i * 4 + 2] = r[i * 3 + 2];
        A[j * 12 + i * 4 + 3] = tw[3 * j + i] - c[i];
      }
    }
    if (Jp) { Jp[3 * j] = tw[3 * j]; Jp[3 * j + 1] = tw[3 * j + 1]; Jp[3 * j + 2] = tw[3 * j + 2]; }
  }
}

// Reverse.  Inputs: dA [52*12] (may be null), dJp [52*3] (may be null), dfeatpose [189] (may be null).
// Outputs (overwritten): dpose [66], dJrest [52*3].
HD void lbs_chain_bwd(const float* pose, const float* Jrest, const int* parents, const float* dA,
                      const float* dJp, const float* dfeatpose, float* dpose, float* dJrest) {
  float rot[LBS_JB * 9];
  float Rl[LBS_JB * 9];
  for (int j = 0; j < LBS_JB; ++j) {
    rodrigues_fwd(pose + 3 * j, Rl + 9 * j);
    if (j == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) rot[e] = Rl[e];
    } else {
      mat3_mul(rot + 9 * parents[j], Rl + 9 * j, rot + 9 * j);
    }
  }
  float drot[LBS_JB * 9];
  float dt[LBS_J * 3];
  for (int e = 0; e < LBS_JB * 9; ++e) drot[e] = 0.f;
  for (int e = 0; e < LBS_J * 3; ++e) { dt[e] = 0.f; dJrest[e] = 0.f; }
  // A_j = [rot_o | tw_j - rot_o J_j],  Jp_j = tw_j
  for (int j = 0; j < LBS_J; ++j) {
    int o = lbs_rot_owner(j);
    float gt[3] = {0.f, 0.f, 0.f};
    if (dA) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        gt[i] = dA[j * 12 + i * 4 + 3];
#pragma unroll
        for (int k = 0; k < 3; ++k) drot[9 * o + i * 3 + k] += dA[j * 12 + i * 4 + k] - gt[i] * Jrest[3 * j + k];
      }
      float back[3];
      mat3_tvec(rot + 9 * o, gt, back);
      dJrest[3 * j] -= back[0]; dJrest[3 * j + 1] -= back[1]; dJrest[3 * j + 2] -= back[2];
    }
    if (dJp) { gt[0] += dJp[3 * j]; gt[1] += dJp[3 * j + 1]; gt[2] += dJp[3 * j + 2]; }
    dt[3 * j] += gt[0]; dt[3 * j + 1] += gt[1]; dt[3 * j + 2] += gt[2];
  }
  // tw_j = rot_{o(p)} (J_j - J_p) + tw_p      (children always have larger index than parents)
  for (int j = LBS_J - 1; j >= 1; --j) {
    int p = parents[j];
    int op = lbs_rot_owner(p);
    const float* g = dt + 3 * j;
    float rel[3] = {Jrest[3 * j] - Jrest[3 * p], Jrest[3 * j + 1] - Jrest[3 * p + 1], Jrest[3 * j + 2] - Jrest[3 * p + 2]};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) drot[9 * op + i * 3 + k] += g[i] * rel[k];
    float back[3];
    mat3_tvec(rot + 9 * op, g, back);
#pragma unroll
    for (int i = 0; i < 3; ++i) { dJrest[3 * j + i] += back[i]; dJrest[3 * p + i] -= back[i]; dt[3 * p + i] += g[i]; }
    if (j < LBS_JB) {
      // rot_j = rot_p R_j
      float dR[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) dR[e] = 0.f;
      mat3_mul_bwd(rot + 9 * p, Rl + 9 * j, drot + 9 * j, drot + 9 * p, dR);
      if (dfeatpose) {
#pragma unroll
        for (int e = 0; e < 9; ++e) dR[e] += dfeatpose[(j - 1) * 9 + e];
      }
      float d3[3] = {0.f, 0.f, 0.f};
      rodrigues_bwd(pose + 3 * j, dR, d3);
      dpose[3 * j] = d3[0]; dpose[3 * j + 1] = d3[1]; dpose[3 * j + 2] = d3[2];
    }
  }
  // root: rot_0 = R_0, tw_0 = J_0
  dJrest[0] += dt[0]; dJrest[1] += dt[1]; dJrest[2] += dt[2];
  float d3[3] = {0.f, 0.f, 0.f};
  rodrigues_bwd(pose, drot, d3);
  dpose[0] = d3[0]; dpose[1] = d3[1]; dpose[2] = d3[2];
}

}  // namespace hb
