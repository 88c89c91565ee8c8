// Fused Stage-III energies and their gradients (reference: humor/fitting/fitting_loss.py
// root_fit :94-181, smpl_fit :183-224, motion_fit :226-309 and the per-term functions :317-518;
// projection humor/fitting/fitting_utils.py:647-676, gmof :250-258).
//
// One block per (sequence b, frame t).  Every gradient element has exactly one owner thread, so the
// outputs are written once (no atomics) and the term values go through a fixed-order reduction:
// results are bit-reproducible run to run.  A term whose coefficient is 0 contributes exactly zero
// (the reference gates on `weight > 0`, fitting_loss.py:193,240,277).
#include "common.cuh"
#include "../../include/humor_b200.h"

namespace hb {

__constant__ int c_smpl2op[25] = {52, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62};
// body_model/utils.py:9 SMPL_PARENTS (the reference's table, shoulders -> neck), used by bone_length_loss
__constant__ int c_parents_ref[22] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 12, 12, 13, 14, 16, 17, 18, 19};
__constant__ int c_nchild[22] = {3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 3, 1, 1, 0, 1, 1, 1, 1, 0, 0};
__constant__ int c_child[22][3] = {{1, 2, 3}, {4, 0, 0}, {5, 0, 0}, {6, 0, 0}, {7, 0, 0}, {8, 0, 0}, {9, 0, 0}, {10, 0, 0},
                                   {11, 0, 0}, {12, 0, 0}, {0, 0, 0}, {0, 0, 0}, {13, 14, 15}, {16, 0, 0}, {17, 0, 0},
                                   {0, 0, 0}, {18, 0, 0}, {19, 0, 0}, {20, 0, 0}, {21, 0, 0}, {0, 0, 0}, {0, 0, 0}};
// datasets/amass_utils.py:22-23 CONTACT_INDS -> slot in the 9 contact logits (-1: not a contact joint)
__constant__ int c_contact_slot[22] = {0, -1, -1, -1, 1, 2, -1, 3, 4, -1, 5, 6, -1, -1, -1, -1, -1, -1, -1, -1, 7, 8};

constexpr int NT = HB_NUM_TERMS;
constexpr int LB = 128;  // block size

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}
__device__ __forceinline__ bool vis(float x) { return !isinf(x); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float bone_len(const float* J, int j) {
  const int p = c_parents_ref[j];
  const float dx = J[3 * j] - J[3 * p], dy = J[3 * j + 1] - J[3 * p + 1], dz = J[3 * j + 2] - J[3 * p + 2];
  return sqrtf(dx * dx + dy * dy + dz * dz);
}
// d(0.5 sum_t (bl_{t+1}-bl_t)^2)/d bl_t for bone j at frame t
__device__ __forceinline__ float bone_dl(const float* Jr_bt, int T, int t, int j) {
  const float l = bone_len(Jr_bt, j);
  float g = 0.f;
  if (t > 0) g += l - bone_len(Jr_bt - 66, j);
  if (t + 1 < T) g -= bone_len(Jr_bt + 66, j) - l;
  return g;
}

__global__ void __launch_bounds__(LB) fit_losses_kernel(HbFitArgs a) {
  __shared__ float red[4];
  const int b = blockIdx.x / a.T, t = blockIdx.x % a.T;
  const int tid = threadIdx.x;
  const int T = a.T, B = a.B, njx = a.njx;
  const size_t bt = (size_t)b * T + t;
  const size_t obt = (size_t)b * a.T_obs + t;
  float v_j2d = 0.f, v_j3d = 0.f, v_smooth = 0.f, v_v3d = 0.f, v_ovp = 0.f, v_ovv = 0.f;
  float v_jc = 0.f, v_bl = 0.f, v_j3r = 0.f, v_cv = 0.f, v_ch = 0.f, v_mp = 0.f, v_pp = 0.f;
  float v_sp = 0.f, v_ovb = 0.f, v_fr = 0.f, v_ovf = 0.f;
  const float* cf = a.coef;

  // ---------------------------------------------------------------- camera-frame joints (owner: joint j)
  if (tid < njx) {
    const int j = tid;
    const float* P = a.cam_joints + (bt * njx + j) * 3;
    const float px = P[0], py = P[1], pz = P[2];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (cf[HB_T_JOINTS2D] != 0.f && a.obs_joints2d) {
      int k = -1;
#pragma unroll
      for (int q = 0; q < 25; ++q) k = (c_smpl2op[q] == j) ? q : k;
      if (k >= 0) {
        const float* o = a.obs_joints2d + (obt * 25 + k) * 3;
        float conf = (k == 1 || k == 9 || k == 12) ? 0.f : o[2];      // OP_IGNORE_JOINTS, fitting_loss.py:350-352
        const float fx = a.cam_f[b * 2], fy = a.cam_f[b * 2 + 1];
        const float iz = 1.f / pz;
        const float rx = px * iz * fx + a.cam_c[b * 2] - o[0];
        const float ry = py * iz * fy + a.cam_c[b * 2 + 1] - o[1];
        const float s2 = a.sigma2d * a.sigma2d, c2 = conf * conf;
        const float dx_ = s2 + rx * rx, dy_ = s2 + ry * ry;
        v_j2d = c2 * (s2 * rx * rx / dx_ + s2 * ry * ry / dy_);
        const float grx = c2 * 2.f * rx * s2 * s2 / (dx_ * dx_) * cf[HB_T_JOINTS2D];
        const float gry = c2 * 2.f * ry * s2 * s2 / (dy_ * dy_) * cf[HB_T_JOINTS2D];
        gx += grx * fx * iz;
        gy += gry * fy * iz;
        gz -= (grx * fx * px + gry * fy * py) * iz * iz;
      }
    }
    if (j < 22) {
      if (cf[HB_T_JOINTS3D] != 0.f && a.obs_joints3d) {
        const float* o = a.obs_joints3d + (obt * 22 + j) * 3;
        const float w = cf[HB_T_JOINTS3D];
        if (vis(o[0])) { float d = px - o[0]; v_j3d += 0.5f * d * d; gx += w * d; }
        if (vis(o[1])) { float d = py - o[1]; v_j3d += 0.5f * d * d; gy += w * d; }
        if (vis(o[2])) { float d = pz - o[2]; v_j3d += 0.5f * d * d; gz += w * d; }
      }
      if (cf[HB_T_SMOOTH] != 0.f) {
        const float w = cf[HB_T_SMOOTH];
        if (t + 1 < T) {
          const float* Q = P + (size_t)njx * 3;
          const float d0 = Q[0] - px, d1 = Q[1] - py, d2 = Q[2] - pz;
          v_smooth += 0.5f * (d0 * d0 + d1 * d1 + d2 * d2);
          gx -= w * d0; gy -= w * d1; gz -= w * d2;
        }
        if (t > 0) {
          const float* Q = P - (size_t)njx * 3;
          gx += w * (px - Q[0]); gy += w * (py - Q[1]); gz += w * (pz - Q[2]);
        }
      }
    }
    float* g = a.d_cam_joints + (bt * njx + j) * 3;
    g[0] = gx; g[1] = gy; g[2] = gz;
  }

  // ---------------------------------------------------------------- camera-frame key vertices (owner: vertex k)
  if (tid < 43) {
    const int k = tid;
    const float* V = a.cam_verts + (bt * 43 + k) * 3;
    float g[3] = {0.f, 0.f, 0.f};
    if (cf[HB_T_VERTS3D] != 0.f && a.obs_verts3d) {
      const float* o = a.obs_verts3d + (obt * 43 + k) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (vis(o[c])) { float d = V[c] - o[c]; v_v3d += 0.5f * d * d; g[c] += cf[HB_T_VERTS3D] * d; }
    }
    if (cf[HB_T_OV_POS] != 0.f && a.seq_interval) {
      const float w = cf[HB_T_OV_POS];
      // role 'c': head of sequence b against the tail of b-1   (fitting_loss.py:136-157)
      if (b > 0) {
        const int ov = a.seq_interval[(b - 1) * 2 + 1] - a.seq_interval[b * 2];
        if (t < ov) {
          const int ta = T - ov + t;
          const float* Aa = a.cam_verts + (((size_t)(b - 1) * T + ta) * 43 + k) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float d = Aa[c] - V[c];                        // d_i = a_i - c_i
            float gd = d;
            v_ovp += 0.5f * d * d;
            if (ov > 1) {
              if (t + 1 < ov) { const float dn = Aa[129 + c] - V[129 + c]; v_ovv += 0.5f * (dn - d) * (dn - d); gd -= dn - d; }
              if (t > 0) { const float dp = Aa[c - 129] - V[c - 129]; gd += d - dp; }
            }
            g[c] -= w * gd;
          }
        }
      }
      // role 'a': tail of sequence b against the head of b+1
      if (b + 1 < B) {
        const int ov = a.seq_interval[b * 2 + 1] - a.seq_interval[(b + 1) * 2];
        const int i = t - (T - ov);
        if (i >= 0 && i < ov) {
          const float* Cc = a.cam_verts + (((size_t)(b + 1) * T + i) * 43 + k) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float d = V[c] - Cc[c];
            float gd = d;
            if (ov > 1) {
              if (i + 1 < ov) { const float dn = V[129 + c] - Cc[129 + c]; gd -= dn - d; }
              if (i > 0) { const float dp = V[c - 129] - Cc[c - 129]; gd += d - dp; }
            }
            g[c] += w * gd;
          }
        }
      }
    }
    float* o = a.d_cam_verts + (bt * 43 + k) * 3;
    o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
  }

  // ---------------------------------------------------------------- prior-frame SMPL joints / rollout joints
  if (tid >= 64 && tid < 64 + 22) {
    const int j = tid - 64;
    const float* Jp = a.prior_joints + (bt * 22 + j) * 3;
    const float* JrF = a.roll_joints + bt * 66;           // this frame's 22 joints
    const float* Jr = JrF + 3 * j;
    float gp[3] = {0.f, 0.f, 0.f}, gr[3] = {0.f, 0.f, 0.f};
    if (cf[HB_T_JOINT_CONSIST] != 0.f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = Jp[c] - Jr[c];
        v_jc += 0.5f * d * d;
        gp[c] += cf[HB_T_JOINT_CONSIST] * d;
        gr[c] -= cf[HB_T_JOINT_CONSIST] * d;
      }
    }
    if (cf[HB_T_BONE_LEN] != 0.f) {
      const float w = cf[HB_T_BONE_LEN];
      if (j > 0) {
        const float l = bone_len(JrF, j);
        if (t + 1 < T) { const float d = bone_len(JrF + 66, j) - l; v_bl += 0.5f * d * d; }
        const float gl = w * bone_dl(JrF, T, t, j) / l;
        const int p = c_parents_ref[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) gr[c] += gl * (Jr[c] - JrF[3 * p + c]);
      }
      for (int q = 0; q < c_nchild[j]; ++q) {
        const int ch = c_child[j][q];
        const float l = bone_len(JrF, ch);
        const float gl = w * bone_dl(JrF, T, t, ch) / l;
#pragma unroll
        for (int c = 0; c < 3; ++c) gr[c] -= gl * (JrF[3 * ch + c] - Jr[c]);
      }
    }
    if (cf[HB_T_J3D_ROLLOUT] != 0.f && a.obs_joints3d) {
      const float* o = a.obs_joints3d + (obt * 22 + j) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (vis(o[c])) { const float d = Jr[c] - o[c]; v_j3r += 0.5f * d * d; gr[c] += cf[HB_T_J3D_ROLLOUT] * d; }
    }
    // contacts (fitting_loss.py:450-469; conf = sigmoid(logits) scattered to CONTACT_INDS, first frame
    // repeated: motion_optimizer.py:982-998)
    const int slot = c_contact_slot[j];
    if (slot >= 0 && T > 1) {
      const float* lg = a.contact_logits + ((size_t)b * (T - 1)) * 9 + slot;
      const float conf = sigmoidf_(lg[(size_t)(t > 0 ? t - 1 : 0) * 9]);
      float dconf = 0.f;                         // d loss / d conf of the logit this thread owns (t>=1)
      if (cf[HB_T_CONTACT_VEL] != 0.f) {
        const float w = cf[HB_T_CONTACT_VEL];
        if (t > 0) {
          const float d0 = Jp[0] - Jp[-66], d1 = Jp[1] - Jp[-65], d2 = Jp[2] - Jp[-64];
          const float n2 = d0 * d0 + d1 * d1 + d2 * d2;
          v_cv += 0.5f * n2 * conf;
          dconf += w * 0.5f * n2;
          gp[0] += w * conf * d0; gp[1] += w * conf * d1; gp[2] += w * conf * d2;
        }
        if (t + 1 < T) {
          const float cn = sigmoidf_(lg[(size_t)t * 9]);
          gp[0] -= w * cn * (Jp[66] - Jp[0]); gp[1] -= w * cn * (Jp[67] - Jp[1]); gp[2] -= w * cn * (Jp[68] - Jp[2]);
        }
      }
      if (cf[HB_T_CONTACT_H] != 0.f) {
        const float w = cf[HB_T_CONTACT_H];
        const float z = Jp[2], az = fabsf(z), r = fmaxf(az - 0.08f, 0.f);
        v_ch += r * conf;
        if (az > 0.08f) gp[2] += w * conf * (z > 0.f ? 1.f : -1.f);
        dconf += w * r;
        if (t == 1) {                            // frame 0 shares logit 0
          const float z0 = Jp[2 - 66];
          dconf += w * fmaxf(fabsf(z0) - 0.08f, 0.f);
        }
      }
      if (t > 0) a.d_contact_logits[((size_t)b * (T - 1) + (t - 1)) * 9 + slot] = dconf * conf * (1.f - conf);
    }
    float* o = a.d_prior_joints + (bt * 22 + j) * 3;
    o[0] = gp[0]; o[1] = gp[1]; o[2] = gp[2];
    o = a.d_roll_joints + (bt * 22 + j) * 3;
    o[0] = gr[0]; o[1] = gr[1]; o[2] = gr[2];
  }

  // ---------------------------------------------------------------- motion prior (owner: latent dim)
  if (tid < 48 && t + 1 < T) {
    const size_t zi = ((size_t)b * (T - 1) + t) * 48 + tid;
    const float z = a.z[zi];
    const float w = cf[HB_T_MOTION_PRIOR];
    if (a.prior_out) {
      const size_t pi = ((size_t)t * B + b) * 96 + tid;
      const float m = a.prior_out[pi], lv = a.prior_out[pi + 48];
      const float var = expf(lv), d = z - m;
      // -log N(z; m, var), fitting_loss.py:404-414,504-518
      v_mp = logf(sqrtf(var)) + 0.91893853320467274178f + d * d / (2.f * var);
      a.d_z[zi] = w * d / var;
      a.d_prior_out[pi] = -w * d / var;
      a.d_prior_out[pi + 48] = w * (0.5f - d * d / (2.f * var));
    } else {
      v_mp = z * z;
      a.d_z[zi] = w * 2.f * z;
    }
  }
  // ---------------------------------------------------------------- pose prior on the re-encoded latent pose
  if (a.latent_pose && tid >= 96 && tid < 128) {
    const size_t li = bt * 32 + (tid - 96);
    const float x = a.latent_pose[li];
    v_pp = x * x;
    a.d_latent_pose[li] = cf[HB_T_POSE_PRIOR] * 2.f * x;
  }
  // ---------------------------------------------------------------- per-sequence terms (frame 0 block)
  if (t == 0) {
    if (tid >= 48 && tid < 64) {
      const int l = tid - 48;
      const float x = a.betas[b * 16 + l];
      float g = 0.f;
      v_sp = x * x;
      g += cf[HB_T_SHAPE_PRIOR] * 2.f * x;
      if (cf[HB_T_OV_BETAS] != 0.f && a.seq_interval) {
        if (b > 0) { const float d = a.betas[(b - 1) * 16 + l] - x; v_ovb = 0.5f * d * d; g -= cf[HB_T_OV_BETAS] * d; }
        if (b + 1 < B) { const float d = x - a.betas[(b + 1) * 16 + l]; g += cf[HB_T_OV_BETAS] * d; }
      }
      a.d_betas[b * 16 + l] = g;
    }
    if (a.floor && tid >= 86 && tid < 89) {
      const int c = tid - 86;
      const float x = a.floor[b * 3 + c];
      float g = 0.f;
      if (cf[HB_T_FLOOR_REG] != 0.f && a.obs_floor) {
        const float d = x - a.obs_floor[b * 4 + c] * a.obs_floor[b * 4 + 3];
        v_fr = 0.5f * d * d;
        g += cf[HB_T_FLOOR_REG] * d;
      }
      if (cf[HB_T_OV_FLOOR] != 0.f && a.seq_interval) {
        if (b > 0) { const float d = a.floor[(b - 1) * 3 + c] - x; v_ovf = 0.5f * d * d; g -= cf[HB_T_OV_FLOOR] * d; }
        if (b + 1 < B) { const float d = x - a.floor[(b + 1) * 3 + c]; g += cf[HB_T_OV_FLOOR] * d; }
      }
      a.d_floor[b * 3 + c] = g;
    }
  }

  // ---------------------------------------------------------------- fixed-order block reduction of the terms
  float* out = a.partials + bt * NT;
  float s;
#define HB_RED(val, id) s = block_sum(val, red); if (tid == 0) out[id] = s;
  HB_RED(v_j2d, HB_T_JOINTS2D) HB_RED(v_j3d, HB_T_JOINTS3D) HB_RED(v_v3d, HB_T_VERTS3D) HB_RED(v_ovp, HB_T_OV_POS)
  HB_RED(v_ovv, HB_T_OV_VEL) HB_RED(v_pp, HB_T_POSE_PRIOR) HB_RED(v_sp, HB_T_SHAPE_PRIOR) HB_RED(v_smooth, HB_T_SMOOTH)
  HB_RED(v_ovb, HB_T_OV_BETAS) HB_RED(v_mp, HB_T_MOTION_PRIOR) HB_RED(v_jc, HB_T_JOINT_CONSIST) HB_RED(v_bl, HB_T_BONE_LEN)
  HB_RED(v_j3r, HB_T_J3D_ROLLOUT) HB_RED(v_cv, HB_T_CONTACT_VEL) HB_RED(v_ch, HB_T_CONTACT_H) HB_RED(v_fr, HB_T_FLOOR_REG)
  HB_RED(v_ovf, HB_T_OV_FLOOR)
#undef HB_RED
  if (tid == 0) { out[HB_T_INIT_PRIOR] = 0.f; for (int i = HB_T_OV_FLOOR + 1; i < NT; ++i) out[i] = 0.f; }
}

// Deterministic two-level reduction of the per-row terms (fixed order at both levels).  Level 1: FIT_RB blocks, each sums its slice of
// rows with lane = term (a row's 24 terms are one coalesced 96-byte read) into partials[(rows + block)][NT]; level 2: one warp
// sums the block sums and forms the weighted loss.  (Round 1: one block, warp = term reading with a 96-byte lane stride:
// 57-240 us on the step's critical path, profiles/r02j_profile_step.txt.)
constexpr int FIT_RB = 64;
__global__ void __launch_bounds__(256) fit_reduce1_kernel(HbFitArgs a) {
  __shared__ float part[8][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = a.B * a.T;
  const int r0 = (int)((long long)rows * blockIdx.x / FIT_RB), r1 = (int)((long long)rows * (blockIdx.x + 1) / FIT_RB);
  float s0 = 0.f, s1 = 0.f;
  if (lane < NT) {
    int r = r0 + w;
    for (; r + 8 < r1; r += 16) { s0 += a.partials[(size_t)r * NT + lane]; s1 += a.partials[(size_t)(r + 8) * NT + lane]; }
    if (r < r1) s0 += a.partials[(size_t)r * NT + lane];
  }
  part[w][lane] = s0 + s1;
  __syncthreads();
  if (w == 0 && lane < NT) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += part[q][lane];
    a.partials[((size_t)rows + blockIdx.x) * NT + lane] = s;
  }
}
__global__ void fit_reduce_kernel(HbFitArgs a) {
  __shared__ float tot[NT];
  const int lane = threadIdx.x;
  const int rows = a.B * a.T;
  if (lane < NT) {
    float s = 0.f;
    for (int q = 0; q < FIT_RB; ++q) s += a.partials[((size_t)rows + q) * NT + lane];
    tot[lane] = s; a.terms[lane] = s;
  }
  __syncwarp();
  if (lane == 0) {
    float l = 0.f;
    for (int i = 0; i < NT; ++i) l += a.coef[i] * tot[i];
    a.loss[0] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// GMM negative log-likelihood (init_motion_prior_loss, fitting_loss.py:504-518) and its gradient.
// One block per GMM_RT = 2 rows.  For every component the lower-triangular Linv_k (D x D, 76 KB at D = 138) is
// staged in shared memory with coalesced loads and serves the block's rows in both passes (round 1 read it row-per-lane
// straight from global memory, one block per row: 684 us per call in situ, profiles/r02b_profile_step.txt):
//   pass 1   y_k = Linv_k (x - mu_k) (4 warps per row, lane = output row i, four partial sums per dot product), maha_k = |y_k|^2,
//            y kept in shared memory
//   LSE      per row over the K components -> nll, responsibilities
//   pass 2   d nll / dx = sum_k resp_k Linv_k^T y_k (thread = column, conflict-free column reads; 8 row accumulators)
// ------------------------------------------------------------------------------------------------
constexpr int GMM_MAXD = 160, GMM_MAXK = 32;
constexpr int GMM_RT = 2;                           // rows per block: 128 blocks at B = 256 (the 8-row form left 116 SMs idle: 323 us)
constexpr int GMM_WPR = 8 / GMM_RT;                 // warps per row in pass 1
static inline size_t gmm_smem_bytes(int D, int K) { return ((size_t)D * D + (size_t)K * GMM_RT * D + (size_t)2 * GMM_RT * D) * sizeof(float); }
__global__ void __launch_bounds__(256) gmm_nll_kernel(int B, int D, int K, const float* __restrict__ x, const float* __restrict__ logw,
                                                       const float* __restrict__ mean, const float* __restrict__ Linv,
                                                       const float* __restrict__ logdet, float* nll, float* dx) {
  HB_DYN_SMEM_F32(sm);
  float* Ls = sm;                                   // [D][D]
  float* ys = Ls + (size_t)D * D;                   // [K][GMM_RT][D]
  float* xs = ys + (size_t)K * GMM_RT * D;          // [GMM_RT][D]
  float* ds = xs + (size_t)GMM_RT * D;              // [GMM_RT][D]  x - mu_k
  __shared__ float lp[GMM_RT][GMM_MAXK];
  __shared__ float resp[GMM_RT][GMM_MAXK];
  __shared__ float mpart[8];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * GMM_RT;
  const int nr = min(GMM_RT, B - b0);
  const int row = w / GMM_WPR, sub = w - row * GMM_WPR;
  for (int i = tid; i < GMM_RT * D; i += 256) { const int r = i / D; xs[i] = r < nr ? x[(size_t)(b0 + r) * D + (i - r * D)] : 0.f; }
  // pass 1
  for (int k = 0; k < K; ++k) {
    __syncthreads();                                // xs ready (k = 0) / previous component's reads of Ls, ds done
    const float* L = Linv + (size_t)k * D * D;
    for (int i = tid; i < D * D; i += 256) Ls[i] = L[i];
    const float* mu = mean + (size_t)k * D;
    for (int i = tid; i < GMM_RT * D; i += 256) { const int r = i / D; ds[i] = xs[i] - mu[i - r * D]; }
    __syncthreads();
    const float* d = ds + row * D;
    float maha = 0.f;
    for (int i = lane + 32 * sub; i < D; i += 32 * GMM_WPR) {
      const float* Li = Ls + (size_t)i * D;
      float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;  // four partial sums: the dot product is not one dependent chain
      int j = 0;
      for (; j + 3 <= i; j += 4) {
        y0 = fmaf(Li[j], d[j], y0); y1 = fmaf(Li[j + 1], d[j + 1], y1);
        y2 = fmaf(Li[j + 2], d[j + 2], y2); y3 = fmaf(Li[j + 3], d[j + 3], y3);
      }
      for (; j <= i; ++j) y0 = fmaf(Li[j], d[j], y0);
      const float y = (y0 + y1) + (y2 + y3);
      ys[((size_t)k * GMM_RT + row) * D + i] = y;
      maha += y * y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maha += __shfl_xor_sync(0xffffffffu, maha, o);
    if (lane == 0) mpart[w] = maha;
    __syncthreads();
    if (tid < GMM_RT) {
      float m = 0.f;
      for (int q = 0; q < GMM_WPR; ++q) m += mpart[tid * GMM_WPR + q];
      lp[tid][k] = logw[k] - 0.5f * ((float)D * 1.8378770664093453f + m) - logdet[k];
    }
  }
  __syncthreads();
  if (tid < GMM_RT) {                               // one thread per row
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, lp[tid][k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(lp[tid][k] - mx);
    const float lse = mx + logf(se);
    if (tid < nr) nll[b0 + tid] = -lse;
    for (int k = 0; k < K; ++k) resp[tid][k] = expf(lp[tid][k] - lse);
  }
  // pass 2: thread tid < D owns column tid of every row of the block
  float g[GMM_RT];
#pragma unroll
  for (int r = 0; r < GMM_RT; ++r) g[r] = 0.f;
  for (int k = 0; k < K; ++k) {
    __syncthreads();                                // resp ready (k = 0) / previous component's reads of Ls done
    const float* L = Linv + (size_t)k * D * D;
    for (int i = tid; i < D * D; i += 256) Ls[i] = L[i];
    __syncthreads();
    if (tid < D) {
      float acc[GMM_RT][2];
#pragma unroll
      for (int r = 0; r < GMM_RT; ++r) acc[r][0] = acc[r][1] = 0.f;
      int i = tid;
      for (; i + 1 < D; i += 2) {
        const float l0 = Ls[(size_t)i * D + tid], l1 = Ls[(size_t)(i + 1) * D + tid];
#pragma unroll
        for (int r = 0; r < GMM_RT; ++r) {
          acc[r][0] = fmaf(l0, ys[((size_t)k * GMM_RT + r) * D + i], acc[r][0]);
          acc[r][1] = fmaf(l1, ys[((size_t)k * GMM_RT + r) * D + i + 1], acc[r][1]);
        }
      }
      if (i < D) {
        const float l0 = Ls[(size_t)i * D + tid];
#pragma unroll
        for (int r = 0; r < GMM_RT; ++r) acc[r][0] = fmaf(l0, ys[((size_t)k * GMM_RT + r) * D + i], acc[r][0]);
      }
#pragma unroll
      for (int r = 0; r < GMM_RT; ++r) g[r] = fmaf(resp[r][k], acc[r][0] + acc[r][1], g[r]);
    }
  }
  if (tid < D)
    for (int r = 0; r < nr; ++r) dx[(size_t)(b0 + r) * D + tid] = g[r];
}

// ------------------------------------------------------------------------------------------------
// The same in three launches over (component, chunk of GMB_ROWS rows) blocks: a block stages ONE Linv_k once and serves 32 rows,
// instead of every 2-row block streaming all K factors twice (128 blocks x 1.8 MB through 24 block-wide syncs: 0.22 ms at B = 256,
// D = 138, K = 12 on the B200).  Scratch in the caller's workspace: y [K][B][D], log-probabilities lp [B][K], per-component
// gradients [K][B][D]; the component sums run in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int GMB_ROWS = 32;
static inline size_t gmb_smem_bytes(int D) { return ((size_t)D * (D + 1) + (size_t)GMB_ROWS * D) * sizeof(float); }
// Linv_k [D][D] -> shared memory rows of LD floats.  Four 16-byte loads in flight per thread: one element per iteration left every
// load's ~1 us of latency exposed (75 dependent round trips per block: 0.10 ms of the first pass, profiles/r02r_profile_step.txt).
__device__ __forceinline__ void gmb_stage_factor(const float* __restrict__ L, float* Ls, int D, int LD, int tid) {
  const int n = D * D;
  if ((n & 3) == 0) {
    const float4* L4 = reinterpret_cast<const float4*>(L);
    const int n4 = n >> 2;
    for (int i0 = tid; i0 < n4; i0 += 4 * 256) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int i4 = i0 + u * 256; v[u] = i4 < n4 ? __ldg(L4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i4 = i0 + u * 256;
        if (i4 < n4) {
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { const int idx = 4 * i4 + q, r = idx / D; Ls[r * LD + (idx - r * D)] = e[q]; }
        }
      }
    }
  } else {
    for (int i = tid; i < n; i += 256) { const int r = i / D; Ls[r * LD + (i - r * D)] = L[i]; }
  }
}
__global__ void __launch_bounds__(256) gmm_pass1_kernel(int B, int D, int K, const float* __restrict__ x, const float* __restrict__ logw,
                                                         const float* __restrict__ mean, const float* __restrict__ Linv,
                                                         const float* __restrict__ logdet, float* __restrict__ y, float* __restrict__ lp) {
  HB_DYN_SMEM_F32(sm);
  const int LD = D + 1;                             // odd row stride: lane = row reads of a column chunk are conflict-free
  float* Ls = sm;                                   // [D][LD]
  float* ds = Ls + (size_t)D * LD;                  // [GMB_ROWS][D]  x - mu_k
  const int k = blockIdx.x, b0 = blockIdx.y * GMB_ROWS, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int nr = min(GMB_ROWS, B - b0);
  gmb_stage_factor(Linv + (size_t)k * D * D, Ls, D, LD, tid);
  const float* mu = mean + (size_t)k * D;
  for (int i = tid; i < GMB_ROWS * D; i += 256) { const int r = i / D, c = i - r * D; ds[i] = r < nr ? x[(size_t)(b0 + r) * D + c] - mu[c] : 0.f; }
  __syncthreads();
  for (int r = w; r < nr; r += 8) {                 // a warp per row, lane = output index i
    const float* d = ds + r * D;
    float maha = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float* Li = Ls + (size_t)i * LD;
      float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
      int j = 0;
      for (; j + 3 <= i; j += 4) {
        y0 = fmaf(Li[j], d[j], y0); y1 = fmaf(Li[j + 1], d[j + 1], y1);
        y2 = fmaf(Li[j + 2], d[j + 2], y2); y3 = fmaf(Li[j + 3], d[j + 3], y3);
      }
      for (; j <= i; ++j) y0 = fmaf(Li[j], d[j], y0);
      const float yv = (y0 + y1) + (y2 + y3);
      y[((size_t)k * B + b0 + r) * D + i] = yv;
      maha += yv * yv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maha += __shfl_xor_sync(0xffffffffu, maha, o);
    if (lane == 0) lp[(size_t)(b0 + r) * K + k] = logw[k] - 0.5f * ((float)D * 1.8378770664093453f + maha) - logdet[k];
  }
}
__global__ void __launch_bounds__(256) gmm_pass2_kernel(int B, int D, int K, const float* __restrict__ Linv, const float* __restrict__ y,
                                                         const float* __restrict__ lp, float* __restrict__ dxk) {
  HB_DYN_SMEM_F32(sm);
  const int LD = D + 1;
  float* Ls = sm;                                   // [D][LD]
  float* ys = Ls + (size_t)D * LD;                  // [GMB_ROWS][D]  resp_k * y_k
  __shared__ float resp[GMB_ROWS];
  const int k = blockIdx.x, b0 = blockIdx.y * GMB_ROWS, tid = threadIdx.x;
  const int nr = min(GMB_ROWS, B - b0);
  gmb_stage_factor(Linv + (size_t)k * D * D, Ls, D, LD, tid);
  if (tid < nr) {                                   // responsibility of component k for every row of the chunk
    const float* l = lp + (size_t)(b0 + tid) * K;
    float mx = -INFINITY;
    for (int q = 0; q < K; ++q) mx = fmaxf(mx, l[q]);
    float se = 0.f;
    for (int q = 0; q < K; ++q) se += expf(l[q] - mx);
    resp[tid] = expf(l[k] - (mx + logf(se)));
  }
  __syncthreads();
  for (int i = tid; i < GMB_ROWS * D; i += 256) { const int r = i / D; ys[i] = r < nr ? resp[r] * y[((size_t)k * B + b0 + r) * D + (i - r * D)] : 0.f; }
  __syncthreads();
  // d x[r][c] = sum_{i >= c} L[i][c] (resp y)[r][i]: thread = (column c, half of the rows)
  const int c = tid % 128, half = tid / 128;
  for (int cc = c; cc < D; cc += 128) {
    float acc[GMB_ROWS / 2];
#pragma unroll
    for (int r = 0; r < GMB_ROWS / 2; ++r) acc[r] = 0.f;
    for (int i = cc; i < D; ++i) {
      const float l = Ls[(size_t)i * LD + cc];
#pragma unroll
      for (int r = 0; r < GMB_ROWS / 2; ++r) acc[r] = fmaf(l, ys[(half * (GMB_ROWS / 2) + r) * D + i], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < GMB_ROWS / 2; ++r) {
      const int rr = half * (GMB_ROWS / 2) + r;
      if (rr < nr) dxk[((size_t)k * B + b0 + rr) * D + cc] = acc[r];
    }
  }
}
__global__ void gmm_reduce_kernel(int B, int D, int K, const float* __restrict__ lp, const float* __restrict__ dxk, float* nll, float* dx) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B * D) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += dxk[(size_t)k * B * D + idx];
    dx[idx] = s;
  }
  if (idx < B) {
    const float* l = lp + (size_t)idx * K;
    float mx = -INFINITY;
    for (int q = 0; q < K; ++q) mx = fmaxf(mx, l[q]);
    float se = 0.f;
    for (int q = 0; q < K; ++q) se += expf(l[q] - mx);
    nll[idx] = -(mx + logf(se));
  }
}

}  // namespace hb
#ifndef HB_HOST_SHIM   // host side of the C-ABI (launch syntax): device builds only
using namespace hb;

extern "C" int humor_fit_losses(const HbFitArgs* a, int64_t* launches, cudaStream_t st) {
  if (!a || a->B <= 0 || a->T <= 0 || !a->terms || !a->loss || !a->partials) return HB_ERR_ARG;
  if (a->njx != 52 && a->njx != 73) return HB_ERR_ARG;
  if (!a->cam_joints || !a->cam_verts || !a->prior_joints || !a->roll_joints || !a->betas || !a->z) return HB_ERR_ARG;
  if (a->T > 1 && !a->contact_logits) return HB_ERR_ARG;
  fit_losses_kernel<<<a->B * a->T, LB, 0, st>>>(*a);
  HB_LAUNCH_CHECK();
  fit_reduce1_kernel<<<FIT_RB, 256, 0, st>>>(*a);
  HB_LAUNCH_CHECK();
  fit_reduce_kernel<<<1, 32, 0, st>>>(*a);
  HB_LAUNCH_CHECK();
  if (launches) *launches = 3;
  return HB_OK;
}

extern "C" int humor_gmm_nll(int B, int D, int K, const float* x, const float* logw, const float* mean, const float* Linv,
                             const float* logdet, float* nll, float* d_x, cudaStream_t st) {
  if (B <= 0 || D <= 0 || D > GMM_MAXD || K <= 0 || K > GMM_MAXK || !x || !logw || !mean || !Linv || !logdet || !nll || !d_x)
    return HB_ERR_ARG;
  static size_t granted = 0;
  const size_t need = gmm_smem_bytes(D, K);     // 134 KB at D = 138, K = 12: Linv_k + K * 8 rows of y + 8 rows of x
  if (need > 227 * 1024) return HB_ERR_ARG;
  if (need > granted) {
    HB_CUDA(cudaFuncSetAttribute(gmm_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    granted = need;
  }
  gmm_nll_kernel<<<cdiv(B, GMM_RT), 256, gmm_smem_bytes(D, K), st>>>(B, D, K, x, logw, mean, Linv, logdet, nll, d_x);
  HB_LAUNCH_CHECK();
  return HB_OK;
}
extern "C" size_t humor_gmm_workspace_bytes(int B, int D, int K) {
  return ((size_t)2 * K * B * D + (size_t)B * K) * sizeof(float);
}
extern "C" int humor_gmm_nll_ws(int B, int D, int K, const float* x, const float* logw, const float* mean, const float* Linv,
                                const float* logdet, float* nll, float* d_x, float* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (B <= 0 || D <= 0 || D > GMM_MAXD || K <= 0 || K > GMM_MAXK || !x || !logw || !mean || !Linv || !logdet || !nll || !d_x || !workspace)
    return HB_ERR_ARG;
  if (workspace_bytes < humor_gmm_workspace_bytes(B, D, K)) return HB_ERR_WORKSPACE;
  static size_t granted = 0;
  const size_t need = gmb_smem_bytes(D);
  if (need > 227 * 1024) return HB_ERR_ARG;
  if (need > granted) {
    HB_CUDA(cudaFuncSetAttribute(gmm_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    HB_CUDA(cudaFuncSetAttribute(gmm_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    granted = need;
  }
  float* y = workspace;
  float* dxk = y + (size_t)K * B * D;
  float* lp = dxk + (size_t)K * B * D;
  const dim3 grid(K, cdiv(B, GMB_ROWS));
  gmm_pass1_kernel<<<grid, 256, need, st>>>(B, D, K, x, logw, mean, Linv, logdet, y, lp);
  HB_LAUNCH_CHECK();
  gmm_pass2_kernel<<<grid, 256, need, st>>>(B, D, K, Linv, y, lp, dxk);
  HB_LAUNCH_CHECK();
  gmm_reduce_kernel<<<cdiv(B * D, 256), 256, 0, st>>>(B, D, K, lp, dxk, nll, d_x);
  HB_LAUNCH_CHECK();
  return HB_OK;
}
#endif  // HB_HOST_SHIM
