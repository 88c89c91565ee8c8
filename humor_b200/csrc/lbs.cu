// SMPL+H linear blend skinning, forward and reverse (exact-fp32 FFMA path).
// Replaces BodyModel.forward -> smplx.SMPLH.forward -> smplx.lbs.lbs
// (humor/body_model/body_model.py:72-115; algorithm SURVEY.md Appendix A.1).
//
// Decomposition
//   lbs_pose_kernel      per frame : Rodrigues x22, pose feature, rest joints from betas, kinematic chain,
//                                    skinning transforms A[52][3x4], posed joints (+trans)
//   lbs_skin_fwd_kernel  (vertex tile x frame tile): v_posed = v_template + [betas|pose_feat] . blend
//                                    (one K=208 contraction: shape blend and pose blend fused), then
//                                    sparse (ELL) skinning  v = sum_w w * (A_j [v_posed;1]) + trans
//   lbs_skin_bwd_kernel  frame tile, loops vertex chunks: recompute v_posed, d v_posed, reduce
//                                    dA[52][3x4], d feature[208], d trans without atomics (deterministic)
//   lbs_pose_bwd_kernel  per frame : chain / Rodrigues reverse, d betas
// (N,6890,4,4) skinning transforms, pose offsets and homogeneous copies that smplx materialises
// (~1.2 MB/frame) never exist; per frame the only HBM traffic is the inputs, A (2.5 KB) and v.
#include <cstdlib>
#include "common.cuh"
#include "lbs_chain.cuh"
#include "umma_launch.cuh"
#include "../../include/humor_b200.h"

namespace hb {

constexpr int SK_VT = 64;    // vertices per block tile
constexpr int SK_FT = 64;    // frames per block tile (forward)
constexpr int SK_FPT = 16;   // frames per thread (forward)
constexpr int BW_FT = 16;    // frames per block (backward)
constexpr int LBS_SEL_LD = 192;   // floats per row of HbLbsModel.sel_blend: 64 vertex slots x 3

constexpr int TC_SLAB = 512;     // frames per tensor-core slab: v_posed slab (512 x 20736 fp32 = 42 MB) stays in L2
constexpr int TC_KF = 224;
      // feature K padded to a multiple of 32 for the TMA/UMMA tiles

struct LbsWs {
  float *feat, *A, *dfeat, *dA, *dtr, *feat_hi, *feat_lo, *vposed, *feat16;
  size_t total;
};
static LbsWs lbs_carve(float* base, int N) {
  LbsWs w;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += align_up(n, 64); return p; };
  size_t Np = align_up((size_t)N, 64);
  w.feat = take(Np * LBS_KF);
  w.A = take(Np * 624);
  w.dfeat = take(Np * LBS_KF);
  w.dA = take(Np * 624);
  w.dtr = take(Np * 4);
  w.feat_hi = take(Np * TC_KF);
  w.feat_lo = take(Np * TC_KF);
  w.vposed = take((size_t)TC_SLAB * 20736);
  w.feat16 = take(Np * 256);                    // fp16 planes of the feature rows: [Np][192] halves (blend form 4) or 2 x [Np][256] (form 5)
  w.total = off;
  return w;
}

// hi plane of an operand of the tensor-core blend: the value ROUNDED (to nearest, ties to even) to tf32 precision; lo = x - hi is
// exact in fp32.  hi + lo == x as before, so the 3xTF32 product is unchanged in accuracy, and hi alone is the best single-pass
// operand (blend form 3 uses one pass on the pose columns).  Inputs are bounded (betas, R - I): no overflow handling needed.
__device__ __forceinline__ float tf32_rn(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float((u + 0x0fffu + ((u >> 13) & 1u)) & 0xffffe000u);
}

__device__ __forceinline__ void rest_joints(const HbLbsModel& m, const float* __restrict__ beta, float* J) {
  for (int e = 0; e < LBS_J * 3; ++e) {
    float a = m.j_template[e];
    const float* d = m.j_dirs + e * LBS_NB;
#pragma unroll
    for (int l = 0; l < LBS_NB; ++l) a = fmaf(d[l], beta[l], a);
    J[e] = a;
  }
}

__global__ void lbs_pose_kernel(HbLbsModel m, int N, int fpb, const float* __restrict__ root_orient,
                                const float* __restrict__ pose_body, const float* __restrict__ betas,
                                const float* __restrict__ trans, float* feat, float* A, float* joints, int njo,
                                float* feat_hi, float* feat_lo, float a_scale, int a_fold) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* beta = betas + (size_t)(n / fpb) * LBS_NB;
  float pose[66], J[LBS_J * 3], Jp[LBS_J * 3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pose[i] = root_orient[(size_t)n * 3 + i];
  for (int i = 0; i < 63; ++i) pose[3 + i] = pose_body[(size_t)n * 63 + i];
  rest_joints(m, beta, J);
  float* f = feat ? feat + (size_t)n * LBS_KF : nullptr;
  if (f) {
#pragma unroll
    for (int l = 0; l < LBS_NB; ++l) f[l] = beta[l];
    f[205] = f[206] = f[207] = 0.f;
  }
  lbs_chain_fwd(pose, J, m.parents, f ? f + LBS_NB : nullptr, A ? A + (size_t)n * 624 : nullptr, Jp);
  if (A && (a_scale != 1.f || a_fold)) {   // transforms of the fused dense pass (see lbs_pose_warp_kernel)
    float* an = A + (size_t)n * 624;
    for (int j = 0; j < LBS_J; ++j)
      for (int i = 0; i < 3; ++i) {
        an[j * 12 + i * 4] *= a_scale; an[j * 12 + i * 4 + 1] *= a_scale; an[j * 12 + i * 4 + 2] *= a_scale;
        if (a_fold) an[j * 12 + i * 4 + 3] += trans[(size_t)n * 3 + i];
      }
  }
  if (feat_hi && f) {                 // hi/lo operand planes (x = hi + lo) of the feature row for the tensor-core blend
    for (int k = 0; k < TC_KF; ++k) {
      const float v = k < 205 ? f[k] : (k == 205 ? 1.f : 0.f);   // column 205 = 1: picks up the template row of the fused blend matrix
      const float h = tf32_rn(v);
      feat_hi[(size_t)n * TC_KF + k] = h;
      feat_lo[(size_t)n * TC_KF + k] = v - h;
    }
  }
  if (joints) {
    float t0 = trans[(size_t)n * 3], t1 = trans[(size_t)n * 3 + 1], t2 = trans[(size_t)n * 3 + 2];
    float* jo = joints + (size_t)n * njo * 3;
    for (int j = 0; j < LBS_J; ++j) { jo[3 * j] = Jp[3 * j] + t0; jo[3 * j + 1] = Jp[3 * j + 1] + t1; jo[3 * j + 2] = Jp[3 * j + 2] + t2; }
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-per-frame versions of the per-frame kernels (the thread-per-frame ones above keep ~3 KB of local arrays per
// thread and cost ~150 us per call whatever the frame count: pure latency).  One warp owns one frame; all per-frame
// state lives in shared memory; the kinematic chain runs level by level (joints of equal tree depth in parallel,
// m.depth / m.max_depth from pack_smplh); the reverse pass lets every parent PULL from its children in CSR order, so
// no atomics are needed and results are bit-reproducible.
// ------------------------------------------------------------------------------------------------
constexpr int PW = 4;            // warps (frames) per block
struct PoseSm {
  float pose[68], beta[16], J[156], Rl[198], rot[198], tw[156];
};
struct PoseBwdSm {
  float drot[198], dt[156], dJ[156], sc[468];
};

__device__ __forceinline__ void pose_forward_warp(const HbLbsModel& m, PoseSm& s, int lane, int n, int fpb,
                                                  const float* __restrict__ root_orient, const float* __restrict__ pose_body,
                                                  const float* __restrict__ betas) {
  for (int i = lane; i < 66; i += 32) s.pose[i] = i < 3 ? root_orient[(size_t)n * 3 + i] : pose_body[(size_t)n * 63 + i - 3];
  if (lane < 16) s.beta[lane] = betas[(size_t)(n / fpb) * LBS_NB + lane];
  __syncwarp();
  for (int e = lane; e < LBS_J * 3; e += 32) {
    float a = m.j_template[e];
    const float4* d = reinterpret_cast<const float4*>(m.j_dirs + e * LBS_NB);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = __ldg(d + q);
      a = fmaf(w.x, s.beta[4 * q], a); a = fmaf(w.y, s.beta[4 * q + 1], a);
      a = fmaf(w.z, s.beta[4 * q + 2], a); a = fmaf(w.w, s.beta[4 * q + 3], a);
    }
    s.J[e] = a;
  }
  if (lane < LBS_JB) rodrigues_fwd(s.pose + 3 * lane, s.Rl + 9 * lane);
  __syncwarp();
  for (int L = 0; L <= m.max_depth; ++L) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      if (j < LBS_J && m.depth[j] == L) {
        if (j == 0) {
#pragma unroll
          for (int e = 0; e < 9; ++e) s.rot[e] = s.Rl[e];
          s.tw[0] = s.J[0]; s.tw[1] = s.J[1]; s.tw[2] = s.J[2];
        } else {
          const int p = m.parents[j];
          if (j < LBS_JB) mat3_mul(s.rot + 9 * p, s.Rl + 9 * j, s.rot + 9 * j);
          const float rel[3] = {s.J[3 * j] - s.J[3 * p], s.J[3 * j + 1] - s.J[3 * p + 1], s.J[3 * j + 2] - s.J[3 * p + 2]};
          float o[3];
          mat3_vec(s.rot + 9 * lbs_rot_owner(p), rel, o);
          s.tw[3 * j] = o[0] + s.tw[3 * p]; s.tw[3 * j + 1] = o[1] + s.tw[3 * p + 1]; s.tw[3 * j + 2] = o[2] + s.tw[3 * p + 2];
        }
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(PW * 32)
lbs_pose_warp_kernel(HbLbsModel m, int N, int fpb, const float* __restrict__ root_orient, const float* __restrict__ pose_body,
                     const float* __restrict__ betas, const float* __restrict__ trans, float* feat, float* A, float* joints,
                     int njo, float* feat_hi, float* feat_lo, float a_scale, int a_fold) {
  // a_scale / a_fold (fused dense pass only, lbs_fuseg.cuh; 1 / 0 everywhere else): A is written with its rotation part times
  // a_scale (the 2^-10 that brings the fp16 planes' accumulators back to metres) and trans added to its translation column (exact
  // for weights that sum to 1, HB_LBS_WEIGHTS_SUM_1) - both would otherwise cost the skinning epilogue 48 instructions per group
  __shared__ PoseSm sm[PW];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * PW + wid;
  if (n >= N) return;
  PoseSm& s = sm[wid];
  pose_forward_warp(m, s, lane, n, fpb, root_orient, pose_body, betas);
  if (feat) {
    float* f = feat + (size_t)n * LBS_KF;
    for (int k = lane; k < TC_KF; k += 32) {
      float v = 0.f;
      if (k < LBS_NB) v = s.beta[k];
      else if (k < 205) {
        const int idx = k - LBS_NB, e = idx % 9;
        v = s.Rl[9 + idx] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
      }
      if (k < LBS_KF) f[k] = v;
      if (feat_hi) {
        const float vp = k == 205 ? 1.f : v;               // column 205 = 1: picks up the template row of the fused blend matrix
        const float h = tf32_rn(vp);
        feat_hi[(size_t)n * TC_KF + k] = h;
        feat_lo[(size_t)n * TC_KF + k] = vp - h;
      }
    }
  }
  const bool need_t = joints || a_fold;
  const float t0 = need_t ? trans[(size_t)n * 3] : 0.f, t1 = need_t ? trans[(size_t)n * 3 + 1] : 0.f, t2 = need_t ? trans[(size_t)n * 3 + 2] : 0.f;
  const float f0 = a_fold ? t0 : 0.f, f1 = a_fold ? t1 : 0.f, f2 = a_fold ? t2 : 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    if (j < LBS_J) {
      if (A) {
        const float* r = s.rot + 9 * lbs_rot_owner(j);
        float c[3];
        mat3_vec(r, s.J + 3 * j, c);
        float4* a = reinterpret_cast<float4*>(A + (size_t)n * 624 + j * 12);
        a[0] = make_float4(r[0] * a_scale, r[1] * a_scale, r[2] * a_scale, s.tw[3 * j] - c[0] + f0);
        a[1] = make_float4(r[3] * a_scale, r[4] * a_scale, r[5] * a_scale, s.tw[3 * j + 1] - c[1] + f1);
        a[2] = make_float4(r[6] * a_scale, r[7] * a_scale, r[8] * a_scale, s.tw[3 * j + 2] - c[2] + f2);
      }
      if (joints) {
        float* jo = joints + ((size_t)n * njo + j) * 3;
        jo[0] = s.tw[3 * j] + t0; jo[1] = s.tw[3 * j + 1] + t1; jo[2] = s.tw[3 * j + 2] + t2;
      }
    }
  }
}

__global__ void __launch_bounds__(PW * 32)
lbs_pose_bwd_warp_kernel(HbLbsModel m, int N, int fpb, const float* __restrict__ root_orient, const float* __restrict__ pose_body,
                         const float* __restrict__ betas, const float* __restrict__ dfeat, const float* __restrict__ dA,
                         const float* __restrict__ dtr, const float* __restrict__ djoints, int njo, float* d_root_orient,
                         float* d_pose_body, float* d_betas, float* d_trans) {
  __shared__ PoseSm sm[PW];
  __shared__ PoseBwdSm sb[PW];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * PW + wid;
  if (n >= N) return;
  PoseSm& s = sm[wid];
  PoseBwdSm& b = sb[wid];
  pose_forward_warp(m, s, lane, n, fpb, root_orient, pose_body, betas);
  const float* dAn = dA ? dA + (size_t)n * 624 : nullptr;
  const float* dj = djoints ? djoints + (size_t)n * njo * 3 : nullptr;
  // ---- a. per-joint seeds: A_j = [rot_o | tw_j - rot_o J_j], Jp_j = tw_j
  for (int i = lane; i < 198; i += 32) b.drot[i] = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    if (j < LBS_J) {
      float gt[3] = {0.f, 0.f, 0.f};
      if (dAn) {
        gt[0] = dAn[j * 12 + 3]; gt[1] = dAn[j * 12 + 7]; gt[2] = dAn[j * 12 + 11];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int k = 0; k < 3; ++k) b.sc[j * 9 + i * 3 + k] = dAn[j * 12 + i * 4 + k] - gt[i] * s.J[3 * j + k];
        float back[3];
        mat3_tvec(s.rot + 9 * lbs_rot_owner(j), gt, back);
        b.dJ[3 * j] = -back[0]; b.dJ[3 * j + 1] = -back[1]; b.dJ[3 * j + 2] = -back[2];
      } else {
#pragma unroll
        for (int e = 0; e < 9; ++e) b.sc[j * 9 + e] = 0.f;
        b.dJ[3 * j] = b.dJ[3 * j + 1] = b.dJ[3 * j + 2] = 0.f;
      }
      if (dj) { gt[0] += dj[3 * j]; gt[1] += dj[3 * j + 1]; gt[2] += dj[3 * j + 2]; }
      b.dt[3 * j] = gt[0]; b.dt[3 * j + 1] = gt[1]; b.dt[3 * j + 2] = gt[2];
    }
  }
  __syncwarp();
  if (lane < LBS_JB) {                                   // rotation owners collect the A-part (wrists also their 15 hand joints)
    float acc[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) acc[e] = b.sc[lane * 9 + e];
    if (lane == 20 || lane == 21) {
      const int h0 = lane == 20 ? 22 : 37;
      for (int hj = h0; hj < h0 + 15; ++hj)
#pragma unroll
        for (int e = 0; e < 9; ++e) acc[e] += b.sc[hj * 9 + e];
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) b.drot[lane * 9 + e] = acc[e];
  }
  __syncwarp();
  for (int i = lane; i < 468; i += 32) b.sc[i] = 0.f;    // sc now carries g (x) rel of hand joints whose parent is a hand joint
  __syncwarp();
  // ---- b. reverse levels: every joint pulls from its children (fixed CSR order)
  for (int L = m.max_depth; L >= 0; --L) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      if (j < LBS_J && m.depth[j] == L) {
        const int o = lbs_rot_owner(j);
        float adt[3] = {b.dt[3 * j], b.dt[3 * j + 1], b.dt[3 * j + 2]};
        float aJ[3] = {b.dJ[3 * j], b.dJ[3 * j + 1], b.dJ[3 * j + 2]};
        float ar[9];
        if (j < LBS_JB) {
#pragma unroll
          for (int e = 0; e < 9; ++e) ar[e] = b.drot[j * 9 + e];
        }
        for (int ci = m.child_start[j]; ci < m.child_start[j + 1]; ++ci) {
          const int c = m.child_list[ci];
          const float g[3] = {b.dt[3 * c], b.dt[3 * c + 1], b.dt[3 * c + 2]};
          const float rel[3] = {s.J[3 * c] - s.J[3 * j], s.J[3 * c + 1] - s.J[3 * j + 1], s.J[3 * c + 2] - s.J[3 * j + 2]};
          float back[3];
          mat3_tvec(s.rot + 9 * o, g, back);
#pragma unroll
          for (int i = 0; i < 3; ++i) { adt[i] += g[i]; aJ[i] -= back[i]; }
          if (j < LBS_JB) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int k = 0; k < 3; ++k) ar[i * 3 + k] += g[i] * rel[k];
            if (c < LBS_JB) {                             // rot_c = rot_j R_c
              const float* dc = b.drot + 9 * c;
              const float* Rc = s.Rl + 9 * c;
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k) ar[i * 3 + k] += dc[i * 3] * Rc[k * 3] + dc[i * 3 + 1] * Rc[k * 3 + 1] + dc[i * 3 + 2] * Rc[k * 3 + 2];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int k = 0; k < 3; ++k) b.sc[c * 9 + i * 3 + k] = g[i] * rel[k];
          }
        }
        if (j == 20 || j == 21) {                         // wrist: + contributions parked by its hand joints
          const int h0 = j == 20 ? 22 : 37;
          for (int hj = h0; hj < h0 + 15; ++hj)
#pragma unroll
            for (int e = 0; e < 9; ++e) ar[e] += b.sc[hj * 9 + e];
        }
        if (j > 0) {
          float back[3];
          mat3_tvec(s.rot + 9 * lbs_rot_owner(m.parents[j]), adt, back);
          aJ[0] += back[0]; aJ[1] += back[1]; aJ[2] += back[2];
        } else {
          aJ[0] += adt[0]; aJ[1] += adt[1]; aJ[2] += adt[2];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) { b.dt[3 * j + i] = adt[i]; b.dJ[3 * j + i] = aJ[i]; }
        if (j < LBS_JB) {
#pragma unroll
          for (int e = 0; e < 9; ++e) b.drot[j * 9 + e] = ar[e];
        }
      }
    }
    __syncwarp();
  }
  // ---- c. rotations -> axis-angle
  if (lane < LBS_JB) {
    float dR[9];
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) dR[e] = b.drot[e];
    } else {
      const float* rp = s.rot + 9 * m.parents[lane];
      const float* dr = b.drot + 9 * lane;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) dR[i * 3 + k] = rp[i] * dr[k] + rp[3 + i] * dr[3 + k] + rp[6 + i] * dr[6 + k];
      if (dfeat) {
#pragma unroll
        for (int e = 0; e < 9; ++e) dR[e] += dfeat[(size_t)n * LBS_KF + LBS_NB + (lane - 1) * 9 + e];
      }
    }
    float d3[3] = {0.f, 0.f, 0.f};
    rodrigues_bwd(s.pose + 3 * lane, dR, d3);
    float* out = lane == 0 ? d_root_orient + (size_t)n * 3 : d_pose_body + (size_t)n * 63 + (lane - 1) * 3;
    out[0] = d3[0]; out[1] = d3[1]; out[2] = d3[2];
  }
  // ---- d. betas:  dfeat[0:16] + J_dirs^T dJ   (two half-sums combined in a fixed order)
  {
    const int l = lane & 15, half = lane >> 4;
    float a = 0.f;
    for (int e = half * 78; e < half * 78 + 78; ++e) a = fmaf(m.j_dirs[e * LBS_NB + l], b.dJ[e], a);
    const float other = __shfl_xor_sync(0xffffffffu, a, 16);
    if (half == 0) d_betas[(size_t)n * LBS_NB + l] = (dfeat ? dfeat[(size_t)n * LBS_KF + l] : 0.f) + (a + other);
  }
  // ---- e. translation
  if (lane < 3) {
    float t = dtr ? dtr[(size_t)n * 4 + lane] : 0.f;
    if (dj)
      for (int j = 0; j < LBS_J; ++j) t += dj[3 * j + lane];
    d_trans[(size_t)n * 3 + lane] = t;
  }
}

// out[(n*out_fs) + slot*3 + c], slot = index within the vertex list (or the vertex id when dense)
__global__ void __launch_bounds__(256)
lbs_skin_fwd_kernel(HbLbsModel m, int N, const float* __restrict__ feat, const float* __restrict__ A,
                    const float* __restrict__ trans, const int* __restrict__ vlist, int nv, float* out, size_t out_fs) {
  HB_DYN_SMEM_F32(Fs);                                     // [SK_FT][LBS_KF]
  const int tid = threadIdx.x;
  const int vi = tid % SK_VT, fg = tid / SK_VT;            // 4 frame groups x 16 frames
  const int f0 = blockIdx.y * SK_FT;
  const int slot = blockIdx.x * SK_VT + vi;
  const bool vok = slot < nv;
  const int vid = vok ? (vlist ? vlist[slot] : slot) : 0;
  for (int i = tid; i < SK_FT * LBS_KF / 4; i += 256) {
    int f = i / (LBS_KF / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f0 + f < N) v = reinterpret_cast<const float4*>(feat + (size_t)(f0 + f) * LBS_KF)[i % (LBS_KF / 4)];
    reinterpret_cast<float4*>(Fs)[i] = v;
  }
  __syncthreads();
  float acc[SK_FPT][3];
#pragma unroll
  for (int f = 0; f < SK_FPT; ++f) acc[f][0] = acc[f][1] = acc[f][2] = 0.f;
  // the model's selected set (key vertices + the 21 vertex-picked joints): their blend columns are packed [208][192], so a warp
  // reads 384 contiguous bytes per k instead of 32 lines 12 bytes each out of the [208][3V] matrix
  const bool tab = m.sel_blend && ((vlist == m.sel_ids && nv == m.sel_nv) || vlist == m.extra_ids);
  const float* bp = tab ? m.sel_blend + (size_t)(slot + (vlist == m.extra_ids ? m.sel_nv : 0)) * 3 : m.blend + (size_t)vid * 3;
  const size_t bld = tab ? (size_t)LBS_SEL_LD : (size_t)m.v3_ld;
  const float* fs = Fs + (size_t)(fg * SK_FPT) * LBS_KF;
  for (int k = 0; k < 205; k += 4) {            // rows 205..207 are padding (zero)
    float p[4][3];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* b = bp + (size_t)(k + kk) * bld;
      p[kk][0] = __ldg(b); p[kk][1] = __ldg(b + 1); p[kk][2] = __ldg(b + 2);
    }
#pragma unroll
    for (int f = 0; f < SK_FPT; ++f) {
      float4 fv = *reinterpret_cast<const float4*>(fs + f * LBS_KF + k);
      const float q[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        acc[f][0] = fmaf(q[kk], p[kk][0], acc[f][0]);
        acc[f][1] = fmaf(q[kk], p[kk][1], acc[f][1]);
        acc[f][2] = fmaf(q[kk], p[kk][2], acc[f][2]);
      }
    }
  }
  if (!vok) return;
  const float vt0 = m.v_template[vid * 3], vt1 = m.v_template[vid * 3 + 1], vt2 = m.v_template[vid * 3 + 2];
  const int* wi = m.w_idx + (size_t)vid * m.wk;
  const float* wv = m.w_val + (size_t)vid * m.wk;
#pragma unroll 1
  for (int f = 0; f < SK_FPT; ++f) {
    const int n = f0 + fg * SK_FPT + f;
    if (n >= N) break;
    const float px = vt0 + acc[f][0], py = vt1 + acc[f][1], pz = vt2 + acc[f][2];
    float ox = 0.f, oy = 0.f, oz = 0.f;
    const float* An = A + (size_t)n * 624;
    for (int w = 0; w < m.wk; ++w) {
      const float wt = wv[w];
      if (wt == 0.f) continue;
      const float4* a = reinterpret_cast<const float4*>(An + wi[w] * 12);
      const float4 r0 = __ldg(a), r1 = __ldg(a + 1), r2 = __ldg(a + 2);
      ox = fmaf(wt, fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))), ox);
      oy = fmaf(wt, fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))), oy);
      oz = fmaf(wt, fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))), oz);
    }
    float* o = out + (size_t)n * out_fs + (size_t)slot * 3;
    o[0] = ox + trans[(size_t)n * 3];
    o[1] = oy + trans[(size_t)n * 3 + 1];
    o[2] = oz + trans[(size_t)n * 3 + 2];
  }
}

// Tensor-core path, second half: v_posed (template + blend, from the UMMA GEMM, L2-resident slab) -> skinned vertices.
// Block = 128 consecutive vertices x SA_F frames; the frames' 52 skinning transforms are staged in shared memory
// (the gather of <=4 transforms per vertex is the shared-memory-bound part of LBS); reads and writes are coalesced.
constexpr int SA_F = 16;
__global__ void __launch_bounds__(128)
lbs_skin_apply_kernel(HbLbsModel m, int nframes, int v3_ld, const float* __restrict__ vposed, const float* __restrict__ A,
                      const float* __restrict__ trans, float* out) {
  __shared__ __align__(16) float As[SA_F][624];
  __shared__ float Ts[SA_F][3];
  const int tid = threadIdx.x;
  const int f0 = blockIdx.y * SA_F;
  const int nf = min(SA_F, nframes - f0);
  for (int i = tid; i < nf * 624 / 4; i += 128) {
    reinterpret_cast<float4*>(&As[0][0])[i] = reinterpret_cast<const float4*>(A + (size_t)f0 * 624)[i];
  }
  if (tid < nf * 3) Ts[tid / 3][tid % 3] = trans[(size_t)f0 * 3 + tid];
  __syncthreads();
  const int v = blockIdx.x * 128 + tid;
  if (v >= m.num_verts) return;
  int wi[8];
  float wv[8];
  const int wk = m.wk < 8 ? m.wk : 8;
#pragma unroll
  for (int w = 0; w < 8; ++w) { wi[w] = w < wk ? m.w_idx[(size_t)v * m.wk + w] * 12 : 0; wv[w] = w < wk ? m.w_val[(size_t)v * m.wk + w] : 0.f; }
  for (int f = 0; f < nf; ++f) {
    const float* p = vposed + (size_t)(f0 + f) * v3_ld + (size_t)v * 3;
    const float px = p[0], py = p[1], pz = p[2];
    float ox = 0.f, oy = 0.f, oz = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (wv[w] != 0.f) {
        const float4* a = reinterpret_cast<const float4*>(&As[f][wi[w]]);
        const float4 r0 = a[0], r1 = a[1], r2 = a[2];
        ox = fmaf(wv[w], fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))), ox);
        oy = fmaf(wv[w], fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))), oy);
        oz = fmaf(wv[w], fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))), oz);
      }
    }
    for (int w = 8; w < m.wk; ++w) {                      // generic tail for models with > 8 weights per vertex
      const float wt = m.w_val[(size_t)v * m.wk + w];
      if (wt != 0.f) {
        const float* a = &As[f][m.w_idx[(size_t)v * m.wk + w] * 12];
        ox = fmaf(wt, fmaf(a[0], px, fmaf(a[1], py, fmaf(a[2], pz, a[3]))), ox);
        oy = fmaf(wt, fmaf(a[4], px, fmaf(a[5], py, fmaf(a[6], pz, a[7]))), oy);
        oz = fmaf(wt, fmaf(a[8], px, fmaf(a[9], py, fmaf(a[10], pz, a[11]))), oz);
      }
    }
    float* o = out + ((size_t)(f0 + f) * m.num_verts + v) * 3;
    o[0] = ox + Ts[f][0]; o[1] = oy + Ts[f][1]; o[2] = oz + Ts[f][2];
  }
}
// joints[n][52 + e] = verts[n][extra_ids[e]]
__global__ void lbs_gather_extra_kernel(HbLbsModel m, int N, const float* __restrict__ verts, float* joints) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 21) return;
  const int n = i / 21, e = i % 21;
  const float* s = verts + ((size_t)n * m.num_verts + m.extra_ids[e]) * 3;
  float* d = joints + ((size_t)n * 73 + 52 + e) * 3;
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

// One block owns BW_FT frames and walks over the vertex list in chunks of 64; all reductions over
// vertices are thread-private (no atomics).  dv[(n*dv_fs) + slot*3 + c].
__global__ void __launch_bounds__(256)
lbs_skin_bwd_kernel(HbLbsModel m, int N, const float* __restrict__ feat, const float* __restrict__ A,
                    const int* __restrict__ vlist, int nv, const float* __restrict__ dv, size_t dv_fs,
                    float* dfeat, float* dA, float* dtr, int accumulate,
                    const int* __restrict__ vlist2, int nv1, const float* __restrict__ dv2, size_t dv2_fs) {
  // slots [0, nv1) come from (vlist, dv); slots [nv1, nv) from the second source (vlist2, dv2) - used to fold the 21
  // vertex-picked joints into the same pass as the listed key vertices (nv1 == nv: single source)
  HB_DYN_SMEM_F32(sm);
  float* Fs = sm;                              // [BW_FT][208]
  float* GP = Fs + BW_FT * LBS_KF;             // [BW_FT][192]   d v_posed
  float* Gs = GP + BW_FT * 192;                // [BW_FT][192]   d v
  float* Ps = Gs + BW_FT * 192;                // [BW_FT][192]   v_posed
  float* dAs = Ps + BW_FT * 192;               // [BW_FT][624]
  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * BW_FT;
  for (int i = tid; i < BW_FT * LBS_KF; i += 256) {
    int f = i / LBS_KF;
    Fs[i] = (f0 + f < N) ? feat[(size_t)(f0 + f) * LBS_KF + (i % LBS_KF)] : 0.f;
  }
  for (int i = tid; i < BW_FT * 624; i += 256) dAs[i] = 0.f;
  float accF[BW_FT];
#pragma unroll
  for (int f = 0; f < BW_FT; ++f) accF[f] = 0.f;
  float accT = 0.f;
  __syncthreads();
  const int vi = tid % 64, fq = tid / 64;      // phase 1: 4 frames per thread
  for (int c0 = 0; c0 < nv; c0 += 64) {
    const int slot = c0 + vi;
    const bool vok = slot < nv;
    const int vid = vok ? (slot < nv1 ? (vlist ? vlist[slot] : slot) : vlist2[slot - nv1]) : 0;
    // ---- phase 1: v_posed and d v_posed for (4 frames, 1 vertex)
    {
      float acc[4][3];
#pragma unroll
      for (int f = 0; f < 4; ++f) acc[f][0] = acc[f][1] = acc[f][2] = 0.f;
      const bool tab = m.sel_blend && ((vlist == m.sel_ids && nv1 == m.sel_nv && (nv1 == nv || vlist2 == m.extra_ids)) ||
                                       (vlist == m.extra_ids && nv1 == nv));
      const float* bp = tab ? m.sel_blend + (size_t)(slot + (vlist == m.extra_ids ? m.sel_nv : 0)) * 3 : m.blend + (size_t)vid * 3;
      const size_t bld = tab ? (size_t)LBS_SEL_LD : (size_t)m.v3_ld;
#pragma unroll 5
      for (int k = 0; k < 205; ++k) {
        const float* b = bp + (size_t)k * bld;
        const float p0 = __ldg(b), p1 = __ldg(b + 1), p2 = __ldg(b + 2);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          const float q = Fs[(fq * 4 + f) * LBS_KF + k];
          acc[f][0] = fmaf(q, p0, acc[f][0]);
          acc[f][1] = fmaf(q, p1, acc[f][1]);
          acc[f][2] = fmaf(q, p2, acc[f][2]);
        }
      }
      const float vt0 = m.v_template[vid * 3], vt1 = m.v_template[vid * 3 + 1], vt2 = m.v_template[vid * 3 + 2];
      const int* wi = m.w_idx + (size_t)vid * m.wk;
      const float* wv = m.w_val + (size_t)vid * m.wk;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int fl = fq * 4 + f, n = f0 + fl;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f;
        if (vok && n < N) {
          const float* g = slot < nv1 ? dv + (size_t)n * dv_fs + (size_t)slot * 3 : dv2 + (size_t)n * dv2_fs + (size_t)(slot - nv1) * 3;
          g0 = g[0]; g1 = g[1]; g2 = g[2];
          const float* An = A + (size_t)n * 624;
          for (int w = 0; w < m.wk; ++w) {
            const float wt = wv[w];
            if (wt == 0.f) continue;
            const float* a = An + wi[w] * 12;
            q0 = fmaf(wt, a[0] * g0 + a[4] * g1 + a[8] * g2, q0);
            q1 = fmaf(wt, a[1] * g0 + a[5] * g1 + a[9] * g2, q1);
            q2 = fmaf(wt, a[2] * g0 + a[6] * g1 + a[10] * g2, q2);
          }
        }
        float* gp = GP + fl * 192 + vi * 3; gp[0] = q0; gp[1] = q1; gp[2] = q2;
        float* gs = Gs + fl * 192 + vi * 3; gs[0] = g0; gs[1] = g1; gs[2] = g2;
        float* ps = Ps + fl * 192 + vi * 3; ps[0] = vt0 + acc[f][0]; ps[1] = vt1 + acc[f][1]; ps[2] = vt2 + acc[f][2];
      }
    }
    __syncthreads();
    const int cn = min(64, nv - c0);
    // ---- phase 2a: dA[f][j][r][c] += w * g[r] * [p;1][c]     thread = (frame, element of the 3x4)
    {
      const int f = tid / 16, e = tid % 16;
      if (e < 12) {
        const int r = e / 4, c = e % 4;
        float* dst = dAs + f * 624 + e;
        for (int i = 0; i < cn; ++i) {
          const int v2 = (c0 + i) < nv1 ? (vlist ? vlist[c0 + i] : c0 + i) : vlist2[c0 + i - nv1];
          const float gr = Gs[f * 192 + i * 3 + r];
          const float pc = c < 3 ? Ps[f * 192 + i * 3 + c] : 1.f;
          const float gpc = gr * pc;
          const int* wi = m.w_idx + (size_t)v2 * m.wk;
          const float* wv = m.w_val + (size_t)v2 * m.wk;
          for (int w = 0; w < m.wk; ++w) {
            const float wt = __ldg(wv + w);
            if (wt != 0.f) dst[__ldg(wi + w) * 12] += wt * gpc;
          }
        }
      }
    }
    // ---- phase 2b: d feat[f][k] += sum_vc GP[f][vc] * blend_t[vc][k]   thread = k
    if (tid < LBS_KF) {
      const bool tab2 = m.sel_blend && vlist == m.sel_ids && nv1 == m.sel_nv && (nv1 == nv || vlist2 == m.extra_ids) && c0 == 0;
      if (tab2) {
        // packed columns: feature row tid of sel_blend holds the 3 * cn values of this chunk contiguously, in GP's own order;
        // 16 of them are fetched before the first is used (one dependent L2 round trip per (vertex, coordinate) otherwise)
        const float4* rowp = reinterpret_cast<const float4*>(m.sel_blend + (size_t)tid * LBS_SEL_LD);
        const int ne = cn * 3;
        for (int e0 = 0; e0 < ne; e0 += 16) {
          float4 q4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) q4[u] = (e0 + 4 * u < ne) ? __ldg(rowp + (e0 >> 2) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float bv[4] = {q4[u].x, q4[u].y, q4[u].z, q4[u].w};
#pragma unroll
            for (int z = 0; z < 4; ++z) {
              const int e = e0 + 4 * u + z;
              if (e < ne) {
#pragma unroll
                for (int f = 0; f < BW_FT; ++f) accF[f] = fmaf(GP[f * 192 + e], bv[z], accF[f]);
              }
            }
          }
        }
      } else {
        for (int i = 0; i < cn; ++i) {
          const int v2 = (c0 + i) < nv1 ? (vlist ? vlist[c0 + i] : c0 + i) : vlist2[c0 + i - nv1];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float bt = __ldg(m.blend_t + ((size_t)v2 * 3 + c) * LBS_KF + tid);
#pragma unroll
            for (int f = 0; f < BW_FT; ++f) accF[f] = fmaf(GP[f * 192 + i * 3 + c], bt, accF[f]);
          }
        }
      }
    } else if (tid < LBS_KF + BW_FT * 3) {
      // ---- phase 2c: d trans[f][c] += sum_i g
      const int f = (tid - LBS_KF) / 3, c = (tid - LBS_KF) % 3;
      for (int i = 0; i < cn; ++i) accT += Gs[f * 192 + i * 3 + c];
    }
    __syncthreads();
  }
  for (int i = tid; i < BW_FT * 624; i += 256) {
    int f = i / 624;
    if (f0 + f < N) {
      float* d = dA + (size_t)(f0 + f) * 624 + (i % 624);
      *d = accumulate ? *d + dAs[i] : dAs[i];
    }
  }
  if (tid < LBS_KF) {
#pragma unroll
    for (int f = 0; f < BW_FT; ++f)
      if (f0 + f < N) {
        float* d = dfeat + (size_t)(f0 + f) * LBS_KF + tid;
        *d = accumulate ? *d + accF[f] : accF[f];
      }
  } else if (tid < LBS_KF + BW_FT * 3) {
    const int f = (tid - LBS_KF) / 3, c = (tid - LBS_KF) % 3;
    if (f0 + f < N) {
      float* d = dtr + (size_t)(f0 + f) * 4 + c;
      *d = accumulate ? *d + accT : accT;
    }
  }
}

__global__ void lbs_pose_bwd_kernel(HbLbsModel m, int N, int fpb, const float* __restrict__ root_orient,
                                    const float* __restrict__ pose_body, const float* __restrict__ betas,
                                    const float* __restrict__ dfeat, const float* __restrict__ dA,
                                    const float* __restrict__ dtr, const float* __restrict__ djoints, int njo,
                                    float* d_root_orient, float* d_pose_body, float* d_betas, float* d_trans) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* beta = betas + (size_t)(n / fpb) * LBS_NB;
  float pose[66], J[LBS_J * 3], dpose[66], dJ[LBS_J * 3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pose[i] = root_orient[(size_t)n * 3 + i];
  for (int i = 0; i < 63; ++i) pose[3 + i] = pose_body[(size_t)n * 63 + i];
  rest_joints(m, beta, J);
  const float* dj = djoints ? djoints + (size_t)n * njo * 3 : nullptr;
  lbs_chain_bwd(pose, J, m.parents, dA ? dA + (size_t)n * 624 : nullptr, dj,
                dfeat ? dfeat + (size_t)n * LBS_KF + LBS_NB : nullptr, dpose, dJ);
#pragma unroll
  for (int i = 0; i < 3; ++i) d_root_orient[(size_t)n * 3 + i] = dpose[i];
  for (int i = 0; i < 63; ++i) d_pose_body[(size_t)n * 63 + i] = dpose[3 + i];
  for (int l = 0; l < LBS_NB; ++l) {
    float a = dfeat ? dfeat[(size_t)n * LBS_KF + l] : 0.f;
    for (int e = 0; e < LBS_J * 3; ++e) a = fmaf(m.j_dirs[e * LBS_NB + l], dJ[e], a);
    d_betas[(size_t)n * LBS_NB + l] = a;
  }
  float t[3] = {dtr ? dtr[(size_t)n * 4] : 0.f, dtr ? dtr[(size_t)n * 4 + 1] : 0.f, dtr ? dtr[(size_t)n * 4 + 2] : 0.f};
  // the 21 vertex-picked joints reach trans through the skinning pass (dtr); only the chain joints add here
  if (dj)
    for (int j = 0; j < LBS_J; ++j) { t[0] += dj[3 * j]; t[1] += dj[3 * j + 1]; t[2] += dj[3 * j + 2]; }
  d_trans[(size_t)n * 3] = t[0]; d_trans[(size_t)n * 3 + 1] = t[1]; d_trans[(size_t)n * 3 + 2] = t[2];
}

// shaped template of every SEQUENCE for the fused dense pass (LbsFusegArgs.vs): out[s][c] = (v_template[c] + sum_l blend[l][c] beta[s][l])
// * scale, c < 3V, zero in the padding columns.  One shape per frames_per_beta frames: 16 of the GEMM's 206 columns leave the per-frame
// product (and the fourth k-block with them)
__global__ void lbs_shape_rows_kernel(HbLbsModel m, int nseq, const float* __restrict__ betas, float scale, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (c >= m.v3_ld || s >= nseq) return;
  float v = 0.f;
  if (c < 3 * m.num_verts) {
    v = m.v_template[c];
#pragma unroll
    for (int l = 0; l < LBS_NB; ++l) v = fmaf(m.blend[(size_t)l * m.v3_ld + c], __ldg(betas + (size_t)s * LBS_NB + l), v);
  }
  out[(size_t)s * m.v3_ld + c] = v * scale;
}

static const bool g_thread_pose = (getenv("HB_LBS_THREAD") != nullptr);
static const bool g_no_vs = (getenv("HB_LBS_NO_SHAPE_ROWS") != nullptr);   // A/B: betas stay in the GEMM (K = 256) whatever frames_per_beta is
static const bool g_no_fold = (getenv("HB_LBS_NO_FOLD") != nullptr);   // A/B: the fused pass adds trans per vertex instead of inside A
// dense skinning pass: 1 = blend GEMM (umma_gemm3_kernel, v_posed slabs through L2) + lane-per-vertex lbs_skin_apply_kernel (the
// round-1 default, kept as the form for meshes without group tables and as the A/B partner), 3 (DEFAULT since round 2: 1.06 ms vs
// 3.35 ms per 15 360 frames on the B200, profiles/r02g_*) = blend GEMM + group skinning fused in one persistent tcgen05 kernel
// (lbs_fuseg.cuh).  Forms 2 (lane-per-frame skin pass, persistent 128x256 blend) and the round-1 fused kernel measured slower and
// were removed (records: profiles/r01*, r02a*).
static int g_skin_form = getenv("HB_LBS_SKIN") ? atoi(getenv("HB_LBS_SKIN")) : 3;
static int g_used_skin = 0, g_used_blend = 0;   // forms the last dense tensor-core call actually ran (0: none yet)
// blend contraction: 1 = three TF32 passes on fp32 hi/lo planes; with skin form 3 also 5 (DEFAULT) = fp16 hi + lo operand planes,
// three products per k-block - the accuracy of three TF32 passes (vertices within 5e-6 m of form 1) from 4-byte operand elements.
// (Forms 3 and 4 - single-pass pose columns in tf32 / fp16, <= 7e-5 m - lost to form 5 on both speed and accuracy: removed.)
static int g_blend_form = getenv("HB_LBS_BLEND") ? atoi(getenv("HB_LBS_BLEND")) : 5;
static int g_slab = getenv("HB_LBS_SLAB") ? atoi(getenv("HB_LBS_SLAB")) : 0;
static const size_t SKIN_FWD_SMEM = (size_t)SK_FT * LBS_KF * sizeof(float);
static const size_t SKIN_BWD_SMEM = (size_t)(BW_FT * LBS_KF + 3 * BW_FT * 192 + BW_FT * 624) * sizeof(float);

}  // namespace hb

#ifndef HB_HOST_SHIM   // host side of the C-ABI (launch syntax): device builds only
using namespace hb;

extern "C" int humor_lbs_configure(int skin_form, int blend_form, int slab_frames) {
  if ((skin_form != 0 && skin_form != 1 && skin_form != 3) || (blend_form != 0 && blend_form != 1 && blend_form != 5) ||
      (slab_frames != 0 && (slab_frames < 128 || slab_frames > TC_SLAB)))
    return HB_ERR_ARG;
  if (skin_form) g_skin_form = skin_form;
  if (blend_form) g_blend_form = blend_form;
  if (slab_frames) g_slab = slab_frames;
  return HB_OK;
}

extern "C" int humor_lbs_set_fuseg_ctas(int n) {
  if (n < 0) return HB_ERR_ARG;
  lbs_set_fuseg_ctas(n);
  return HB_OK;
}

extern "C" int humor_lbs_forms_used(int* skin_form, int* blend_form) {
  if (skin_form) *skin_form = g_used_skin;
  if (blend_form) *blend_form = g_used_blend;
  return HB_OK;
}

// measurement: CUDA events around the fused kernel of the next dense calls, on the stream it is launched on (eager calls only: an
// event pair cannot be timed inside a stream capture)
static int g_time_fuseg = 0, g_time_count = 0;
static cudaEvent_t g_time_ev[2] = {nullptr, nullptr};
static float g_time_ms = 0.f;
static bool g_time_open = false;
static void fuseg_time_collect() {
  if (!g_time_open) return;
  float ms = 0.f;
  if (cudaEventSynchronize(g_time_ev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, g_time_ev[0], g_time_ev[1]) == cudaSuccess) {
    g_time_ms += ms; ++g_time_count;
  }
  g_time_open = false;
}
extern "C" int humor_lbs_fuseg_timing(int enable, float* ms_sum, int* launches) {
  fuseg_time_collect();
  if (ms_sum) *ms_sum = g_time_ms;
  if (launches) *launches = g_time_count;
  if (enable && !g_time_ev[0] && (cudaEventCreate(&g_time_ev[0]) != cudaSuccess || cudaEventCreate(&g_time_ev[1]) != cudaSuccess)) return (int)cudaErrorUnknown;
  if (enable != g_time_fuseg) { g_time_ms = 0.f; g_time_count = 0; }
  g_time_fuseg = enable ? 1 : 0;
  return HB_OK;
}

extern "C" size_t humor_lbs_workspace_bytes(int N) { return lbs_carve(nullptr, N).total * sizeof(float); }

extern "C" int humor_lbs_fwd(const HbLbsModel* m, int N, int fpb, const float* root_orient, const float* pose_body,
                             const float* betas, const float* trans, float* workspace, size_t workspace_bytes,
                             const int* vlist, int nv, float* verts, float* joints, int njo, int64_t* launches,
                             cudaStream_t st) {
  if (!m || N <= 0 || fpb <= 0 || !root_orient || !pose_body || !betas || !trans || !workspace) return HB_ERR_ARG;
  if (njo != 52 && njo != 73) return HB_ERR_ARG;
  LbsWs ws = lbs_carve(workspace, N);
  if (workspace_bytes < ws.total * sizeof(float)) return HB_ERR_WORKSPACE;
  int64_t nl = 0;
  const bool need_skin = (verts != nullptr) || (joints && njo == 73);
  // dense output of >= 128 frames: blend on the 5th-gen tensor cores (UMMA 3xTF32) + shared-memory skinning pass
  const bool tc = verts && !vlist && N >= 128 && m->use_umma && m->blend_t_hi && m->v3_ld <= 20736 && umma_available();
  // skin form 3: one persistent tcgen05 kernel, blend accumulators skinned straight out of TMEM by lane = frame (lbs_fuseg.cuh)
  const bool fuseg = tc && g_skin_form == 3 && (m->flags & HB_LBS_PLANES_TEMPLATE) && m->ft_tab && m->ft_rec && m->ft_rec_stride > 0 &&
                     m->num_groups > 0 && m->ft_nct == cdiv(m->num_groups, 8) && (m->num_verts % 2) == 0 && m->v3_ld % 4 == 0;
  const bool f16x3 = fuseg && g_blend_form == 5 && m->blend16a_h && m->blend16a_l;
  // one shape per >= 32 frames: template + shape blend per sequence (rows in the form-1 slab, unused on this path), pose columns only
  const int nseq = cdiv(N, fpb);
  const bool vsrows = f16x3 && fpb >= 32 && nseq <= TC_SLAB && m->blend16p_h && m->blend16p_l && !g_no_vs;
  // the fused pass is the only reader of A then: its transforms come pre-scaled, and with the translation when that is exact
  const float a_scale = f16x3 ? 0.0009765625f : 1.f;
  const int a_fold = (fuseg && (m->flags & HB_LBS_WEIGHTS_SUM_1) && !g_no_fold) ? 1 : 0;
  const bool tf32_planes = tc && !f16x3;        // the fp16 forms read their own planes (feat_f16_kernel): 27 MB of tf32 planes not written
  if (m->depth && m->child_start && !g_thread_pose)
    lbs_pose_warp_kernel<<<cdiv(N, PW), PW * 32, 0, st>>>(*m, N, fpb, root_orient, pose_body, betas, trans,
                                                         need_skin ? ws.feat : nullptr, need_skin ? ws.A : nullptr, joints, njo,
                                                         tf32_planes ? ws.feat_hi : nullptr, tf32_planes ? ws.feat_lo : nullptr, a_scale, a_fold);
  else
    lbs_pose_kernel<<<cdiv(N, 64), 64, 0, st>>>(*m, N, fpb, root_orient, pose_body, betas, trans,
                                               need_skin ? ws.feat : nullptr, need_skin ? ws.A : nullptr, joints, njo,
                                               tf32_planes ? ws.feat_hi : nullptr, tf32_planes ? ws.feat_lo : nullptr, a_scale, a_fold);
  HB_LAUNCH_CHECK(); ++nl;
  if (fuseg) {
    LbsFusegArgs fa;
    fa.N = N; fa.num_verts = m->num_verts; fa.num_groups = m->num_groups; fa.nrt = fa.nct = 0;
    fa.nkb16 = f16x3 ? 4 : 0;
    fa.ft_tab = m->ft_tab; fa.ft_rec = static_cast<const unsigned char*>(m->ft_rec); fa.ft_rec_stride = m->ft_rec_stride;
    fa.A = ws.A; fa.trans = a_fold ? nullptr : trans; fa.out = verts;
    fa.vs = nullptr; fa.vs_ld = 0; fa.fpb = fpb;
    if (vsrows) {
      // pose columns as fp16 hi + (unscaled) lo planes, K = 192; template + shape blend per sequence, added by the epilogue
      unsigned short* f16h = reinterpret_cast<unsigned short*>(ws.feat16);
      unsigned short* f16l = f16h + (size_t)align_up((size_t)N, 64) * 256;
      lbs_shape_rows_kernel<<<dim3(cdiv(m->v3_ld, 256), nseq), 256, 0, st>>>(*m, nseq, betas, 1024.f, ws.vposed);
      HB_LAUNCH_CHECK(); ++nl;
      HB_CUDA(launch_feat_f16(ws.feat, LBS_KF, 205, N, LBS_NB, 3, f16h, f16l, -1, st));
      ++nl;
      fa.nkb16 = 3; fa.vs = ws.vposed; fa.vs_ld = m->v3_ld;
      if (g_time_fuseg) { fuseg_time_collect(); HB_CUDA(cudaEventRecord(g_time_ev[0], st)); }
      HB_CUDA(launch_lbs_fuseg(nullptr, nullptr, TC_KF, nullptr, nullptr, TC_KF, m->v3_ld, 0, f16h, m->blend16p_h, f16l, m->blend16p_l, 192,
                               fa, st));
      if (g_time_fuseg) { HB_CUDA(cudaEventRecord(g_time_ev[1], st)); g_time_open = true; }
    } else if (f16x3) {
      // every column as fp16 hi + (unscaled) lo planes, K = 256: three products per k-block, no tf32 k-blocks
      unsigned short* f16h = reinterpret_cast<unsigned short*>(ws.feat16);
      unsigned short* f16l = f16h + (size_t)align_up((size_t)N, 64) * 256;
      HB_CUDA(launch_feat_f16(ws.feat, LBS_KF, LBS_KF, N, 0, 4, f16h, f16l, 205, st));
      ++nl;
      HB_CUDA(launch_lbs_fuseg(nullptr, nullptr, TC_KF, nullptr, nullptr, TC_KF, m->v3_ld, 0, f16h, m->blend16a_h, f16l, m->blend16a_l, 256,
                               fa, st));
    } else {
      HB_CUDA(launch_lbs_fuseg(ws.feat_hi, ws.feat_lo, TC_KF, m->blend_t_hi, m->blend_t_lo, TC_KF, m->v3_ld, TC_KF, nullptr, nullptr, nullptr,
                               nullptr, 0, fa, st));
    }
    g_used_skin = 3; g_used_blend = f16x3 ? 5 : 1;
    ++nl;
    if (joints && njo == 73) {
      lbs_gather_extra_kernel<<<cdiv(N * 21, 256), 256, 0, st>>>(*m, N, verts, joints);
      HB_LAUNCH_CHECK(); ++nl;
    }
  } else if (tc) {
    GemmEpi ep;
    ep.bias = (m->flags & HB_LBS_PLANES_TEMPLATE) ? nullptr : m->v_template;     // planes with the template in column 205 need no bias
    ep.gamma = ep.beta = nullptr; ep.xhat = ep.rstd = nullptr; ep.ldxh = 0; ep.Cch = 0; ep.gsize = 64;
    static const int b_const = getenv("HB_UMMA_PREFETCH_B") ? 1 : 0;      // B = the model's blend planes: constant
    ep.b_const = b_const;
    // frames per slab: the v_posed slab must stay in L2 between the two kernels (<= TC_SLAB rows of workspace)
    const int slab = (g_slab >= 128 && g_slab <= TC_SLAB) ? g_slab : TC_SLAB;
    g_used_skin = g_used_blend = 1;
    for (int f0 = 0; f0 < N; f0 += slab) {
      const int nf = (N - f0 < slab) ? N - f0 : slab;
      HB_CUDA(launch_umma_gemm3_bn(ws.feat_hi + (size_t)f0 * TC_KF, ws.feat_lo + (size_t)f0 * TC_KF, TC_KF, m->blend_t_hi, m->blend_t_lo,
                                   TC_KF, nf, 3 * m->num_verts, TC_KF, ws.vposed, nullptr, nullptr, m->v3_ld, EPI_BIAS, ep, 128, st));
      dim3 grid(cdiv(m->num_verts, 128), cdiv(nf, SA_F));
      lbs_skin_apply_kernel<<<grid, 128, 0, st>>>(*m, nf, m->v3_ld, ws.vposed, ws.A + (size_t)f0 * 624, trans + (size_t)f0 * 3,
                                                  verts + (size_t)f0 * m->num_verts * 3);
      HB_LAUNCH_CHECK();
      nl += 2;
    }
    if (joints && njo == 73) {
      lbs_gather_extra_kernel<<<cdiv(N * 21, 256), 256, 0, st>>>(*m, N, verts, joints);
      HB_LAUNCH_CHECK(); ++nl;
    }
  } else if (need_skin) {
    static bool attr_fwd = false;     // once per process (and outside any stream capture: the first call is a warm-up)
    if (!attr_fwd) {
      HB_CUDA(cudaFuncSetAttribute(lbs_skin_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKIN_FWD_SMEM));
      attr_fwd = true;
    }
    if (verts) {
      const int nvv = vlist ? nv : m->num_verts;
      if (nvv > 0) {
        dim3 grid(cdiv(nvv, SK_VT), cdiv(N, SK_FT));
        lbs_skin_fwd_kernel<<<grid, 256, SKIN_FWD_SMEM, st>>>(*m, N, ws.feat, ws.A, trans, vlist, nvv, verts, (size_t)nvv * 3);
        HB_LAUNCH_CHECK(); ++nl;
      }
    }
    if (joints && njo == 73) {
      dim3 grid(1, cdiv(N, SK_FT));
      lbs_skin_fwd_kernel<<<grid, 256, SKIN_FWD_SMEM, st>>>(*m, N, ws.feat, ws.A, trans, m->extra_ids, 21, joints + 52 * 3, (size_t)73 * 3);
      HB_LAUNCH_CHECK(); ++nl;
    }
  }
  if (launches) *launches = nl;
  return HB_OK;
}

extern "C" int humor_lbs_bwd(const HbLbsModel* m, int N, int fpb, const float* root_orient, const float* pose_body,
                             const float* betas, const float* trans, float* workspace, size_t workspace_bytes,
                             const int* vlist, int nv, const float* d_verts, const float* d_joints, int njo,
                             float* d_root_orient, float* d_pose_body, float* d_betas, float* d_trans,
                             int64_t* launches, cudaStream_t st) {
  if (!m || N <= 0 || fpb <= 0 || !root_orient || !pose_body || !betas || !workspace) return HB_ERR_ARG;
  if (!d_root_orient || !d_pose_body || !d_betas || !d_trans) return HB_ERR_ARG;
  if (njo != 52 && njo != 73) return HB_ERR_ARG;
  LbsWs ws = lbs_carve(workspace, N);
  if (workspace_bytes < ws.total * sizeof(float)) return HB_ERR_WORKSPACE;
  int64_t nl = 0;
  const bool xj = d_joints && njo == 73;
  const bool need_skin = (d_verts != nullptr) || xj;
  if (need_skin) {
    // recompute the per-frame forward (feature rows, skinning transforms): cheaper than keeping them
    if (m->depth && m->child_start && !g_thread_pose)
      lbs_pose_warp_kernel<<<cdiv(N, PW), PW * 32, 0, st>>>(*m, N, fpb, root_orient, pose_body, betas, trans, ws.feat, ws.A, nullptr, 52,
                                                           nullptr, nullptr, 1.f, 0);
    else
      lbs_pose_kernel<<<cdiv(N, 64), 64, 0, st>>>(*m, N, fpb, root_orient, pose_body, betas, trans, ws.feat, ws.A, nullptr, 52, nullptr, nullptr, 1.f, 0);
    HB_LAUNCH_CHECK(); ++nl;
    static bool attr_bwd = false;
    if (!attr_bwd) {
      HB_CUDA(cudaFuncSetAttribute(lbs_skin_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKIN_BWD_SMEM));
      attr_bwd = true;
    }
    int acc = 0;
    if (d_verts && vlist && xj) {
      // listed key vertices + the 21 vertex-picked joints in ONE pass (43 + 21 = 64 = one chunk on the Stage-III path)
      lbs_skin_bwd_kernel<<<cdiv(N, BW_FT), 256, SKIN_BWD_SMEM, st>>>(*m, N, ws.feat, ws.A, vlist, nv + 21, d_verts, (size_t)nv * 3,
                                                                     ws.dfeat, ws.dA, ws.dtr, 0, m->extra_ids, nv, d_joints + 52 * 3,
                                                                     (size_t)73 * 3);
      HB_LAUNCH_CHECK(); ++nl;
    } else {
      if (d_verts) {
        const int nvv = vlist ? nv : m->num_verts;
        lbs_skin_bwd_kernel<<<cdiv(N, BW_FT), 256, SKIN_BWD_SMEM, st>>>(*m, N, ws.feat, ws.A, vlist, nvv, d_verts, (size_t)nvv * 3,
                                                                       ws.dfeat, ws.dA, ws.dtr, acc, nullptr, nvv, nullptr, 0);
        HB_LAUNCH_CHECK(); ++nl;
        acc = 1;
      }
      if (xj) {
        lbs_skin_bwd_kernel<<<cdiv(N, BW_FT), 256, SKIN_BWD_SMEM, st>>>(*m, N, ws.feat, ws.A, m->extra_ids, 21, d_joints + 52 * 3,
                                                                       (size_t)73 * 3, ws.dfeat, ws.dA, ws.dtr, acc, nullptr, 21, nullptr, 0);
        HB_LAUNCH_CHECK(); ++nl;
      }
    }
  }
  if (m->depth && m->child_start && !g_thread_pose)
    lbs_pose_bwd_warp_kernel<<<cdiv(N, PW), PW * 32, 0, st>>>(*m, N, fpb, root_orient, pose_body, betas, need_skin ? ws.dfeat : nullptr,
                                                             need_skin ? ws.dA : nullptr, need_skin ? ws.dtr : nullptr, d_joints, njo,
                                                             d_root_orient, d_pose_body, d_betas, d_trans);
  else
    lbs_pose_bwd_kernel<<<cdiv(N, 64), 64, 0, st>>>(*m, N, fpb, root_orient, pose_body, betas, need_skin ? ws.dfeat : nullptr,
                                                   need_skin ? ws.dA : nullptr, need_skin ? ws.dtr : nullptr, d_joints, njo,
                                                   d_root_orient, d_pose_body, d_betas, d_trans);
  HB_LAUNCH_CHECK(); ++nl;
  if (launches) *launches = nl;
  return HB_OK;
}
#endif  // HB_HOST_SHIM
