// Batched rotation conversions and their reverse modes (utils/transforms.py:139-170, :243-389).
#include "common.cuh"
#include "geom.cuh"
#include "../../include/humor_b200.h"

namespace hb {
__global__ void rodrigues_fwd_kernel(int n, const float* __restrict__ aa, float* R) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r[3] = {aa[3 * (size_t)i], aa[3 * (size_t)i + 1], aa[3 * (size_t)i + 2]}, M[9];
  rodrigues_fwd(r, M);
#pragma unroll
  for (int e = 0; e < 9; ++e) R[9 * (size_t)i + e] = M[e];
}
__global__ void rodrigues_bwd_kernel(int n, const float* __restrict__ aa, const float* __restrict__ dR, float* daa) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r[3] = {aa[3 * (size_t)i], aa[3 * (size_t)i + 1], aa[3 * (size_t)i + 2]}, G[9], d[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 9; ++e) G[e] = dR[9 * (size_t)i + e];
  rodrigues_bwd(r, G, d);
  daa[3 * (size_t)i] = d[0]; daa[3 * (size_t)i + 1] = d[1]; daa[3 * (size_t)i + 2] = d[2];
}
__global__ void mat2aa_fwd_kernel(int n, const float* __restrict__ R, float* aa) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float M[9], a[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) M[e] = R[9 * (size_t)i + e];
  mat2aa_fwd(M, a);
  aa[3 * (size_t)i] = a[0]; aa[3 * (size_t)i + 1] = a[1]; aa[3 * (size_t)i + 2] = a[2];
}
__global__ void mat2aa_bwd_kernel(int n, const float* __restrict__ R, const float* __restrict__ daa, float* dR) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float M[9], G[9], d[3] = {daa[3 * (size_t)i], daa[3 * (size_t)i + 1], daa[3 * (size_t)i + 2]};
#pragma unroll
  for (int e = 0; e < 9; ++e) { M[e] = R[9 * (size_t)i + e]; G[e] = 0.f; }
  mat2aa_bwd(M, d, G);
#pragma unroll
  for (int e = 0; e < 9; ++e) dR[9 * (size_t)i + e] = G[e];
}

// ------------------------------------------------------------------------------------------------
// camera -> prior frame of one sub-sequence (humor/fitting/fitting_utils.py:61-103 parse_floor_plane, :149-190
// compute_cam2prior), once per closure and sequence.  As ~60 torch ops + their ~150 autograd nodes this was the largest group of
// tiny launches left in the Stage-III graph; here: one thread per sequence, forward and hand-written reverse.
//   plane   N = sgn f/|f|, D = sgn |f|, sgn = -1 if f_y/|f| > 0 (normal up = -y in camera space)
//   up = N;  a = D - N.t0;  br = -Rodrigues(r0)[:,0];  c = N.br;  s = a/c
//   right = normalise(sign(s) a (br/c - N));  fwd = normalise(up x right);  R = rows (right, fwd, up)
//   t = -t0;  root_height = N.j0 - D
// ------------------------------------------------------------------------------------------------
struct Cam2PriorFwd {
  float N[3], D, sgn, invf, a, c, br[3], sg, u[3], q[3], qn, right[3], w[3], wn, fwd[3];
};
HD void cam2prior_eval(const float* f, const float* t0, const float* r0, Cam2PriorFwd& v) {
  const float fn = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  v.invf = 1.f / fn;
  v.sgn = (f[1] * v.invf > 0.f) ? -1.f : 1.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) v.N[i] = f[i] * v.invf * v.sgn;
  v.D = fn * v.sgn;
  v.a = v.D - (v.N[0] * t0[0] + v.N[1] * t0[1] + v.N[2] * t0[2]);
  float Rm[9];
  rodrigues_fwd(r0, Rm);
  v.br[0] = -Rm[0]; v.br[1] = -Rm[3]; v.br[2] = -Rm[6];
  v.c = v.N[0] * v.br[0] + v.N[1] * v.br[1] + v.N[2] * v.br[2];
  const float s = v.a / v.c;
  v.sg = (s < 0.f) ? -1.f : 1.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    // (hit - foot) = (t0 + s br) - (t0 + a N), formed as the reference forms it: through the two intersection points
    const float hit = t0[i] + s * v.br[i], foot = t0[i] + (-v.a) * (-v.N[i]);
    v.u[i] = v.br[i] / v.c - v.N[i];
    v.q[i] = (hit - foot) * v.sg;
  }
  v.qn = sqrtf(v.q[0] * v.q[0] + v.q[1] * v.q[1] + v.q[2] * v.q[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) v.right[i] = v.q[i] / v.qn;
  v.w[0] = v.N[1] * v.right[2] - v.N[2] * v.right[1];
  v.w[1] = v.N[2] * v.right[0] - v.N[0] * v.right[2];
  v.w[2] = v.N[0] * v.right[1] - v.N[1] * v.right[0];
  v.wn = sqrtf(v.w[0] * v.w[0] + v.w[1] * v.w[1] + v.w[2] * v.w[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) v.fwd[i] = v.w[i] / v.wn;
}
__global__ void cam2prior_fwd_kernel(int B, const float* __restrict__ floor, const float* __restrict__ trans0, int ld_t,
                                     const float* __restrict__ orient0, int ld_r, const float* __restrict__ joint0, int ld_j,
                                     float* __restrict__ R, float* __restrict__ t, float* __restrict__ h) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float f[3] = {floor[3 * b], floor[3 * b + 1], floor[3 * b + 2]};
  const float t0[3] = {trans0[(size_t)b * ld_t], trans0[(size_t)b * ld_t + 1], trans0[(size_t)b * ld_t + 2]};
  const float r0[3] = {orient0[(size_t)b * ld_r], orient0[(size_t)b * ld_r + 1], orient0[(size_t)b * ld_r + 2]};
  const float j0[3] = {joint0[(size_t)b * ld_j], joint0[(size_t)b * ld_j + 1], joint0[(size_t)b * ld_j + 2]};
  Cam2PriorFwd v;
  cam2prior_eval(f, t0, r0, v);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    R[9 * b + i] = v.right[i]; R[9 * b + 3 + i] = v.fwd[i]; R[9 * b + 6 + i] = v.N[i];
    t[3 * b + i] = -t0[i];
  }
  // s_root = (D - N.j0) / (N.(-N)), as the reference forms it
  const float nn = -(v.N[0] * v.N[0] + v.N[1] * v.N[1] + v.N[2] * v.N[2]);
  h[b] = (v.D - (v.N[0] * j0[0] + v.N[1] * j0[1] + v.N[2] * j0[2])) / nn;
}
// reverse: gR [B][9], gt [B][3], gh [B] (any of them nullable = zero) -> d floor, d trans0, d orient0, d joint0, all [B][3]
__global__ void cam2prior_bwd_kernel(int B, const float* __restrict__ floor, const float* __restrict__ trans0, int ld_t,
                                     const float* __restrict__ orient0, int ld_r, const float* __restrict__ joint0, int ld_j,
                                     const float* __restrict__ gR, const float* __restrict__ gt, const float* __restrict__ gh,
                                     float* __restrict__ d_floor, float* __restrict__ d_trans0, float* __restrict__ d_orient0,
                                     float* __restrict__ d_joint0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float f[3] = {floor[3 * b], floor[3 * b + 1], floor[3 * b + 2]};
  const float t0[3] = {trans0[(size_t)b * ld_t], trans0[(size_t)b * ld_t + 1], trans0[(size_t)b * ld_t + 2]};
  const float r0[3] = {orient0[(size_t)b * ld_r], orient0[(size_t)b * ld_r + 1], orient0[(size_t)b * ld_r + 2]};
  const float j0[3] = {joint0[(size_t)b * ld_j], joint0[(size_t)b * ld_j + 1], joint0[(size_t)b * ld_j + 2]};
  Cam2PriorFwd v;
  cam2prior_eval(f, t0, r0, v);
  float gN[3] = {0.f, 0.f, 0.f}, gD = 0.f, g_t0[3] = {0.f, 0.f, 0.f}, g_j0[3] = {0.f, 0.f, 0.f};
  if (gt) for (int i = 0; i < 3; ++i) g_t0[i] = -gt[3 * b + i];
  if (gh) {                                     // h = N.j0 - D  (|N| = 1: the reference's denominator N.(-N) carries no gradient beyond rounding)
    const float g = gh[b];
    for (int i = 0; i < 3; ++i) { gN[i] += g * j0[i]; g_j0[i] = g * v.N[i]; }
    gD -= g;
  }
  float g_br[3] = {0.f, 0.f, 0.f};
  if (gR) {
    float g_right[3], g_fwd[3];
    for (int i = 0; i < 3; ++i) { g_right[i] = gR[9 * b + i]; g_fwd[i] = gR[9 * b + 3 + i]; gN[i] += gR[9 * b + 6 + i]; }
    // fwd = w/|w|, w = N x right
    const float fd = v.fwd[0] * g_fwd[0] + v.fwd[1] * g_fwd[1] + v.fwd[2] * g_fwd[2];
    float g_w[3];
    for (int i = 0; i < 3; ++i) g_w[i] = (g_fwd[i] - v.fwd[i] * fd) / v.wn;
    // w = a x b: g_a = b x g_w, g_b = g_w x a
    gN[0] += v.right[1] * g_w[2] - v.right[2] * g_w[1];
    gN[1] += v.right[2] * g_w[0] - v.right[0] * g_w[2];
    gN[2] += v.right[0] * g_w[1] - v.right[1] * g_w[0];
    g_right[0] += g_w[1] * v.N[2] - g_w[2] * v.N[1];
    g_right[1] += g_w[2] * v.N[0] - g_w[0] * v.N[2];
    g_right[2] += g_w[0] * v.N[1] - g_w[1] * v.N[0];
    // right = q/|q|, q = sg a u, u = br/c - N
    const float rd = v.right[0] * g_right[0] + v.right[1] * g_right[1] + v.right[2] * g_right[2];
    float g_q[3], g_u[3];
    for (int i = 0; i < 3; ++i) g_q[i] = (g_right[i] - v.right[i] * rd) / v.qn;
    const float g_a = v.sg * (g_q[0] * v.u[0] + g_q[1] * v.u[1] + g_q[2] * v.u[2]);
    for (int i = 0; i < 3; ++i) g_u[i] = v.sg * v.a * g_q[i];
    float g_c = -(g_u[0] * v.br[0] + g_u[1] * v.br[1] + g_u[2] * v.br[2]) / (v.c * v.c);
    for (int i = 0; i < 3; ++i) { g_br[i] += g_u[i] / v.c; gN[i] -= g_u[i]; }
    for (int i = 0; i < 3; ++i) { gN[i] += g_c * v.br[i]; g_br[i] += g_c * v.N[i]; }
    // a = D - N.t0
    gD += g_a;
    for (int i = 0; i < 3; ++i) { gN[i] -= g_a * t0[i]; g_t0[i] -= g_a * v.N[i]; }
  }
  // br = -Rodrigues(r0)[:,0]
  float G[9] = {-g_br[0], 0.f, 0.f, -g_br[1], 0.f, 0.f, -g_br[2], 0.f, 0.f}, dr[3] = {0.f, 0.f, 0.f};
  rodrigues_bwd(r0, G, dr);
  // N = sgn f/|f|, D = sgn |f|
  const float nh[3] = {f[0] * v.invf, f[1] * v.invf, f[2] * v.invf};
  const float nd = nh[0] * gN[0] + nh[1] * gN[1] + nh[2] * gN[2];
  for (int i = 0; i < 3; ++i) {
    d_floor[3 * b + i] = v.sgn * ((gN[i] - nh[i] * nd) * v.invf + gD * nh[i]);
    d_trans0[3 * b + i] = g_t0[i];
    d_orient0[3 * b + i] = dr[i];
    d_joint0[3 * b + i] = g_j0[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Outputs of the latent roll-out in the layout the energies read (humor/fitting/motion_optimizer.py:964-1019, :678-741):
// world rows (S,B,348) of the CVAE chain + the frame-0 state -> (B,T,.) prior-frame translation / root orientation / body pose
// (axis-angle) / joints / contact logits, contact confidences and labels, and - with a camera->prior transform (R, t) - the
// camera-frame root orientation and translation of apply_cam2prior(inverse=True).  As torch ops this was a permute, 5 slices, the
// matrix->axis-angle kernel, 4 concatenations, sigmoid / index_copy / cat / compare, and Rodrigues + 2 batched matmuls + matrix->
// axis-angle for the camera frame; in reverse 5 zero-filled (S,B,348) buffers added together.  One block per sub-sequence.
// World columns: 0 trans | 3 trans_vel | 6 root R | 15 root vel | 18 body R (21) | 207 joints | 273 joint vel | 339 contact logits.
// ------------------------------------------------------------------------------------------------
constexpr int RO_LD = 348;
constexpr int RO_THREADS = 128;
struct RollOut {
  int B, S;
  const float* world;      // [S][B][348]
  const float* trans0;     // [B][3]   frame 0 (prior frame)
  const float* orient0;    // [B][3]
  const float* pose0;      // [B][63]
  const float* joints0;    // [B][66]
  const float* R;          // [B][9] rows right/forward/up of the prior frame in camera coordinates, or NULL (no camera frame)
  const float* t;          // [B][3]
  const int* contact_idx;  // [9] joint of every contact logit
  float thresh;
  float *trans, *orient, *pose, *joints;     // [B][T][3|3|63|66]
  float *logits;                              // [B][S][9]
  float *conf, *labels;                       // [B][T][22]
  float *cam_trans, *cam_orient;              // [B][T][3] (R != NULL)
};
__global__ void __launch_bounds__(RO_THREADS) rollout_outputs_fwd_kernel(RollOut a) {
  const int b = blockIdx.x, tid = threadIdx.x, T = a.S + 1;
  const float* R = a.R ? a.R + 9 * b : nullptr;
  for (int idx = tid; idx < T * 22; idx += RO_THREADS) {         // rotations: root (j = 0) and the 21 body joints
    const int t = idx / 22, j = idx - 22 * t;
    float aa[3];
    if (t == 0) {
      const float* src = j == 0 ? a.orient0 + 3 * b : a.pose0 + 63 * b + 3 * (j - 1);
      aa[0] = src[0]; aa[1] = src[1]; aa[2] = src[2];
    } else {
      const float* w = a.world + ((size_t)(t - 1) * a.B + b) * RO_LD + (j == 0 ? 6 : 18 + 9 * (j - 1));
      float M[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) M[e] = w[e];
      mat2aa_fwd(M, aa);
    }
    float* dst = j == 0 ? a.orient + ((size_t)b * T + t) * 3 : a.pose + ((size_t)b * T + t) * 63 + 3 * (j - 1);
    dst[0] = aa[0]; dst[1] = aa[1]; dst[2] = aa[2];
    if (j == 0 && R) {                                          // camera frame: R^T Rodrigues(aa) -> axis-angle
      float Rq[9], Mc[9], ca[3];
      rodrigues_fwd(aa, Rq);
      mat3_mul_tn(R, Rq, Mc);
      mat2aa_fwd(Mc, ca);
      float* o = a.cam_orient + ((size_t)b * T + t) * 3;
      o[0] = ca[0]; o[1] = ca[1]; o[2] = ca[2];
    }
  }
  for (int idx = tid; idx < T * 66; idx += RO_THREADS) {         // joints
    const int t = idx / 66, e = idx - 66 * t;
    a.joints[((size_t)b * T + t) * 66 + e] = t == 0 ? a.joints0[66 * b + e] : a.world[((size_t)(t - 1) * a.B + b) * RO_LD + 207 + e];
  }
  for (int t = tid; t < T; t += RO_THREADS) {                    // translation, prior and camera frame
    float p[3], p0[3] = {a.trans0[3 * b], a.trans0[3 * b + 1], a.trans0[3 * b + 2]};
    for (int i = 0; i < 3; ++i) p[i] = t == 0 ? p0[i] : a.world[((size_t)(t - 1) * a.B + b) * RO_LD + i];
    float* o = a.trans + ((size_t)b * T + t) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    if (R) {
      const float d[3] = {p[0] - p0[0], p[1] - p0[1], p[2] - p0[2]};
      float c[3];
      mat3_tvec(R, d, c);
      float* oc = a.cam_trans + ((size_t)b * T + t) * 3;
      for (int i = 0; i < 3; ++i) oc[i] = c[i] - a.t[3 * b + i];
    }
  }
  for (int idx = tid; idx < a.S * 9; idx += RO_THREADS) {        // contact logits
    const int s = idx / 9, c = idx - 9 * s;
    a.logits[((size_t)b * a.S + s) * 9 + c] = a.world[((size_t)s * a.B + b) * RO_LD + 339 + c];
  }
  for (int idx = tid; idx < T * 22; idx += RO_THREADS) {         // confidences on the 22 joints (frame 0 repeats step 0), labels
    const int t = idx / 22, j = idx - 22 * t;
    const int s = t == 0 ? 0 : t - 1;
    float cf = 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c)
      if (a.contact_idx[c] == j) cf = 1.f / (1.f + expf(-a.world[((size_t)s * a.B + b) * RO_LD + 339 + c]));
    a.conf[((size_t)b * T + t) * 22 + j] = cf;
    a.labels[((size_t)b * T + t) * 22 + j] = cf > a.thresh ? 1.f : 0.f;
  }
}
struct RollOutBwd {
  int B, S;
  const float* world; const float* trans0; const float* orient0; const float* R;      // forward inputs the reverse needs
  const float *g_trans, *g_orient, *g_pose, *g_joints, *g_logits, *g_cam_trans, *g_cam_orient;   // nullable = zero
  float* d_world;          // [S][B][348], every element written
  float *d_trans0, *d_orient0, *d_pose0, *d_joints0;   // [B][3|3|63|66]
  float *d_R, *d_t;        // [B][9], [B][3] (R != NULL)
};
__global__ void __launch_bounds__(RO_THREADS) rollout_outputs_bwd_kernel(RollOutBwd a) {
  __shared__ float red[RO_THREADS][15];                          // per-thread partials of d R (9), d t (3), d trans0 via the camera frame (3)
  const int b = blockIdx.x, tid = threadIdx.x, T = a.S + 1;
  const float* R = a.R ? a.R + 9 * b : nullptr;
  float acc[15];
#pragma unroll
  for (int e = 0; e < 15; ++e) acc[e] = 0.f;
  for (int idx = tid; idx < T * 22; idx += RO_THREADS) {         // rotations
    const int t = idx / 22, j = idx - 22 * t;
    const float* gsrc = j == 0 ? (a.g_orient ? a.g_orient + ((size_t)b * T + t) * 3 : nullptr)
                               : (a.g_pose ? a.g_pose + ((size_t)b * T + t) * 63 + 3 * (j - 1) : nullptr);
    float g[3] = {gsrc ? gsrc[0] : 0.f, gsrc ? gsrc[1] : 0.f, gsrc ? gsrc[2] : 0.f};
    const float* w = t == 0 ? nullptr : a.world + ((size_t)(t - 1) * a.B + b) * RO_LD + (j == 0 ? 6 : 18 + 9 * (j - 1));
    float M[9];
    if (w) {
#pragma unroll
      for (int e = 0; e < 9; ++e) M[e] = w[e];
    }
    if (j == 0 && R && a.g_cam_orient) {
      float aa[3];
      if (w) mat2aa_fwd(M, aa);
      else { aa[0] = a.orient0[3 * b]; aa[1] = a.orient0[3 * b + 1]; aa[2] = a.orient0[3 * b + 2]; }
      const float* gc = a.g_cam_orient + ((size_t)b * T + t) * 3;
      float Rq[9], Mc[9], gM[9], gRq[9], ga[3] = {0.f, 0.f, 0.f};
      rodrigues_fwd(aa, Rq);
      mat3_mul_tn(R, Rq, Mc);
#pragma unroll
      for (int e = 0; e < 9; ++e) gM[e] = 0.f;
      mat2aa_bwd(Mc, gc, gM);
      mat3_mul(R, gM, gRq);                                      // Mc = R^T Rq: d Rq = R d Mc, d R = Rq d Mc^T
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[3 * i + k] += Rq[3 * i] * gM[3 * k] + Rq[3 * i + 1] * gM[3 * k + 1] + Rq[3 * i + 2] * gM[3 * k + 2];
      rodrigues_bwd(aa, gRq, ga);
      g[0] += ga[0]; g[1] += ga[1]; g[2] += ga[2];
    }
    if (t == 0) {
      float* d = j == 0 ? a.d_orient0 + 3 * b : a.d_pose0 + 63 * b + 3 * (j - 1);
      d[0] = g[0]; d[1] = g[1]; d[2] = g[2];
    } else {
      float gW[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) gW[e] = 0.f;
      mat2aa_bwd(M, g, gW);
      float* d = a.d_world + ((size_t)(t - 1) * a.B + b) * RO_LD + (j == 0 ? 6 : 18 + 9 * (j - 1));
#pragma unroll
      for (int e = 0; e < 9; ++e) d[e] = gW[e];
    }
  }
  for (int idx = tid; idx < T * 66; idx += RO_THREADS) {         // joints; velocity columns carry no gradient
    const int t = idx / 66, e = idx - 66 * t;
    const float g = a.g_joints ? a.g_joints[((size_t)b * T + t) * 66 + e] : 0.f;
    if (t == 0) a.d_joints0[66 * b + e] = g;
    else {
      float* d = a.d_world + ((size_t)(t - 1) * a.B + b) * RO_LD;
      d[207 + e] = g;
      d[273 + e] = 0.f;
    }
  }
  for (int idx = tid; idx < a.S * 15; idx += RO_THREADS) {       // trans_vel (3), root velocity (3), contact logits (9)
    const int s = idx / 15, e = idx - 15 * s;
    float* d = a.d_world + ((size_t)s * a.B + b) * RO_LD;
    if (e < 3) d[3 + e] = 0.f;
    else if (e < 6) d[15 + e - 3] = 0.f;
    else d[339 + e - 6] = a.g_logits ? a.g_logits[((size_t)b * a.S + s) * 9 + e - 6] : 0.f;
  }
  float g0[3] = {0.f, 0.f, 0.f};                                  // d trans0 of the thread that owns frame 0
  for (int t = tid; t < T; t += RO_THREADS) {                    // translation
    float g[3];
    for (int i = 0; i < 3; ++i) g[i] = a.g_trans ? a.g_trans[((size_t)b * T + t) * 3 + i] : 0.f;
    if (R && a.g_cam_trans) {
      const float* gc = a.g_cam_trans + ((size_t)b * T + t) * 3;
      float p[3], d[3], gd[3];
      for (int i = 0; i < 3; ++i) {
        p[i] = t == 0 ? a.trans0[3 * b + i] : a.world[((size_t)(t - 1) * a.B + b) * RO_LD + i];
        d[i] = p[i] - a.trans0[3 * b + i];
      }
      mat3_vec(R, gc, gd);                                       // c = R^T d - t: d d = R g, d R[i][k] += d[i] g[k], d t = -g
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[3 * i + k] += d[i] * gc[k];
        acc[9 + i] -= gc[i];
        acc[12 + i] -= gd[i];                                    // through d = p - p0 into frame 0
        g[i] += gd[i];
      }
    }
    if (t == 0) { g0[0] = g[0]; g0[1] = g[1]; g0[2] = g[2]; }
    else {
      float* d = a.d_world + ((size_t)(t - 1) * a.B + b) * RO_LD;
      d[0] = g[0]; d[1] = g[1]; d[2] = g[2];
    }
  }
#pragma unroll
  for (int e = 0; e < 15; ++e) red[tid][e] = acc[e];
  __syncthreads();
  if (tid < 15) {                                                // fixed order: deterministic
    float s = 0.f;
    for (int i = 0; i < RO_THREADS; ++i) s += red[i][tid];
    red[0][tid] = s;
  }
  __syncthreads();
  if (tid == 0) {                                                // thread 0 owns frame 0
    for (int i = 0; i < 3; ++i) a.d_trans0[3 * b + i] = g0[i] + red[0][12 + i];
    if (R) {
      for (int e = 0; e < 9; ++e) a.d_R[9 * b + e] = red[0][e];
      for (int i = 0; i < 3; ++i) a.d_t[3 * b + i] = red[0][9 + i];
    }
  }
}
}  // namespace hb
#ifndef HB_HOST_SHIM   // host side of the C-ABI (launch syntax): device builds only
using namespace hb;
extern "C" int humor_rodrigues_fwd(int n, const float* aa, float* R, cudaStream_t st) {
  if (n <= 0 || !aa || !R) return HB_ERR_ARG;
  rodrigues_fwd_kernel<<<cdiv(n, 128), 128, 0, st>>>(n, aa, R);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" int humor_rodrigues_bwd(int n, const float* aa, const float* dR, float* daa, cudaStream_t st) {
  if (n <= 0 || !aa || !dR || !daa) return HB_ERR_ARG;
  rodrigues_bwd_kernel<<<cdiv(n, 128), 128, 0, st>>>(n, aa, dR, daa);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" int humor_mat2aa_fwd(int n, const float* R, float* aa, cudaStream_t st) {
  if (n <= 0 || !aa || !R) return HB_ERR_ARG;
  mat2aa_fwd_kernel<<<cdiv(n, 128), 128, 0, st>>>(n, R, aa);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" int humor_mat2aa_bwd(int n, const float* R, const float* daa, float* dR, cudaStream_t st) {
  if (n <= 0 || !R || !daa || !dR) return HB_ERR_ARG;
  mat2aa_bwd_kernel<<<cdiv(n, 128), 128, 0, st>>>(n, R, daa, dR);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" int humor_cam2prior_fwd(int B, const float* floor_plane, const float* trans0, int ld_t, const float* orient0, int ld_r,
                                   const float* joint0, int ld_j, float* R, float* t, float* root_height, cudaStream_t st) {
  if (B <= 0 || !floor_plane || !trans0 || !orient0 || !joint0 || !R || !t || !root_height || ld_t < 3 || ld_r < 3 || ld_j < 3) return HB_ERR_ARG;
  cam2prior_fwd_kernel<<<cdiv(B, 64), 64, 0, st>>>(B, floor_plane, trans0, ld_t, orient0, ld_r, joint0, ld_j, R, t, root_height);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" int humor_cam2prior_bwd(int B, const float* floor_plane, const float* trans0, int ld_t, const float* orient0, int ld_r,
                                   const float* joint0, int ld_j, const float* gR, const float* gt, const float* gh, float* d_floor,
                                   float* d_trans0, float* d_orient0, float* d_joint0, cudaStream_t st) {
  if (B <= 0 || !floor_plane || !trans0 || !orient0 || !joint0 || !d_floor || !d_trans0 || !d_orient0 || !d_joint0 || ld_t < 3 || ld_r < 3 ||
      ld_j < 3)
    return HB_ERR_ARG;
  cam2prior_bwd_kernel<<<cdiv(B, 64), 64, 0, st>>>(B, floor_plane, trans0, ld_t, orient0, ld_r, joint0, ld_j, gR, gt, gh, d_floor, d_trans0,
                                                  d_orient0, d_joint0);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" int humor_rollout_outputs_fwd(int B, int S, const float* world, const float* trans0, const float* orient0, const float* pose0,
                                         const float* joints0, const float* R, const float* t, const int* contact_idx, float thresh,
                                         float* trans, float* orient, float* pose, float* joints, float* logits, float* conf,
                                         float* labels, float* cam_trans, float* cam_orient, cudaStream_t st) {
  if (B <= 0 || S <= 0 || !world || !trans0 || !orient0 || !pose0 || !joints0 || !contact_idx || !trans || !orient || !pose || !joints ||
      !logits || !conf || !labels || ((R != nullptr) != (t != nullptr)) || (R && (!cam_trans || !cam_orient)))
    return HB_ERR_ARG;
  RollOut a;
  a.B = B; a.S = S; a.world = world; a.trans0 = trans0; a.orient0 = orient0; a.pose0 = pose0; a.joints0 = joints0; a.R = R; a.t = t;
  a.contact_idx = contact_idx; a.thresh = thresh; a.trans = trans; a.orient = orient; a.pose = pose; a.joints = joints; a.logits = logits;
  a.conf = conf; a.labels = labels; a.cam_trans = cam_trans; a.cam_orient = cam_orient;
  rollout_outputs_fwd_kernel<<<B, RO_THREADS, 0, st>>>(a);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" int humor_rollout_outputs_bwd(int B, int S, const float* world, const float* trans0, const float* orient0, const float* R,
                                         const float* g_trans, const float* g_orient, const float* g_pose, const float* g_joints,
                                         const float* g_logits, const float* g_cam_trans, const float* g_cam_orient, float* d_world,
                                         float* d_trans0, float* d_orient0, float* d_pose0, float* d_joints0, float* d_R, float* d_t,
                                         cudaStream_t st) {
  if (B <= 0 || S <= 0 || !world || !trans0 || !orient0 || !d_world || !d_trans0 || !d_orient0 || !d_pose0 || !d_joints0 ||
      (R && (!d_R || !d_t)))
    return HB_ERR_ARG;
  RollOutBwd a;
  a.B = B; a.S = S; a.world = world; a.trans0 = trans0; a.orient0 = orient0; a.R = R; a.g_trans = g_trans; a.g_orient = g_orient;
  a.g_pose = g_pose; a.g_joints = g_joints; a.g_logits = g_logits; a.g_cam_trans = R ? g_cam_trans : nullptr;
  a.g_cam_orient = R ? g_cam_orient : nullptr; a.d_world = d_world; a.d_trans0 = d_trans0; a.d_orient0 = d_orient0; a.d_pose0 = d_pose0;
  a.d_joints0 = d_joints0; a.d_R = d_R; a.d_t = d_t;
  rollout_outputs_bwd_kernel<<<B, RO_THREADS, 0, st>>>(a);
  HB_LAUNCH_CHECK(); return HB_OK;
}
extern "C" const char* humor_b200_version(void) { return "humor_b200 0.1 (sm_100a)"; }
#endif  // HB_HOST_SHIM
