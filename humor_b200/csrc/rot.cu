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
extern "C" const char* humor_b200_version(void) { return "humor_b200 0.1 (sm_100a)"; }
#endif  // HB_HOST_SHIM
