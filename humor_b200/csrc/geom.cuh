// Per-element geometry used by the HuMoR Stage-III kernels: forward + hand-derived reverse mode.
// Every function is __host__ __device__ so the exact code the kernels run is also compiled with
// g++ (tests/host/geom_host.cpp) and checked against torch autograd on CPU.
//
// Reference semantics followed (paths relative to /root/reference/humor):
//   rodrigues        utils/transforms.py:139-170   (angle = ||r + 1e-8||)
//   mat2aa           utils/transforms.py:243-389   (4-branch quaternion, atan2, NaN -> 0)
//   world2aligned    utils/transforms.py:17-42     (acos(clamp(x/(||xy||+1e-6))), +1e-6 normaliser)
// Matrices are row-major float[9].
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace hb {

HD void mat3_mul(const float* A, const float* B, float* C) {          // C = A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
HD void mat3_mul_tn(const float* A, const float* B, float* C) {       // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[0 * 3 + i] * B[0 * 3 + j] + A[1 * 3 + i] * B[1 * 3 + j] + A[2 * 3 + i] * B[2 * 3 + j];
}
HD void mat3_mul_nt(const float* A, const float* B, float* C) {       // C = A B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
HD void mat3_vec(const float* A, const float* v, float* o) {          // o = A v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
HD void mat3_tvec(const float* A, const float* v, float* o) {         // o = A^T v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
// given C = A B and dC: dA += dC B^T, dB += A^T dC
HD void mat3_mul_bwd(const float* A, const float* B, const float* dC, float* dA, float* dB) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a += dC[i * 3 + k] * B[j * 3 + k];
        b += A[k * 3 + i] * dC[k * 3 + j];
      }
      if (dA) dA[i * 3 + j] += a;
      if (dB) dB[i * 3 + j] += b;
    }
}

// ------------------------------------------------------------------------------------------------
// axis-angle -> matrix
// ------------------------------------------------------------------------------------------------
HD void rodrigues_fwd(const float* r, float* R) {
  const float e = 1e-8f;
  float ax = r[0] + e, ay = r[1] + e, az = r[2] + e;
  float th = sqrtf(ax * ax + ay * ay + az * az);
  float dx = r[0] / th, dy = r[1] / th, dz = r[2] / th;
  float s = sinf(th), c1 = 1.f - cosf(th);
  float dd = dx * dx + dy * dy + dz * dz;
  // K = [[0,-dz,dy],[dz,0,-dx],[-dy,dx,0]] ; K^2 = d d^T - |d|^2 I
  R[0] = 1.f + c1 * (dx * dx - dd);
  R[1] = -s * dz + c1 * dx * dy;
  R[2] = s * dy + c1 * dx * dz;
  R[3] = s * dz + c1 * dx * dy;
  R[4] = 1.f + c1 * (dy * dy - dd);
  R[5] = -s * dx + c1 * dy * dz;
  R[6] = -s * dy + c1 * dx * dz;
  R[7] = s * dx + c1 * dy * dz;
  R[8] = 1.f + c1 * (dz * dz - dd);
}

// dr += J^T dR
HD void rodrigues_bwd(const float* r, const float* G, float* dr) {
  const float e = 1e-8f;
  float ax = r[0] + e, ay = r[1] + e, az = r[2] + e;
  float th = sqrtf(ax * ax + ay * ay + az * az);
  float inv = 1.f / th;
  float d[3] = {r[0] * inv, r[1] * inv, r[2] * inv};
  float s = sinf(th), c = cosf(th), c1 = 1.f - c;
  float dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  float trG = G[0] + G[4] + G[8];
  // sum G .* K and sum G .* K^2
  float gK = d[0] * (G[7] - G[5]) + d[1] * (G[2] - G[6]) + d[2] * (G[3] - G[1]);
  float gK2 = -dd * trG;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) gK2 += G[i * 3 + j] * d[i] * d[j];
  float g_th = c * gK + s * gK2;
  float gd[3];
  gd[0] = s * (G[7] - G[5]);
  gd[1] = s * (G[2] - G[6]);
  gd[2] = s * (G[3] - G[1]);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) a += (G[k * 3 + j] + G[j * 3 + k]) * d[j];
    gd[k] += c1 * (a - 2.f * d[k] * trG);
  }
  // d = r / th
  float g_th_d = -(gd[0] * r[0] + gd[1] * r[1] + gd[2] * r[2]) * inv * inv;
  float gt = g_th + g_th_d;
  dr[0] += gd[0] * inv + gt * ax * inv;
  dr[1] += gd[1] * inv + gt * ay * inv;
  dr[2] += gd[2] * inv + gt * az * inv;
}

// ------------------------------------------------------------------------------------------------
// matrix -> axis-angle
// ------------------------------------------------------------------------------------------------
struct Mat2AACtx {
  int branch;
  float qraw[4], t, q[4], s2, s, k, two_theta;
};

HD void mat2aa_core(const float* R, float* aa, Mat2AACtx& cx) {
  // the reference indexes the transposed matrix m = R^T : m[i][j] = R[j][i]
  float m00 = R[0], m11 = R[4], m22 = R[8];
  float m01 = R[3], m10 = R[1], m02 = R[6], m20 = R[2], m12 = R[7], m21 = R[5];
  bool d2 = m22 < 1e-6f, d01 = m00 > m11, d0n1 = m00 < -m11;
  float* q = cx.qraw;
  if (d2 && d01) {
    cx.branch = 0; cx.t = 1.f + m00 - m11 - m22;
    q[0] = m12 - m21; q[1] = cx.t; q[2] = m01 + m10; q[3] = m20 + m02;
  } else if (d2) {
    cx.branch = 1; cx.t = 1.f - m00 + m11 - m22;
    q[0] = m20 - m02; q[1] = m01 + m10; q[2] = cx.t; q[3] = m12 + m21;
  } else if (d0n1) {
    cx.branch = 2; cx.t = 1.f - m00 - m11 + m22;
    q[0] = m01 - m10; q[1] = m20 + m02; q[2] = m12 + m21; q[3] = cx.t;
  } else {
    cx.branch = 3; cx.t = 1.f + m00 + m11 + m22;
    q[0] = cx.t; q[1] = m12 - m21; q[2] = m20 - m02; q[3] = m01 - m10;
  }
  float sc = 0.5f / sqrtf(cx.t);
#pragma unroll
  for (int i = 0; i < 4; ++i) cx.q[i] = q[i] * sc;
  cx.s2 = cx.q[1] * cx.q[1] + cx.q[2] * cx.q[2] + cx.q[3] * cx.q[3];
  cx.s = sqrtf(cx.s2);
  float c = cx.q[0];
  cx.two_theta = 2.f * (c < 0.f ? atan2f(-cx.s, -c) : atan2f(cx.s, c));
  cx.k = cx.s2 > 0.f ? cx.two_theta / cx.s : 2.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float v = cx.q[i + 1] * cx.k;
    aa[i] = (v != v) ? 0.f : v;
  }
}
HD void mat2aa_fwd(const float* R, float* aa) {
  Mat2AACtx cx;
  mat2aa_core(R, aa, cx);
}
// dR += J^T daa   (recomputes the forward)
HD void mat2aa_bwd(const float* R, const float* daa_in, float* dR) {
  Mat2AACtx cx;
  float aa[3];
  mat2aa_core(R, aa, cx);
  float daa[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float v = cx.q[i + 1] * cx.k;
    daa[i] = (v != v) ? 0.f : daa_in[i];
  }
  float gq[4] = {0.f, daa[0] * cx.k, daa[1] * cx.k, daa[2] * cx.k};
  if (cx.s2 > 0.f) {
    float gk = daa[0] * cx.q[1] + daa[1] * cx.q[2] + daa[2] * cx.q[3];
    float g_tt = gk / cx.s;
    float g_s = -gk * cx.two_theta / cx.s2;
    float c = cx.q[0], n = cx.s2 + c * c;
    g_s += g_tt * 2.f * c / n;
    gq[0] = g_tt * (-2.f * cx.s / n);
    float g_s2 = g_s / (2.f * cx.s);
    gq[1] += 2.f * cx.q[1] * g_s2;
    gq[2] += 2.f * cx.q[2] * g_s2;
    gq[3] += 2.f * cx.q[3] * g_s2;
  }
  float sc = 0.5f / sqrtf(cx.t);
  float g_t = -(gq[0] * cx.q[0] + gq[1] * cx.q[1] + gq[2] * cx.q[2] + gq[3] * cx.q[3]) / (2.f * cx.t);
  float g[4] = {gq[0] * sc, gq[1] * sc, gq[2] * sc, gq[3] * sc};
  // scatter to m (transposed indexing): gm_ij is the grad wrt m[i][j] = R[j][i]
  float g00 = 0, g11 = 0, g22 = 0, g01 = 0, g10 = 0, g02 = 0, g20 = 0, g12 = 0, g21 = 0;
  switch (cx.branch) {
    case 0:
      g_t += g[1];
      g12 += g[0]; g21 -= g[0]; g01 += g[2]; g10 += g[2]; g20 += g[3]; g02 += g[3];
      g00 += g_t; g11 -= g_t; g22 -= g_t; break;
    case 1:
      g_t += g[2];
      g20 += g[0]; g02 -= g[0]; g01 += g[1]; g10 += g[1]; g12 += g[3]; g21 += g[3];
      g00 -= g_t; g11 += g_t; g22 -= g_t; break;
    case 2:
      g_t += g[3];
      g01 += g[0]; g10 -= g[0]; g20 += g[1]; g02 += g[1]; g12 += g[2]; g21 += g[2];
      g00 -= g_t; g11 -= g_t; g22 += g_t; break;
    default:
      g_t += g[0];
      g12 += g[1]; g21 -= g[1]; g20 += g[2]; g02 -= g[2]; g01 += g[3]; g10 -= g[3];
      g00 += g_t; g11 += g_t; g22 += g_t; break;
  }
  dR[0] += g00; dR[4] += g11; dR[8] += g22;
  dR[3] += g01; dR[1] += g10; dR[6] += g02; dR[2] += g20; dR[7] += g12; dR[5] += g21;
}

// ------------------------------------------------------------------------------------------------
// yaw that aligns the body-right axis (-R[:,0]) with +x
// ------------------------------------------------------------------------------------------------
HD void w2a_aa(const float* R0, float* aa) {
  const float e = 1e-6f;
  float rx = -R0[0], ry = -R0[3];
  float nxy = sqrtf(rx * rx + ry * ry);
  float xp = rx / (nxy + e);
  float xc = fminf(fmaxf(xp, -1.f), 1.f);
  float ang = acosf(xc);
  float az = -ry;                               // cross((rx,ry,0),(1,0,0)) = (0,0,-ry)
  float u = az / (fabsf(az) + e);
  aa[0] = 0.f; aa[1] = 0.f; aa[2] = u * ang;
}
HD void w2a_fwd(const float* R0, float* Ra) {
  float aa[3];
  w2a_aa(R0, aa);
  rodrigues_fwd(aa, Ra);
}
// dR0 += J^T dRa
HD void w2a_bwd(const float* R0, const float* dRa, float* dR0) {
  const float e = 1e-6f;
  float aa[3];
  w2a_aa(R0, aa);
  float daa[3] = {0.f, 0.f, 0.f};
  rodrigues_bwd(aa, dRa, daa);
  float rx = -R0[0], ry = -R0[3];
  float nxy = sqrtf(rx * rx + ry * ry);
  float xp = rx / (nxy + e);
  float xc = fminf(fmaxf(xp, -1.f), 1.f);
  float ang = acosf(xc);
  float az = -ry, naz = fabsf(az);
  float u = az / (naz + e);
  float g_ang = daa[2] * u, g_u = daa[2] * ang;
  // the x / y components of the axis are identically 0 * (.) : no gradient
  float sgn = az > 0.f ? 1.f : (az < 0.f ? -1.f : 0.f);
  float g_az = g_u * (1.f / (naz + e) - az * sgn / ((naz + e) * (naz + e)));
  float g_ry = -g_az, g_rx = 0.f;
  float g_xp = 0.f;
  if (xp >= -1.f && xp <= 1.f) g_xp = -g_ang / sqrtf(1.f - xc * xc);
  g_rx += g_xp / (nxy + e);
  float g_n = -g_xp * rx / ((nxy + e) * (nxy + e));
  if (nxy > 0.f) { g_rx += g_n * rx / nxy; g_ry += g_n * ry / nxy; }
  dR0[0] -= g_rx;
  dR0[3] -= g_ry;
}

}  // namespace hb
