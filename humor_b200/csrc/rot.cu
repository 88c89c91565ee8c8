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
extern "C" const char* humor_b200_version(void) { return "humor_b200 0.1 (sm_100a)"; }
#endif  // HB_HOST_SHIM
