// Test infrastructure: the chamfer KERNELS of humor_b200/csrc/chamfer.cu executed on the CPU through the SIMT shim
// (tests/host/shim/cuda_runtime.h) — same source, same grid/block shapes as humor_chamfer_fwd / humor_chamfer_bwd.
//   g++ -O1 -std=c++20 -ffp-contract=off -pthread -shared -fPIC -Itests/host/shim -DHB_HOST_SHIM chamfer_host.cpp
#include "../../humor_b200/csrc/chamfer.cu"
using namespace hb;
extern "C" {
void h_chamfer_nn(int b, int nq, const float* Q, int np, const float* P, float* dist, int* idx) {
  if (nq == 0 || b == 0) return;
  if (np == 0) {
    const size_t count = (size_t)b * nq;
    shim::launch(dim3((unsigned)((count + 255) / 256)), dim3(256), [&] { chamfer_fill_zero_kernel(count, dist, idx); });
    return;
  }
  shim::launch(dim3(cdiv(nq, CH_THREADS * CH_QPT), b), dim3(CH_THREADS), [&] { chamfer_nn_kernel(nq, Q, np, P, dist, idx); });
}
void h_chamfer_bwd(int b, int n, const float* xyz1, int m, const float* xyz2, const float* gd1, const int* idx1,
                   const float* gd2, const int* idx2, float* g1, float* g2) {
  if (b == 0 || (n == 0 && m == 0)) return;
  const bool both = n > 0 && m > 0;
  shim::launch(dim3(b), dim3(CH_THREADS), [&] {
    chamfer_bwd_kernel(n, xyz1, m, xyz2, both ? gd1 : nullptr, idx1, both ? gd2 : nullptr, idx2, n > 0 ? g1 : nullptr,
                       m > 0 ? g2 : nullptr);
  });
}
}
