// Test infrastructure: the tcgen05 kernels of the dense LBS forward executed on the CPU: SIMT shim (threads, barriers) +
// functional emulation of mbarriers / TMA / UMMA / TMEM (tests/host/shim/tc_emul.h).
//   g++ -O2 -std=c++20 -pthread -shared -fPIC -Itests/host/shim -DHB_HOST_SHIM tc_host.cpp
#include "../../humor_b200/csrc/lbs_fuseg.cuh"
#include "../../humor_b200/csrc/umma_gemm16.cuh"
using namespace hb;
extern "C" {
// skin form 3 (lbs_fuseg.cuh): blend + lane = frame group skinning, `grid` persistent CTAs; transforms via un-swizzled TMA boxes
long long h_lbs_fuseg(const float* feat_hi, const float* feat_lo, int ldf, const float* bt_hi, const float* bt_lo, int ldb, int b_rows,
                      int K, int N, int num_verts, int num_groups, const int* ft_tab, const unsigned char* ft_rec, int ft_rec_stride,
                      const float* A, const float* trans, float* out, int grid, long long* tma_count, const void* feat16,
                      const void* bt16, int ld16, int nkb16, const void* feat16l, const void* bt16l, const float* vs, int vs_ld, int fpb) {
  CUtensorMap a_hi{feat_hi, (unsigned long long)N, (unsigned long long)K, (unsigned long long)ldf, 32, UM_BM, 0};
  CUtensorMap a_lo{feat_lo, (unsigned long long)N, (unsigned long long)K, (unsigned long long)ldf, 32, UM_BM, 0};
  CUtensorMap b_hi{bt_hi, (unsigned long long)b_rows, (unsigned long long)K, (unsigned long long)ldb, 32, FG_BN, 0};
  CUtensorMap b_lo{bt_lo, (unsigned long long)b_rows, (unsigned long long)K, (unsigned long long)ldb, 32, FG_BN, 0};
  CUtensorMap tt{A, (unsigned long long)N, 624ull, 624ull, 12, UM_BM, 1};
  CUtensorMap a16{static_cast<const float*>(feat16), (unsigned long long)N, 64ull * nkb16, (unsigned long long)ld16, 64, UM_BM, 0, 1};
  CUtensorMap b16{static_cast<const float*>(bt16), (unsigned long long)b_rows, 64ull * nkb16, (unsigned long long)ld16, 64, FG_BN, 0, 1};
  CUtensorMap a16l{static_cast<const float*>(feat16l), (unsigned long long)N, 64ull * nkb16, (unsigned long long)ld16, 64, UM_BM, 0, 1};
  CUtensorMap b16l{static_cast<const float*>(bt16l), (unsigned long long)b_rows, 64ull * nkb16, (unsigned long long)ld16, 64, FG_BN, 0, 1};
  LbsFusegArgs a;
  a.nkb16 = nkb16; a.dbg = 0;
  a.N = N; a.num_verts = num_verts; a.num_groups = num_groups; a.nrt = cdiv(N, UM_BM); a.nct = cdiv(num_groups, FG_GPT);
  a.ft_tab = ft_tab; a.ft_rec = ft_rec; a.ft_rec_stride = ft_rec_stride;
  a.A = A; a.trans = trans; a.out = out; a.vs = vs; a.vs_ld = vs_ld; a.fpb = fpb;
  tcemu::reset();
  shim::launch(dim3(grid), dim3(FG_THREADS), [&] { lbs_fuseg_kernel(a_hi, a_lo, b_hi, b_lo, tt, a16, b16, a16l, b16l, K, a); });
  if (tma_count) *tma_count = tcemu::g_tma_count;
  return tcemu::g_mma_count;
}
// fp16 hi/lo GEMM (umma_gemm16.cuh) with its epilogues: planes [rows][ld] halves; BN = 64, split-K over `ks` (1 or 4) CTAs
long long h_umma_gemm16(const unsigned short* A_h, const unsigned short* A_l, int lda, const unsigned short* B_h, const unsigned short* B_l,
                        int ldb, int M, int N, int K, float* C, int ldc, unsigned short* C16_h, unsigned short* C16_l, int ld16, int epi,
                        const float* bias, const float* gamma, const float* beta, float* xhat, int ldxh, float* rstd, int gsize, int ks,
                        int bn) {
  auto map = [](const unsigned short* p, int rows, int K, int ld, unsigned box_rows) {
    return CUtensorMap{reinterpret_cast<const float*>(p), (unsigned long long)rows, (unsigned long long)K, (unsigned long long)ld, 64, box_rows, 0, 1};
  };
  CUtensorMap a_h = map(A_h, M, K, lda, UM_BM), a_l = map(A_l, M, K, lda, UM_BM), b_h = map(B_h, N, K, ldb, (unsigned)bn),
              b_l = map(B_l, N, K, ldb, (unsigned)bn);
  GemmEpi ep;
  ep.bias = bias; ep.gamma = gamma; ep.beta = beta; ep.xhat = xhat; ep.rstd = rstd; ep.ldxh = ldxh; ep.Cch = N; ep.gsize = gsize;
  tcemu::reset();
  const dim3 grid(cdiv(N, bn) * ks, cdiv(M, UM_BM));
  if (bn == 128 && epi == EPI_BIAS) shim::launch_cluster(grid, dim3(192), 1, [&] { umma_gemm16_kernel<128, EPI_BIAS, 1>(a_h, a_l, b_h, b_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep); });
  else if (bn == 128) shim::launch_cluster(grid, dim3(192), 1, [&] { umma_gemm16_kernel<128, EPI_GN_RELU, 1>(a_h, a_l, b_h, b_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep); });
  else if (epi == EPI_BIAS && ks == 1) shim::launch_cluster(grid, dim3(192), 1, [&] { umma_gemm16_kernel<64, EPI_BIAS, 1>(a_h, a_l, b_h, b_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep); });
  else if (epi == EPI_BIAS) shim::launch_cluster(grid, dim3(192), 4, [&] { umma_gemm16_kernel<64, EPI_BIAS, 4>(a_h, a_l, b_h, b_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep); });
  else if (ks == 1) shim::launch_cluster(grid, dim3(192), 1, [&] { umma_gemm16_kernel<64, EPI_GN_RELU, 1>(a_h, a_l, b_h, b_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep); });
  else shim::launch_cluster(grid, dim3(192), 4, [&] { umma_gemm16_kernel<64, EPI_GN_RELU, 4>(a_h, a_l, b_h, b_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep); });
  return tcemu::g_mma_count;
}

// persistent 128x128-tile 3xTF32 GEMM with two epilogue groups (umma_gemm3p_kernel): `grid` CTAs walk the tiles round-robin
long long h_umma_gemm3p(const float* A_hi, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb, int M, int N, int K,
                        float* C, float* C_hi, float* C_lo, int ldc, int epi, const float* bias, const float* gamma, const float* beta,
                        float* xhat, int ldxh, float* rstd, int gsize, int cch, int grid) {
  CUtensorMap a_hi{A_hi, (unsigned long long)M, (unsigned long long)K, (unsigned long long)lda, 32, UM_BM};
  CUtensorMap a_lo{A_lo, (unsigned long long)M, (unsigned long long)K, (unsigned long long)lda, 32, UM_BM};
  CUtensorMap b_hi{B_hi, (unsigned long long)N, (unsigned long long)K, (unsigned long long)ldb, 32, 128};
  CUtensorMap b_lo{B_lo, (unsigned long long)N, (unsigned long long)K, (unsigned long long)ldb, 32, 128};
  GemmEpi ep;
  ep.bias = bias; ep.gamma = gamma; ep.beta = beta; ep.xhat = xhat; ep.rstd = rstd; ep.ldxh = ldxh; ep.Cch = cch; ep.gsize = gsize;
  tcemu::reset();
  if (epi == EPI_BIAS) shim::launch(dim3(grid), dim3(UMP_THREADS), [&] { umma_gemm3p_kernel<EPI_BIAS>(a_hi, a_lo, b_hi, b_lo, M, N, K, C, C_hi, C_lo, ldc, ep); });
  else if (epi == EPI_GN_RELU) shim::launch(dim3(grid), dim3(UMP_THREADS), [&] { umma_gemm3p_kernel<EPI_GN_RELU>(a_hi, a_lo, b_hi, b_lo, M, N, K, C, C_hi, C_lo, ldc, ep); });
  else shim::launch(dim3(grid), dim3(UMP_THREADS), [&] { umma_gemm3p_kernel<EPI_GN_RELU_BWD>(a_hi, a_lo, b_hi, b_lo, M, N, K, C, C_hi, C_lo, ldc, ep); });
  return tcemu::g_mma_count;
}
}
