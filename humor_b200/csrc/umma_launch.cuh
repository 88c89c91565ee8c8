#pragma once
#include "gemm.cuh"
#include "chain_args.cuh"
namespace hb {
// C[M,N] = (A_hi+A_lo)[M,K] * (B_hi+B_lo)[N,K]^T on tcgen05 (3xTF32).  Operand planes are fp32 arrays with the
// given leading dimensions (multiples of 4); K a multiple of 32.  Outputs: C (exact value) and/or the hi/lo planes.
cudaError_t launch_umma_gemm3(const float* A_hi, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb,
                              int M, int N, int K, float* C, float* C_hi, float* C_lo, int ldc, int epi, const GemmEpi& ep,
                              cudaStream_t st);
// same with the tile width chosen by the caller (bn = 64 or 128)
cudaError_t launch_umma_gemm3_bn(const float* A_hi, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb,
                                 int M, int N, int K, float* C, float* C_hi, float* C_lo, int ldc, int epi, const GemmEpi& ep,
                                 int bn, cudaStream_t st);
// hi = top 11 mantissa bits of x, lo = x - hi (exact); n elements
cudaError_t launch_split_hilo(const float* x, float* hi, float* lo, size_t n, cudaStream_t st);
bool umma_available();
// persistent decoder chain (chain_persist.cuh): S steps of one direction in one launch; zeroes a.flags on the stream first
cudaError_t launch_chain(const ChainLaunch& a, cudaStream_t st);
int chain_max_clusters();
void chain_set_debug(long long* buf, size_t bytes);
// fp16 hi/lo operand planes (umma_gemm16.cuh): 4 bytes per operand element instead of 8; planes [rows][ld] halves
// outputs: C (fp32, nullable) and/or the fp16 hi/lo planes of the result (ld16 halves per row, nullable); epi = EPI_BIAS | EPI_GN_RELU
cudaError_t launch_umma_gemm16(const void* A_h, const void* A_l, int lda, const void* B_h, const void* B_l, int ldb, int M, int N, int K,
                               float* C, int ldc, void* C16_h, void* C16_l, int ld16, int epi, const GemmEpi& ep, cudaStream_t st);
cudaError_t launch_split16(const float* x, void* h, void* l, size_t n, cudaStream_t st);
// dense LBS forward, skin form 3: blend GEMM + lane = frame group skinning in one persistent kernel (lbs_fuseg.cuh)
struct LbsFusegArgs {
  int N;                   // frames
  int num_verts;
  int num_groups;
  int nrt, nct;            // row tiles (128 frames), column tiles (64 vertices); set by the launcher
  int nkb16;               // > 0 (blend form 5): this many 64-wide fp16 k-blocks (kind::f16) on hi + lo planes: h.h + l.h + h.l
  int ft_rec_stride;       // bytes per column tile of ft_rec
  int dbg;                 // measurement switches (HB_LBS_FUSEG_DBG; results are WRONG with any of them): 1 no output stores, 2 no skinning
                           // at all (the epilogue only hands the TMEM buffer back), 4 operand loads without the L2 hint
  const int* ft_tab;       // [nct][FG_TAB]
  const unsigned char* ft_rec;   // [nct][ft_rec_stride] skinning records (HbLbsModel.ft_rec)
  const float* A;          // [N][52][12] skinning transforms OF THIS PASS: rotation part times the accumulator scale (2^-10 when
                           // the blend planes are pre-scaled for the fp16 range), translation column + trans unless `trans` is set
  const float* trans;      // [N][3] added to every vertex, or nullptr: already inside A
  // shaped template per SEQUENCE, or nullptr.  HuMoR fits one shape per sub-sequence (frames_per_beta = T), so template + shape
  // blend is a property of the sequence, not of the frame: with vs set, the GEMM carries the 189 pose columns only (K = 192: three
  // k-blocks instead of four, a quarter of the operand bytes the kernel is bound by) and the epilogue adds row frame / fpb of
  // vs [sequences][vs_ld] = (v_template + shapedirs . betas) / accumulator scale, staged per tile by the producer (fpb >= 32)
  const float* vs;
  int vs_ld;               // floats per row of vs (multiple of 4)
  int fpb;                 // frames per sequence
  float* out;              // [N][num_verts][3]
};
// bt_* = blend_t hi/lo planes [b_rows][K] (ldb floats per row) WITH the template in column 205 (HB_LBS_PLANES_TEMPLATE; the
// feature planes carry 1 there); the caller fills every field of `a` except nrt / nct.
// a.nkb16 > 0: feat16 [N][ld16] / bt16 [b_rows][ld16] fp16 hi planes (ld16 halves per row, >= 64 * nkb16), feat16l / bt16l the
// lo planes (same shapes); K (columns of the tf32 planes) is then 0
cudaError_t launch_lbs_fuseg(const float* feat_hi, const float* feat_lo, int ldf, const float* bt_hi, const float* bt_lo, int ldb,
                             int b_rows, int K, const void* feat16, const void* bt16, const void* feat16l, const void* bt16l, int ld16,
                             LbsFusegArgs a, cudaStream_t st);
// CTAs of the next lbs_fuseg launches (0: one per SM); > SMs = shorter chunks that the scheduler slots next to other streams' kernels
void lbs_set_fuseg_ctas(int n);
// fp16 plane(s) of the feature columns [c0, c0 + 64 * nkb16) of feat[N][ldf] (columns >= ncols read as zero): out = fp16(x),
// out_lo (nullable) = fp16(x - out); column `one_col` (>= 0) is written as 1 whatever feat holds there
cudaError_t launch_feat_f16(const float* feat, int ldf, int ncols, int N, int c0, int nkb16, void* out, void* out_lo, int one_col,
                            cudaStream_t st);
}  // namespace hb
