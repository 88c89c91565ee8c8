// Host side of the tcgen05 GEMM: TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry
// point, so the library does not link libcuda) and launches.
#include <cstdlib>
#ifndef HB_HOST_SHIM
#include <cuda_fp16.h>
#endif
#include "umma_gemm.cuh"
#include "umma_launch.cuh"
#include "lbs_fuseg.cuh"
#include "umma_gemm16.cuh"
#include "chain_persist.cuh"
#include "../../include/humor_b200.h"

namespace hb {

#ifndef HB_HOST_SHIM   // host side (TMA descriptors, launches): device builds only
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static const bool g_splitk = (getenv("HB_NO_SPLITK") == nullptr);
static const bool g_pdl = (getenv("HB_NO_PDL") == nullptr);
static const bool g_persist = (getenv("HB_GEMM_ONE_TILE") == nullptr);   // A/B: one 128x128 tile per CTA (round 1) instead of persistent CTAs
static int g_encode_state = 0;      // 0 unknown, 1 ok, -1 unavailable

static bool load_encode() {
  if (g_encode_state == 0) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn) { g_encode = (EncodeTiledFn)fn; g_encode_state = 1; }
    else { g_encode_state = -1; (void)cudaGetLastError(); }
  }
  return g_encode_state == 1;
}
bool umma_available() { return load_encode(); }

// rows x K fp32 matrix, row stride ld floats; box = {32 floats, box_rows}, 128-byte swizzle, OOB rows read as zero
static bool make_map(CUtensorMap* m, const float* base, int rows, int K, int ld, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)UM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// rows x cols fp32 matrix, row stride ld floats; un-swizzled box = {box_cols floats, box_rows} (box_cols * 4 a multiple of 16):
// shared memory receives the box rows back to back
static bool make_map_plain(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// rows x K fp16 matrix, row stride ld halves; box = {64 halves = 128 B, box_rows}, 128-byte swizzle, OOB reads as zero
static bool make_map_f16(CUtensorMap* m, const void* base, int rows, int K, int ld, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int EPI, int KS>
static cudaError_t launch_t(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                            int M, int N, int K, float* C, float* C_hi, float* C_lo, int ldc, const GemmEpi& ep, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm3_kernel<BN, EPI, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaSmem<BN>::TOTAL);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(N, BN) * KS, cdiv(M, UM_BM));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = UmmaSmem<BN>::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // PDL: see griddepcontrol in the kernel
  at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = KS; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = KS > 1 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, umma_gemm3_kernel<BN, EPI, KS>, a_hi, a_lo, b_hi, b_lo, M, N, K, C, C_hi, C_lo, ldc, ep);
}

// persistent 128x128 tiles with two epilogue groups (umma_gemm3p_kernel): the batched products
template <int EPI>
static cudaError_t launch_p(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                            int M, int N, int K, float* C, float* C_hi, float* C_lo, int ldc, const GemmEpi& ep, cudaStream_t st) {
  static int sms = 0, want = 0;
  if (!sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_gemm3p_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, UMP_SMEM);
    if (e != cudaSuccess) { sms = 0; return e; }
    const char* w = getenv("HB_GEMM_CTAS");                    // tests: fewer CTAs than tiles on small problems
    want = w ? atoi(w) : 0;
  }
  const int ntiles = cdiv(N, 128) * cdiv(M, UM_BM);
  int grid = ntiles < sms ? ntiles : sms;
  if (want > 0 && want < grid) grid = want;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(UMP_THREADS);
  cfg.dynamicSmemBytes = UMP_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, umma_gemm3p_kernel<EPI>, a_hi, a_lo, b_hi, b_lo, M, N, K, C, C_hi, C_lo, ldc, ep);
}

cudaError_t launch_umma_gemm3(const float* A_hi, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb,
                              int M, int N, int K, float* C, float* C_hi, float* C_lo, int ldc, int epi, const GemmEpi& ep,
                              cudaStream_t st) {
  // few row tiles (the sequential decoder steps, M = sub-sequences per GPU): 64-column tiles double the CTA count
  return launch_umma_gemm3_bn(A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, C, C_hi, C_lo, ldc, epi, ep, (M <= 1024) ? 64 : 128, st);
}

cudaError_t launch_umma_gemm3_bn(const float* A_hi, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb,
                                 int M, int N, int K, float* C, float* C_hi, float* C_lo, int ldc, int epi, const GemmEpi& ep,
                                 int bn, cudaStream_t st) {
  if (!load_encode()) return cudaErrorNotSupported;
  if (bn != 64 && bn != 128) return cudaErrorInvalidValue;
  if (K % UM_BK || lda % 4 || ldb % 4 || ldc % 4) return cudaErrorInvalidValue;
  if ((C_hi == nullptr) != (C_lo == nullptr)) return cudaErrorInvalidValue;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  if (!make_map(&ta_hi, A_hi, M, K, lda, UM_BM) || !make_map(&ta_lo, A_lo, M, K, lda, UM_BM) ||
      !make_map(&tb_hi, B_hi, N, K, ldb, bn) || !make_map(&tb_lo, B_lo, N, K, ldb, bn))
    return cudaErrorInvalidValue;
  // split-K over a 4-CTA cluster when there are few output tiles and K is long enough to share
  const bool splitk = bn == 64 && g_splitk && cdiv(N, 64) * cdiv(M, UM_BM) <= 64 && K >= 8 * UM_BK;
#define HB_UMMA_CASE(E)                                                                                              \
  case E:                                                                                                            \
    if (splitk) return launch_t<64, E, 4>(ta_hi, ta_lo, tb_hi, tb_lo, M, N, K, C, C_hi, C_lo, ldc, ep, st);          \
    if (bn == 64) return launch_t<64, E, 1>(ta_hi, ta_lo, tb_hi, tb_lo, M, N, K, C, C_hi, C_lo, ldc, ep, st);         \
    return (g_persist && (E == EPI_BIAS || ep.gsize == 64)) ? launch_p<E>(ta_hi, ta_lo, tb_hi, tb_lo, M, N, K, C, C_hi, C_lo, ldc, ep, st)                  \
                     : launch_t<128, E, 1>(ta_hi, ta_lo, tb_hi, tb_lo, M, N, K, C, C_hi, C_lo, ldc, ep, st);
  switch (epi) {
    HB_UMMA_CASE(EPI_BIAS)
    HB_UMMA_CASE(EPI_GN_RELU)
    HB_UMMA_CASE(EPI_GN_RELU_BWD)
  }
#undef HB_UMMA_CASE
  return cudaErrorInvalidValue;
}

static int g_fuseg_ctas = 0;
void lbs_set_fuseg_ctas(int n) { g_fuseg_ctas = n > 0 ? n : 0; }
// skin form 3: blend + group skinning in one persistent kernel (lbs_fuseg.cuh).  `a` arrives with the model tables, A, trans,
// out, N, num_verts, num_groups filled in; tile counts are set here.
cudaError_t launch_lbs_fuseg(const float* feat_hi, const float* feat_lo, int ldf, const float* bt_hi, const float* bt_lo, int ldb,
                             int b_rows, int K, const void* feat16, const void* bt16, const void* feat16l, const void* bt16l, int ld16,
                             LbsFusegArgs a, cudaStream_t st) {
  if (!load_encode()) return cudaErrorNotSupported;
  if (K % UM_BK || ldf % 4 || ldb % 4 || a.N <= 0 || a.num_groups <= 0 || (a.num_verts & 1) ||
      a.num_groups != cdiv(a.num_verts, FG_G) || !a.ft_tab || !a.ft_rec || a.ft_rec_stride < FG_REC_HEAD + FG_REC_ENTRY ||
      a.ft_rec_stride > FG_REC_MAX || a.ft_rec_stride % 16 || (reinterpret_cast<uintptr_t>(a.ft_rec) & 15u) || !a.A || !a.out)
    return cudaErrorInvalidValue;
  if (a.nkb16 < 0 || (a.nkb16 > 0 && (!feat16 || !bt16 || !feat16l || !bt16l || ld16 % 8 || ld16 < 64 * a.nkb16))) return cudaErrorInvalidValue;
  if ((K == 0) == (a.nkb16 <= 0)) return cudaErrorInvalidValue;          // either the tf32 planes or the fp16 planes
  if (a.vs && (a.fpb < 32 || a.vs_ld % 4 || (reinterpret_cast<uintptr_t>(a.vs) & 15u))) return cudaErrorInvalidValue;
  static int sms = 0, want = 0;
  if (!sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lbs_fuseg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FG_SMEM);
    if (e != cudaSuccess) { sms = 0; return e; }
    // persistent CTAs hold an SM (223 KB of shared memory, all of TMEM) for the whole pass: HB_LBS_FUSEG_CTAS leaves SMs to
    // kernels of other streams (the latency-bound decoder chain runs next to the dense pass)
    const char* w = getenv("HB_LBS_FUSEG_CTAS");
    want = w ? atoi(w) : 0;
  }
  static const int dbg = getenv("HB_LBS_FUSEG_DBG") ? atoi(getenv("HB_LBS_FUSEG_DBG")) : 0;
  a.dbg = dbg;
  a.nrt = cdiv(a.N, UM_BM);
  a.nct = cdiv(a.num_groups, FG_GPT);
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tt, ta16, tb16, ta16l, tb16l;
  if (!make_map_plain(&tt, a.A, a.N, 624, 624, 12, UM_BM)) return cudaErrorInvalidValue;
  if (K > 0 && (!make_map(&ta_hi, feat_hi, a.N, K, ldf, UM_BM) || !make_map(&ta_lo, feat_lo, a.N, K, ldf, UM_BM) ||
                !make_map(&tb_hi, bt_hi, b_rows, K, ldb, FG_BN) || !make_map(&tb_lo, bt_lo, b_rows, K, ldb, FG_BN)))
    return cudaErrorInvalidValue;
  if (a.nkb16 > 0 && (!make_map_f16(&ta16, feat16, a.N, 64 * a.nkb16, ld16, UM_BM) || !make_map_f16(&tb16, bt16, b_rows, 64 * a.nkb16, ld16, FG_BN) ||
                      !make_map_f16(&ta16l, feat16l, a.N, 64 * a.nkb16, ld16, UM_BM) || !make_map_f16(&tb16l, bt16l, b_rows, 64 * a.nkb16, ld16, FG_BN)))
    return cudaErrorInvalidValue;
  // descriptors of the operand kind this launch does not use are never dereferenced: any valid one stands in
  if (K == 0) { ta_hi = ta_lo = ta16; tb_hi = tb_lo = tb16; }
  else { ta16 = ta16l = ta_hi; tb16 = tb16l = tb_hi; }
  const int ntiles = a.nrt * a.nct;
  int grid = ntiles < sms ? ntiles : sms;
  // HB_LBS_FUSEG_CTAS: fewer CTAs leave SMs to kernels of other streams; MORE CTAs than SMs (each walks a shorter chunk of the tile
  // list and retires) let the hardware scheduler slot the pass into whatever SMs the main stream's kernels leave idle
  if (want > 0) grid = want < ntiles ? want : ntiles;
  if (g_fuseg_ctas > 0) grid = g_fuseg_ctas < ntiles ? g_fuseg_ctas : ntiles;      // per-call override (lbs_set_fuseg_ctas)
  lbs_fuseg_kernel<<<grid, FG_THREADS, FG_SMEM, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, tt, ta16, tb16, ta16l, tb16l, K, a);
  return cudaGetLastError();
}

// fp16 hi/lo GEMM (umma_gemm16.cuh): planes [rows][ld] halves, ld % 8 == 0, K % 64 == 0
template <int BN, int EPI, int KS>
static cudaError_t launch16_t(const CUtensorMap& a_h, const CUtensorMap& a_l, const CUtensorMap& b_h, const CUtensorMap& b_l, int M, int N,
                              int K, float* C, int ldc, unsigned short* C16_h, unsigned short* C16_l, int ld16, const GemmEpi& ep,
                              cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm16_kernel<BN, EPI, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaSmem<BN>::TOTAL);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cdiv(N, BN) * KS, cdiv(M, UM_BM));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = UmmaSmem<BN>::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // PDL: see griddepcontrol in the kernel
  at[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = KS; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = KS > 1 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, umma_gemm16_kernel<BN, EPI, KS>, a_h, a_l, b_h, b_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep);
}

cudaError_t launch_umma_gemm16(const void* A_h, const void* A_l, int lda, const void* B_h, const void* B_l, int ldb, int M, int N, int K,
                               float* C, int ldc, void* C16_h_, void* C16_l_, int ld16, int epi, const GemmEpi& ep, cudaStream_t st) {
  if (!load_encode()) return cudaErrorNotSupported;
  if (K % U16_BK || lda % 8 || ldb % 8 || lda < K || ldb < K || (C && ldc % 4) || (!C && !C16_h_)) return cudaErrorInvalidValue;
  if ((C16_h_ == nullptr) != (C16_l_ == nullptr) || (C16_h_ && ld16 % 4) || (epi != EPI_BIAS && epi != EPI_GN_RELU)) return cudaErrorInvalidValue;
  unsigned short* C16_h = static_cast<unsigned short*>(C16_h_);
  unsigned short* C16_l = static_cast<unsigned short*>(C16_l_);
  const int bn = (M <= 1024) ? 64 : 128;                     // few row tiles: narrow tiles + split-K, as launch_umma_gemm3
  CUtensorMap ta_h, ta_l, tb_h, tb_l;
  if (!make_map_f16(&ta_h, A_h, M, K, lda, UM_BM) || !make_map_f16(&ta_l, A_l, M, K, lda, UM_BM) ||
      !make_map_f16(&tb_h, B_h, N, K, ldb, bn) || !make_map_f16(&tb_l, B_l, N, K, ldb, bn))
    return cudaErrorInvalidValue;
  const bool splitk = bn == 64 && g_splitk && cdiv(N, 64) * cdiv(M, UM_BM) <= 64 && K >= 4 * U16_BK;
#define HB_U16_CASE(E)                                                                                                                  \
  case E:                                                                                                                               \
    if (splitk) return launch16_t<64, E, 4>(ta_h, ta_l, tb_h, tb_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep, st);                        \
    return bn == 64 ? launch16_t<64, E, 1>(ta_h, ta_l, tb_h, tb_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep, st)                          \
                    : launch16_t<128, E, 1>(ta_h, ta_l, tb_h, tb_l, M, N, K, C, ldc, C16_h, C16_l, ld16, ep, st);
  switch (epi) {
    HB_U16_CASE(EPI_BIAS)
    HB_U16_CASE(EPI_GN_RELU)
  }
#undef HB_U16_CASE
  return cudaErrorInvalidValue;
}

// persistent decoder chain (chain_persist.cuh): one launch per direction.  Clusters resident at once are bounded by what the
// device can co-schedule (cudaOccupancyMaxActiveClusters): the kernel's data-flow flags need every launched cluster running.
static int g_chain_clusters = 0;        // 0: not yet queried
static long long* g_chain_dbg = nullptr;   // humor_chain_debug: per-phase clock stamps of CTA 0 for tools/chain_timeline.py
static size_t g_chain_dbg_bytes = 0;
void chain_set_debug(long long* buf, size_t bytes) { g_chain_dbg = buf; g_chain_dbg_bytes = bytes; }
int chain_max_clusters() { return g_chain_clusters; }
cudaError_t launch_chain(const ChainLaunch& a, cudaStream_t st) {
  if (!load_encode()) return cudaErrorNotSupported;
  if (a.B <= 0 || a.B > CH_MAX_MT * UM_BM || a.S <= 0 || !a.flags) return cudaErrorInvalidValue;
  if (!g_chain_clusters) {              // once per process, on the first (un-captured) call
    cudaError_t e = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(32 * CH_CS); q.blockDim = dim3(CHAIN_THREADS); q.dynamicSmemBytes = CH_SMEM;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = CH_CS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, chain_kernel, &q);
    if (e != cudaSuccess) return e;
    const char* w = getenv("HB_CHAIN_CLUSTERS");
    int want = w ? atoi(w) : 32;
    if (want < 1) want = 1;
    g_chain_clusters = n < want ? n : want;
    if (g_chain_clusters < 1) return cudaErrorInvalidConfiguration;
  }
  static ChainParams p;                  // 64-byte aligned (CUtensorMap); filled per launch, copied by the launch itself
  for (int i = 0; i < CH_NMAPS; ++i) {
    const ChainPlane& pl = a.planes[i];
    if (!pl.hi) { if (i == 0) return cudaErrorInvalidValue; p.map_hi[i] = p.map_hi[0]; p.map_lo[i] = p.map_lo[0]; continue; }
    if (!pl.lo) return cudaErrorInvalidValue;
    if (pl.half) {                        // fp16 hi / scaled-lo planes: boxes of 64 halves
      if (pl.cols % 64 || pl.ld % 8) return cudaErrorInvalidValue;
      if (!make_map_f16(&p.map_hi[i], pl.hi, pl.rows, pl.cols, pl.ld, pl.box_rows) ||
          !make_map_f16(&p.map_lo[i], pl.lo, pl.rows, pl.cols, pl.ld, pl.box_rows))
        return cudaErrorInvalidValue;
      continue;
    }
    if (pl.cols % UM_BK || pl.ld % 4) return cudaErrorInvalidValue;
    if (!make_map(&p.map_hi[i], static_cast<const float*>(pl.hi), pl.rows, pl.cols, pl.ld, pl.box_rows) ||
        !make_map(&p.map_lo[i], static_cast<const float*>(pl.lo), pl.rows, pl.cols, pl.ld, pl.box_rows))
      return cudaErrorInvalidValue;
  }
  for (int i = 0; i < CH_NGEMM; ++i) {
    p.g[i] = a.g[i];
    if (a.g[i].ntn < 1 || a.g[i].ntn > CH_MAX_NT || a.g[i].nkb < 1 || (a.g[i].gsize != 64 && a.g[i].gsize != 32)) return cudaErrorInvalidValue;
  }
  p.glue = a.glue; p.flags = a.flags; p.B = a.B; p.S = a.S; p.dir = a.dir; p.f16 = (a.f16 && a.dir == 0) ? 1 : 0;
  p.dbg = (g_chain_dbg && g_chain_dbg_bytes >= (size_t)a.S * 5 * CH_DBG_EV * sizeof(long long)) ? g_chain_dbg : nullptr;
  cudaError_t e = cudaMemsetAsync(a.flags, 0, CH_FLAGS * sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g_chain_clusters * CH_CS);
  cfg.blockDim = dim3(CHAIN_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CH_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, chain_kernel, p);
}

cudaError_t launch_split16(const float* x, void* h, void* l, size_t n, cudaStream_t st) {
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split16_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, static_cast<unsigned short*>(h), static_cast<unsigned short*>(l), n);
  return cudaGetLastError();
}

#endif  // HB_HOST_SHIM

__global__ void split_hilo_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  }
}
// fp16 operand plane of the pose-feature columns (blend form 4): out[n][j] = fp16(feat[n][c0 + j]), j < 64 * nkb16, columns
// past ncols read as zero.  Round to nearest: the same 11-bit significand the tf32-rounded hi plane carries.
#ifdef HB_HOST_SHIM
static inline unsigned short f16_bits(float x) { return tcemu::f32_to_f16_bits(x); }
#else
__device__ __forceinline__ unsigned short f16_bits(float x) { return __half_as_ushort(__float2half_rn(x)); }
#endif
#ifdef HB_HOST_SHIM
static inline float f16_value(unsigned short b) { _Float16 h; std::memcpy(&h, &b, 2); return (float)h; }
#else
__device__ __forceinline__ float f16_value(unsigned short b) { return __half2float(__ushort_as_half(b)); }
#endif
__global__ void feat_f16_kernel(const float* __restrict__ feat, int ldf, int ncols, int N, int c0, int w, unsigned short* __restrict__ out,
                                unsigned short* __restrict__ out_lo, int one_col) {
  const size_t n = (size_t)N * w;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / w;
    const int j = (int)(i - r * w);
    const float x = c0 + j == one_col ? 1.f : (c0 + j < ncols ? feat[r * ldf + c0 + j] : 0.f);
    const unsigned short h = f16_bits(x);
    out[i] = h;
    if (out_lo) out_lo[i] = f16_bits(x - f16_value(h));        // unscaled lo plane of blend form 5
  }
}
#ifndef HB_HOST_SHIM
cudaError_t launch_feat_f16(const float* feat, int ldf, int ncols, int N, int c0, int nkb16, void* out, void* out_lo, int one_col,
                            cudaStream_t st) {
  if (!feat || !out || N <= 0 || nkb16 <= 0) return cudaErrorInvalidValue;
  const size_t n = (size_t)N * 64 * nkb16;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  feat_f16_kernel<<<(unsigned)blocks, 256, 0, st>>>(feat, ldf, ncols, N, c0, 64 * nkb16, static_cast<unsigned short*>(out),
                                                    static_cast<unsigned short*>(out_lo), one_col);
  return cudaGetLastError();
}
cudaError_t launch_split_hilo(const float* x, float* hi, float* lo, size_t n, cudaStream_t st) {
  if (n % 4) return cudaErrorInvalidValue;
  const size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  split_hilo_kernel<<<blocks, 256, 0, st>>>(x, hi, lo, n4);
  return cudaGetLastError();
}

#endif  // HB_HOST_SHIM
}  // namespace hb
#ifndef HB_HOST_SHIM
using namespace hb;

// Test / utility entry point: C = A * B^T (+bias) in fp32-level accuracy on the tensor cores.
// workspace: 2*(M*lda + N*ldb) floats for the operand planes.
extern "C" size_t humor_umma_gemm_workspace_bytes(int M, int N, int lda, int ldb) {
  return (size_t)2 * ((size_t)M * lda + (size_t)N * ldb) * sizeof(float);
}
extern "C" int humor_umma_gemm(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M,
                               int N, int K, float* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!A || !B || !C || !workspace || M <= 0 || N <= 0 || K <= 0) return HB_ERR_ARG;
  if (workspace_bytes < humor_umma_gemm_workspace_bytes(M, N, lda, ldb)) return HB_ERR_WORKSPACE;
  float* a_hi = workspace;
  float* a_lo = a_hi + (size_t)M * lda;
  float* b_hi = a_lo + (size_t)M * lda;
  float* b_lo = b_hi + (size_t)N * ldb;
  HB_CUDA(launch_split_hilo(A, a_hi, a_lo, (size_t)M * lda, st));
  HB_CUDA(launch_split_hilo(B, b_hi, b_lo, (size_t)N * ldb, st));
  GemmEpi ep;
  ep.bias = bias; ep.gamma = ep.beta = nullptr; ep.xhat = ep.rstd = nullptr; ep.ldxh = 0; ep.Cch = 0; ep.gsize = 64;
  HB_CUDA(launch_umma_gemm3(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, C, nullptr, nullptr, ldc, EPI_BIAS, ep, st));
  return HB_OK;
}

// Utility entry point: the same product from 4-byte operand elements (fp16 hi + scaled fp16 lo planes, umma_gemm16.cuh).
// lda / ldb: row strides of A / B in elements, multiples of 8; K a multiple of 64.
extern "C" size_t humor_umma_gemm16_workspace_bytes(int M, int N, int lda, int ldb) {
  return (size_t)2 * ((size_t)M * lda + (size_t)N * ldb) * sizeof(unsigned short);
}
extern "C" int humor_umma_gemm16(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M,
                                 int N, int K, float* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (!A || !B || !C || !workspace || M <= 0 || N <= 0 || K <= 0) return HB_ERR_ARG;
  if (workspace_bytes < humor_umma_gemm16_workspace_bytes(M, N, lda, ldb)) return HB_ERR_WORKSPACE;
  unsigned short* a_h = reinterpret_cast<unsigned short*>(workspace);
  unsigned short* a_l = a_h + (size_t)M * lda;
  unsigned short* b_h = a_l + (size_t)M * lda;
  unsigned short* b_l = b_h + (size_t)N * ldb;
  HB_CUDA(launch_split16(A, a_h, a_l, (size_t)M * lda, st));
  HB_CUDA(launch_split16(B, b_h, b_l, (size_t)N * ldb, st));
  GemmEpi ep;
  ep.bias = bias; ep.gamma = ep.beta = nullptr; ep.xhat = ep.rstd = nullptr; ep.ldxh = 0; ep.Cch = 0; ep.gsize = 64;
  HB_CUDA(launch_umma_gemm16(a_h, a_l, lda, b_h, b_l, ldb, M, N, K, C, ldc, nullptr, nullptr, 0, EPI_BIAS, ep, st));
  return HB_OK;
}
#endif  // HB_HOST_SHIM
