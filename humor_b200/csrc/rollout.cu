// HuMoR CVAE rollout (reference: models/humor_model.py:785-1017 roll_out, :1019-1059 sample_step,
// :407-418 prior, :445-498 decode) and its reverse pass, as a sequence of kernels on one stream.
//
// B200-first restructuring of the reference's Python loop:
//   * only the DECODER is on the sequential critical path (4 fused GEMM+GroupNorm+ReLU launches and
//     one glue launch per step).  The conditional PRIOR of step t depends only on the step's input
//     state, so all S steps are evaluated afterwards as ONE batched MLP over M = S*B rows.
//   * the same holds in reverse: the prior's input-gradients for all steps come from one batched
//     backward; the BPTT loop then only walks decoder + glue.
//   * activations live time-major ([S][B][ld]) so each step is a dense row block for the GEMMs;
//     cat(x, z) skip connections are free (z is written once into the padded tail columns).
#include <cstdlib>
#include "gemm.cuh"
#include "rollout_glue.cuh"
#include "glue_warp.cuh"
#include "umma_launch.cuh"
#include "umma_split16.cuh"
#include "../../include/humor_b200.h"

namespace hb {

constexpr int BP_LD = 224 + 512 + 1024 + 1024;   // reverse operand planes of the persistent chain: d raw | d pre3 | d pre2 | d pre1
constexpr int BP_C3 = 224, BP_C2 = 736, BP_C1 = 1760;
constexpr int X16_LD = 448;         // decoder input row (339 state + 48 z, XIN_LD = 416) padded to a multiple of 64 halves

struct Tape {
  float *xins, *raws, *Gs, *t2j;
  float *dxh1, *dxh2, *dxh3, *drs1, *drs2, *drs3;
  float *pxh1, *pxh2, *pxh3, *pxh4, *prs1, *prs2, *prs3, *prs4;
  float *h1, *h2, *h3;                 // decoder transients [B][1088],[B][1088],[B][576]
  float *pa, *pb;                      // prior ping-pong [S*B][1024]
  float *dh3, *dh2, *dh1, *da0, *draw, *dxres, *dnsum, *dG0, *dG1, *dt2j, *dpx;
  // hi/lo operand planes of the tensor-core path (x = hi + lo); h1/h2/h3, pa/pb, dh1/dh2/dh3 hold the hi plane
  float *xin_hi, *xin_lo, *pa_lo, *pb_lo, *dpo_hi, *dpo_lo, *h1_lo, *h2_lo, *h3_lo, *dh1_lo, *dh2_lo, *dh3_lo, *draw_hi, *draw_lo;
  // fp16 hi/lo operand planes of the forward pass (use_umma == 2): step inputs [S][B][448] (decoder K = 448, prior K = 384), decoder
  // transients [B][1088], [B][1088], [B][576], prior ping-pong [S*B][1024]; halves
  unsigned short *x16_h, *x16_l, *h1_16h, *h1_16l, *h2_16h, *h2_16l, *h3_16h, *h3_16l, *pa16_h, *pa16_l, *pb16_h, *pb16_l;
  // persistent decoder chain (chain_persist.cuh): data-flow flags, the reverse pass's per-step operand planes
  // [S*B][BP_LD] = d raw 224 | d pre3 512 | d pre2 1024 | d pre1 1024, and d z of every step [S*B][48]
  unsigned* chain_flags;
  float *bp_hi, *bp_lo, *dz_all;
  size_t total;
};

static Tape carve(float* base, int B, int S) {
  Tape t;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += align_up(n, 64); return p; };
  const size_t M = (size_t)S * B;
  t.xins = take((size_t)(S + 1) * B * XIN_LD);
  t.raws = take(M * RAW_LD);
  t.Gs = take((size_t)(S + 1) * B * 12);
  t.t2j = take((size_t)B * 4);
  t.dxh1 = take(M * 1024); t.dxh2 = take(M * 1024); t.dxh3 = take(M * 512);
  t.drs1 = take(M * 16); t.drs2 = take(M * 16); t.drs3 = take(M * 16);
  t.pxh1 = take(M * 1024); t.pxh2 = take(M * 1024); t.pxh3 = take(M * 1024); t.pxh4 = take(M * 1024);
  t.prs1 = take(M * 16); t.prs2 = take(M * 16); t.prs3 = take(M * 16); t.prs4 = take(M * 16);
  t.h1 = take((size_t)B * 1088); t.h2 = take((size_t)B * 1088); t.h3 = take((size_t)B * 576);
  t.pa = take(M * 1024); t.pb = take(M * 1024);
  t.dh3 = take((size_t)B * 576); t.dh2 = take((size_t)B * 1088); t.dh1 = take((size_t)B * 1088);
  t.da0 = take((size_t)B * XIN_LD); t.draw = take((size_t)B * RAW_LD);
  t.dxres = take((size_t)B * 340); t.dnsum = take((size_t)B * 340);
  t.dG0 = take((size_t)B * 12); t.dG1 = take((size_t)B * 12); t.dt2j = take((size_t)B * 4);
  t.dpx = take(M * 352);
  t.xin_hi = take((size_t)(S + 1) * B * XIN_LD); t.xin_lo = take((size_t)(S + 1) * B * XIN_LD);
  t.pa_lo = take(M * 1024); t.pb_lo = take(M * 1024);
  t.dpo_hi = take(M * 96); t.dpo_lo = take(M * 96);
  t.h1_lo = take((size_t)B * 1088); t.h2_lo = take((size_t)B * 1088); t.h3_lo = take((size_t)B * 576);
  t.dh1_lo = take((size_t)B * 1088); t.dh2_lo = take((size_t)B * 1088); t.dh3_lo = take((size_t)B * 576);
  t.draw_hi = take((size_t)B * RAW_LD); t.draw_lo = take((size_t)B * RAW_LD);
  auto take16 = [&](size_t halves) { return reinterpret_cast<unsigned short*>(take((halves + 1) / 2)); };
  t.x16_h = take16(M * X16_LD); t.x16_l = take16(M * X16_LD);
  t.h1_16h = take16((size_t)B * 1088); t.h1_16l = take16((size_t)B * 1088);
  t.h2_16h = take16((size_t)B * 1088); t.h2_16l = take16((size_t)B * 1088);
  t.h3_16h = take16((size_t)B * 576); t.h3_16l = take16((size_t)B * 576);
  t.pa16_h = take16(M * 1024); t.pa16_l = take16(M * 1024); t.pb16_h = take16(M * 1024); t.pb16_l = take16(M * 1024);
  t.chain_flags = reinterpret_cast<unsigned*>(take(CH_FLAGS));
  t.bp_hi = take(M * BP_LD); t.bp_lo = take(M * BP_LD); t.dz_all = take(M * 48);
  t.total = off;
  return t;
}

__device__ __forceinline__ float hi11(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// write v as (hi, lo) planes when a lo plane exists, else exactly
__device__ __forceinline__ void put_split(float* hi, float* lo, size_t i, float v) {
  if (lo) { const float h = hi11(v); hi[i] = h; lo[i] = v - h; } else { hi[i] = v; }
}
__device__ __forceinline__ float get_split(const float* hi, const float* lo, size_t i) { return lo ? hi[i] + lo[i] : hi[i]; }

// use_umma == 2: the decoder input of one step as fp16 hi/lo planes (x = h + l * 2^-11) - the state | z row written by the
// previous glue kernel (or rollout_init_kernel) - and z into the skip-connection columns of the three hidden-activation planes
__global__ void chain16_pack_kernel(int B, const float* __restrict__ xin, unsigned short* x_h, unsigned short* x_l, unsigned short* h1_h,
                                    unsigned short* h1_l, unsigned short* h2_h, unsigned short* h2_l, unsigned short* h3_h,
                                    unsigned short* h3_l) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const float* x = xin + (size_t)b * XIN_LD;
  for (int i = threadIdx.x; i < XIN_LD; i += blockDim.x) {
    unsigned short h, l;
    split16(x[i], h, l);
    x_h[(size_t)b * X16_LD + i] = h; x_l[(size_t)b * X16_LD + i] = l;
    if (i >= STATE_D && i < STATE_D + 48) {
      const int j = i - STATE_D;
      h1_h[(size_t)b * 1088 + 1024 + j] = h; h1_l[(size_t)b * 1088 + 1024 + j] = l;
      h2_h[(size_t)b * 1088 + 1024 + j] = h; h2_l[(size_t)b * 1088 + 1024 + j] = l;
      h3_h[(size_t)b * 576 + 512 + j] = h; h3_l[(size_t)b * 576 + 512 + j] = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void rollout_init_kernel(int B, int S, const float* __restrict__ init, const float* __restrict__ z,
                                    float* xin0, float* xin0_hi, float* xin0_lo, float* G0, float* t2j, float* h1, float* h2,
                                    float* h3, float* h1_lo, float* h2_lo, float* h3_lo) {
  int b = blockIdx.x;
  float* x = xin0 + (size_t)b * XIN_LD;
  for (int i = threadIdx.x; i < XIN_LD; i += blockDim.x) {
    float v = 0.f;
    if (i < STATE_D) v = init[(size_t)b * STATE_D + i];
    else if (i < STATE_D + 48) v = z[((size_t)b * S) * 48 + (i - STATE_D)];
    x[i] = v;
    if (xin0_lo) put_split(xin0_hi, xin0_lo, (size_t)b * XIN_LD + i, v);
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    float v = i < 48 ? z[((size_t)b * S) * 48 + i] : 0.f;
    put_split(h1, h1_lo, (size_t)b * 1088 + 1024 + i, v);
    put_split(h2, h2_lo, (size_t)b * 1088 + 1024 + i, v);
    put_split(h3, h3_lo, (size_t)b * 576 + 512 + i, v);
  }
  if (threadIdx.x < 12) G0[(size_t)b * 12 + threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.f : 0.f;
  if (threadIdx.x < 3) t2j[b * 4 + threadIdx.x] = threadIdx.x < 2 ? -init[(size_t)b * STATE_D + 207 + threadIdx.x] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// Glue kernels: ONE WARP PER SEQUENCE ROW, lane = joint.  Rows are staged through shared memory with
// coalesced loads/stores; the root-level reverse-mode sums (dRa, dGr, dGt, dta, dt2j) are warp-shuffle
// reductions.  Same arithmetic as glue_step_fwd / glue_step_bwd in rollout_glue.cuh (the scalar
// host/device statement that tests/host validates against autograd).
// ------------------------------------------------------------------------------------------------
constexpr int GLUE_WARPS = 4;

__global__ void __launch_bounds__(GLUE_WARPS * 32)
glue_fwd_kernel(int B, int S, int t, const float* __restrict__ xin, const float* __restrict__ raw,
                const float* __restrict__ G, const float* __restrict__ t2jg, const float* __restrict__ z,
                float* xnext, float* xnext_hi, float* xnext_lo, float* world, float* Gnext, float* h1, float* h2, float* h3,
                float* h1_lo, float* h2_lo, float* h3_lo) {
  __shared__ float s_x[GLUE_WARPS][340], s_r[GLUE_WARPS][216], s_n[GLUE_WARPS][340], s_w[GLUE_WARPS][348];
  pdl_launch_dependents();   // PDL: let the next GEMM start its prologue
  pdl_wait();                // ... and wait for the GEMM that produced `raw`
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * GLUE_WARPS + wid;
  if (b >= B) return;
  GlueFwdRow io;
  io.xr = xin + (size_t)b * XIN_LD; io.rr = raw + (size_t)b * RAW_LD; io.G = G + (size_t)b * 12; io.t2j = t2jg + b * 4;
  io.zt = (t + 1 < S) ? z + ((size_t)b * S + (t + 1)) * 48 : nullptr;
  io.xn = xnext + (size_t)b * XIN_LD;
  io.xn_hi = xnext_lo ? xnext_hi + (size_t)b * XIN_LD : nullptr; io.xn_lo = xnext_lo ? xnext_lo + (size_t)b * XIN_LD : nullptr;
  io.wo = world + (size_t)b * WORLD_LD; io.gn = Gnext + (size_t)b * 12;
  io.h1 = h1 + (size_t)b * 1088 + 1024; io.h2 = h2 + (size_t)b * 1088 + 1024; io.h3 = h3 + (size_t)b * 576 + 512;
  io.h1_lo = h1_lo ? h1_lo + (size_t)b * 1088 + 1024 : nullptr; io.h2_lo = h2_lo ? h2_lo + (size_t)b * 1088 + 1024 : nullptr;
  io.h3_lo = h3_lo ? h3_lo + (size_t)b * 576 + 512 : nullptr;
  io.xn16_h = io.xn16_l = io.h1_16h = io.h1_16l = io.h2_16h = io.h2_16l = io.h3_16h = io.h3_16l = nullptr;
  glue_fwd_warp<false>(io, lane, s_x[wid], s_r[wid], s_n[wid], s_w[wid]);
}

// reverse of one step.  have_next: grads from step t+1 exist (da0/dxres/dpx_next) and dz[:,t+1] is emitted.
__global__ void __launch_bounds__(GLUE_WARPS * 32)
glue_bwd_kernel(int B, int S, int t, int have_next, const float* __restrict__ xin, const float* __restrict__ raw,
                const float* __restrict__ G, const float* __restrict__ t2jg, const float* __restrict__ dworld,
                const float* __restrict__ da0, const float* __restrict__ dpx_next, const float* __restrict__ dh1,
                const float* __restrict__ dh2, const float* __restrict__ dh3, const float* __restrict__ dh1_lo,
                const float* __restrict__ dh2_lo, const float* __restrict__ dh3_lo, float* dxres, float* dnsum,
                const float* __restrict__ dGn_g, float* dG, float* dt2j, float* draw, float* draw_hi, float* draw_lo, float* dz) {
  __shared__ float s_x[GLUE_WARPS][340], s_r[GLUE_WARPS][216], s_dn[GLUE_WARPS][340], s_dw[GLUE_WARPS][348],
      s_dx[GLUE_WARPS][340], s_dr[GLUE_WARPS][224];
  pdl_launch_dependents();
  pdl_wait();
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * GLUE_WARPS + wid;
  if (b >= B) return;
  GlueBwdRow io;
  io.xr = xin + (size_t)b * XIN_LD; io.rr = raw + (size_t)b * RAW_LD; io.wr = dworld + (size_t)b * WORLD_LD;
  io.G = G + (size_t)b * 12; io.t2j = t2jg + b * 4; io.have_next = have_next; io.staged = 0;
  io.a0 = da0 + (size_t)b * XIN_LD; io.px = dpx_next + (size_t)b * 352; io.xs = dxres + (size_t)b * 340;
  io.dGn = dGn_g + (size_t)b * 12; io.dG = dG + (size_t)b * 12; io.dt2j = dt2j + b * 4;
  io.dzt = dz + ((size_t)b * S + (t + 1)) * 48;
  io.dh1 = dh1 + (size_t)b * 1088 + 1024; io.dh2 = dh2 + (size_t)b * 1088 + 1024; io.dh3 = dh3 + (size_t)b * 576 + 512;
  io.dh1_lo = dh1_lo ? dh1_lo + (size_t)b * 1088 + 1024 : nullptr; io.dh2_lo = dh2_lo ? dh2_lo + (size_t)b * 1088 + 1024 : nullptr;
  io.dh3_lo = dh3_lo ? dh3_lo + (size_t)b * 576 + 512 : nullptr;
  io.draw = draw + (size_t)b * RAW_LD;
  io.draw_hi = draw_lo ? draw_hi + (size_t)b * RAW_LD : nullptr; io.draw_lo = draw_lo ? draw_lo + (size_t)b * RAW_LD : nullptr;
  (void)dnsum;
  glue_bwd_warp<false>(io, lane, s_x[wid], s_r[wid], s_dn[wid], s_dw[wid], s_dx[wid], s_dr[wid]);
}

__global__ void rollout_bwd_final_kernel(int B, int S, const float* __restrict__ dxres, const float* __restrict__ da0,
                                         const float* __restrict__ dpx0, const float* __restrict__ dh1,
                                         const float* __restrict__ dh2, const float* __restrict__ dh3,
                                         const float* __restrict__ dh1_lo, const float* __restrict__ dh2_lo,
                                         const float* __restrict__ dh3_lo, const float* __restrict__ dt2j, float* dinit, float* dz) {
  int b = blockIdx.x;
  for (int i = threadIdx.x; i < STATE_D; i += blockDim.x) {
    float v = dxres[(size_t)b * 340 + i] + da0[(size_t)b * XIN_LD + i] + dpx0[(size_t)b * 352 + i];
    if (i == 207 || i == 208) v -= dt2j[b * 4 + (i - 207)];     // t2j = -(joints0.x, joints0.y, 0)
    dinit[(size_t)b * STATE_D + i] = v;
  }
  for (int i = threadIdx.x; i < 48; i += blockDim.x)
    dz[((size_t)b * S) * 48 + i] = da0[(size_t)b * XIN_LD + STATE_D + i] + get_split(dh1, dh1_lo, (size_t)b * 1088 + 1024 + i) +
                                   get_split(dh2, dh2_lo, (size_t)b * 1088 + 1024 + i) + get_split(dh3, dh3_lo, (size_t)b * 576 + 512 + i);
}

// persistent chain, after the last reverse step: d init, and d z [B][S][48] from the batched tail GEMM's [S*B][48]
__global__ void chain_bwd_final_kernel(int B, int S, const float* __restrict__ dxres, const float* __restrict__ da0,
                                       const float* __restrict__ dpx0, const float* __restrict__ dt2j, const float* __restrict__ dz_all,
                                       float* dinit, float* dz) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < STATE_D; i += blockDim.x) {
    float v = dxres[(size_t)b * 340 + i] + da0[(size_t)b * XIN_LD + i] + dpx0[(size_t)b * 352 + i];
    if (i == 207 || i == 208) v -= dt2j[b * 4 + (i - 207)];     // t2j = -(joints0.x, joints0.y, 0)
    dinit[(size_t)b * STATE_D + i] = v;
  }
  for (int i = threadIdx.x; i < S * 48; i += blockDim.x) {
    const int t = i / 48, j = i - t * 48;
    dz[((size_t)b * S + t) * 48 + j] = dz_all[((size_t)t * B + b) * 48 + j];
  }
}

// every GEMM of the rollout multiplies activations by a (frozen) weight matrix: B never depends on the previous kernel
static const int g_b_const = getenv("HB_UMMA_PREFETCH_B") ? 1 : 0;
static GemmEpi epi_gn(const float* bias, const float* g, const float* be, float* xh, int ldxh, float* rs, int C, int gs) {
  GemmEpi e;
  e.bias = bias; e.gamma = g; e.beta = be; e.xhat = xh; e.rstd = rs; e.ldxh = ldxh; e.Cch = C; e.gsize = gs;
  e.b_const = g_b_const;
  return e;
}
static GemmEpi epi_bias(const float* bias) {
  GemmEpi e;
  e.bias = bias; e.gamma = e.beta = nullptr; e.xhat = e.rstd = nullptr; e.ldxh = 0; e.Cch = 0; e.gsize = 4;
  e.b_const = g_b_const;
  return e;
}

}  // namespace hb

#ifndef HB_HOST_SHIM   // host side of the C-ABI (launch syntax): device builds only
using namespace hb;

// ---- persistent decoder chain (chain_persist.cuh): default for use_umma == 1; HB_CHAIN=0 keeps the launch-per-layer chain
static bool chain_ok(const HbHumorWeights* w, int B) {
  const char* e = getenv("HB_CHAIN");            // read per call: tests switch between the two chains inside one process
  const int on = e ? atoi(e) : 1;
  return on && (w->use_umma == 1 || w->use_umma == 2) && umma_available() && w->dec_wz_hi && w->dec_wz_lo && B <= CH_MAX_MT * 128;
}
static ChainGemm chain_gemm(int a_map, int a_col0, int a_row_step, int nkb, int b_map, int ntn, int N, int epi, int gsize, int dep_ntn) {
  ChainGemm g = {};
  g.a_map = a_map; g.b_map = b_map; g.a_col0 = a_col0; g.a_row_step = a_row_step; g.nkb = nkb; g.ntn = ntn; g.N = N; g.epi = epi;
  g.gsize = gsize; g.dep_ntn = dep_ntn;
  return g;
}
static ChainPlane chain_plane(const void* hi, const void* lo, int rows, int cols, int ld, int box_rows, int half = 0) {
  ChainPlane p; p.hi = hi; p.lo = lo; p.rows = rows; p.cols = cols; p.ld = ld; p.box_rows = box_rows; p.half = half; return p;
}
static void chain_glue_common(ChainGlue& gl, const Tape& tp, const float* z_seq, float* world) {
  gl.z = z_seq; gl.xins = tp.xins; gl.xin_hi = tp.xin_hi; gl.xin_lo = tp.xin_lo; gl.raws = tp.raws; gl.Gs = tp.Gs; gl.t2j = tp.t2j;
  gl.world = world; gl.h1 = tp.h1; gl.h1_lo = tp.h1_lo; gl.h2 = tp.h2; gl.h2_lo = tp.h2_lo; gl.h3 = tp.h3; gl.h3_lo = tp.h3_lo;
}
static cudaError_t chain_forward(const HbHumorWeights* w, const Tape& tp, int B, int S, const float* z_seq, float* world, bool f16,
                                 cudaStream_t st) {
  ChainLaunch a = {};
  a.B = B; a.S = S; a.dir = 0; a.flags = tp.chain_flags;
  if (f16) {
    // forward chain on fp16 hi / scaled-lo operand planes (x = h + l 2^-11, umma_gemm16.cuh): 4 bytes per operand element, k-blocks
    // of 64, half the MMA issue slots; the fp32 tape the reverse pass reads (xins, raws, x-hat, 1/sigma) is written exactly as before
    a.f16 = 1;
    a.planes[0] = chain_plane(tp.x16_h, tp.x16_l, S * B, X16_LD, X16_LD, 128, 1);
    a.planes[1] = chain_plane(tp.h1_16h, tp.h1_16l, B, 1088, 1088, 128, 1);
    a.planes[2] = chain_plane(tp.h2_16h, tp.h2_16l, B, 1088, 1088, 128, 1);
    a.planes[3] = chain_plane(tp.h3_16h, tp.h3_16l, B, 576, 576, 128, 1);
    a.planes[4] = chain_plane(w->dec_w16_h[0], w->dec_w16_l[0], 1024, X16_LD, X16_LD, CH_BN, 1);
    a.planes[5] = chain_plane(w->dec_w16_h[1], w->dec_w16_l[1], 1024, 1088, 1088, CH_BN, 1);
    a.planes[6] = chain_plane(w->dec_w16_h[2], w->dec_w16_l[2], 512, 1088, 1088, CH_BN, 1);
    a.planes[7] = chain_plane(w->dec_w16_h[3], w->dec_w16_l[3], 216, 576, 576, CH_BN, 1);
    ChainGemm& g0 = a.g[0]; ChainGemm& g1 = a.g[1]; ChainGemm& g2 = a.g[2]; ChainGemm& g3 = a.g[3];
    g0 = chain_gemm(0, 0, B, X16_LD / 64, 4, 16, 1024, EPI_GN_RELU, 64, 0);
    g0.bias = w->dec_b[0]; g0.gamma = w->dec_g[0]; g0.beta = w->dec_be[0]; g0.xhat = tp.dxh1; g0.ldxh = 1024; g0.rstd = tp.drs1;
    g0.C16_h = tp.h1_16h; g0.C16_l = tp.h1_16l; g0.ld16 = 1088;
    g1 = chain_gemm(1, 0, 0, 17, 5, 16, 1024, EPI_GN_RELU, 64, 16);
    g1.bias = w->dec_b[1]; g1.gamma = w->dec_g[1]; g1.beta = w->dec_be[1]; g1.xhat = tp.dxh2; g1.ldxh = 1024; g1.rstd = tp.drs2;
    g1.C16_h = tp.h2_16h; g1.C16_l = tp.h2_16l; g1.ld16 = 1088;
    g2 = chain_gemm(2, 0, 0, 17, 6, 8, 512, EPI_GN_RELU, 32, 16);
    g2.bias = w->dec_b[2]; g2.gamma = w->dec_g[2]; g2.beta = w->dec_be[2]; g2.xhat = tp.dxh3; g2.ldxh = 512; g2.rstd = tp.drs3;
    g2.C16_h = tp.h3_16h; g2.C16_l = tp.h3_16l; g2.ld16 = 576;
    g3 = chain_gemm(3, 0, 0, 9, 7, 4, 216, EPI_BIAS, 64, 8);
    g3.bias = w->dec_b[3]; g3.C = tp.raws; g3.ldc = RAW_LD; g3.c_row_step = B;
    chain_glue_common(a.glue, tp, z_seq, world);
    a.glue.x16_h = tp.x16_h; a.glue.x16_l = tp.x16_l; a.glue.x16_ld = X16_LD;
    a.glue.h1_16h = tp.h1_16h; a.glue.h1_16l = tp.h1_16l; a.glue.h2_16h = tp.h2_16h; a.glue.h2_16l = tp.h2_16l;
    a.glue.h3_16h = tp.h3_16h; a.glue.h3_16l = tp.h3_16l;
    return launch_chain(a, st);
  }
  a.planes[0] = chain_plane(tp.xin_hi, tp.xin_lo, (S + 1) * B, XIN_LD, XIN_LD, 128);
  a.planes[1] = chain_plane(tp.h1, tp.h1_lo, B, 1088, 1088, 128);
  a.planes[2] = chain_plane(tp.h2, tp.h2_lo, B, 1088, 1088, 128);
  a.planes[3] = chain_plane(tp.h3, tp.h3_lo, B, 576, 576, 128);
  a.planes[4] = chain_plane(w->dec_w_hi[0], w->dec_w_lo[0], 1024, 416, 416, CH_BN);
  a.planes[5] = chain_plane(w->dec_w_hi[1], w->dec_w_lo[1], 1024, 1088, 1088, CH_BN);
  a.planes[6] = chain_plane(w->dec_w_hi[2], w->dec_w_lo[2], 512, 1088, 1088, CH_BN);
  a.planes[7] = chain_plane(w->dec_w_hi[3], w->dec_w_lo[3], 216, 576, 576, CH_BN);
  ChainGemm& g0 = a.g[0]; ChainGemm& g1 = a.g[1]; ChainGemm& g2 = a.g[2]; ChainGemm& g3 = a.g[3];
  g0 = chain_gemm(0, 0, B, 13, 4, 16, 1024, EPI_GN_RELU, 64, 0);
  g0.bias = w->dec_b[0]; g0.gamma = w->dec_g[0]; g0.beta = w->dec_be[0]; g0.xhat = tp.dxh1; g0.ldxh = 1024; g0.rstd = tp.drs1;
  g0.C_hi = tp.h1; g0.C_lo = tp.h1_lo; g0.ldc = 1088;
  g1 = chain_gemm(1, 0, 0, 34, 5, 16, 1024, EPI_GN_RELU, 64, 16);
  g1.bias = w->dec_b[1]; g1.gamma = w->dec_g[1]; g1.beta = w->dec_be[1]; g1.xhat = tp.dxh2; g1.ldxh = 1024; g1.rstd = tp.drs2;
  g1.C_hi = tp.h2; g1.C_lo = tp.h2_lo; g1.ldc = 1088;
  g2 = chain_gemm(2, 0, 0, 34, 6, 8, 512, EPI_GN_RELU, 32, 16);
  g2.bias = w->dec_b[2]; g2.gamma = w->dec_g[2]; g2.beta = w->dec_be[2]; g2.xhat = tp.dxh3; g2.ldxh = 512; g2.rstd = tp.drs3;
  g2.C_hi = tp.h3; g2.C_lo = tp.h3_lo; g2.ldc = 576;
  g3 = chain_gemm(3, 0, 0, 18, 7, 4, 216, EPI_BIAS, 64, 8);
  g3.bias = w->dec_b[3]; g3.C = tp.raws; g3.ldc = RAW_LD; g3.c_row_step = B;
  chain_glue_common(a.glue, tp, z_seq, world);
  return launch_chain(a, st);
}
static cudaError_t chain_backward(const HbHumorWeights* w, const Tape& tp, int B, int S, const float* d_world, cudaStream_t st) {
  ChainLaunch a = {};
  a.B = B; a.S = S; a.dir = 1; a.flags = tp.chain_flags;
  a.planes[0] = chain_plane(tp.bp_hi, tp.bp_lo, S * B, BP_LD, BP_LD, 128);
  a.planes[1] = chain_plane(w->dec_wt_hi[3], w->dec_wt_lo[3], 576, 224, 224, CH_BN);
  a.planes[2] = chain_plane(w->dec_wt_hi[2], w->dec_wt_lo[2], 1088, 512, 512, CH_BN);
  a.planes[3] = chain_plane(w->dec_wt_hi[1], w->dec_wt_lo[1], 1088, 1024, 1024, CH_BN);
  a.planes[4] = chain_plane(w->dec_wt_hi[0], w->dec_wt_lo[0], 416, 1024, 1024, CH_BN);
  ChainGemm& g0 = a.g[0]; ChainGemm& g1 = a.g[1]; ChainGemm& g2 = a.g[2]; ChainGemm& g3 = a.g[3];
  g0 = chain_gemm(0, 0, B, 7, 1, 8, 512, EPI_GN_RELU_BWD, 32, 0);
  g0.gamma = w->dec_g[2]; g0.beta = w->dec_be[2]; g0.xhat = tp.dxh3; g0.ldxh = 512; g0.rstd = tp.drs3;
  g0.C_hi = tp.bp_hi; g0.C_lo = tp.bp_lo; g0.ldc = BP_LD; g0.c_col0 = BP_C3; g0.c_row_step = B;
  g1 = chain_gemm(0, BP_C3, B, 16, 2, 16, 1024, EPI_GN_RELU_BWD, 64, 8);
  g1.gamma = w->dec_g[1]; g1.beta = w->dec_be[1]; g1.xhat = tp.dxh2; g1.ldxh = 1024; g1.rstd = tp.drs2;
  g1.C_hi = tp.bp_hi; g1.C_lo = tp.bp_lo; g1.ldc = BP_LD; g1.c_col0 = BP_C2; g1.c_row_step = B;
  g2 = chain_gemm(0, BP_C2, B, 32, 3, 16, 1024, EPI_GN_RELU_BWD, 64, 16);
  g2.gamma = w->dec_g[0]; g2.beta = w->dec_be[0]; g2.xhat = tp.dxh1; g2.ldxh = 1024; g2.rstd = tp.drs1;
  g2.C_hi = tp.bp_hi; g2.C_lo = tp.bp_lo; g2.ldc = BP_LD; g2.c_col0 = BP_C1; g2.c_row_step = B;
  g3 = chain_gemm(0, BP_C1, B, 32, 4, 6, STATE_D, EPI_BIAS, 64, 16);
  g3.C = tp.da0; g3.ldc = XIN_LD;
  chain_glue_common(a.glue, tp, nullptr, nullptr);
  a.glue.dworld = d_world; a.glue.da0 = tp.da0; a.glue.dpx = tp.dpx; a.glue.dxres = tp.dxres; a.glue.dG0 = tp.dG0; a.glue.dG1 = tp.dG1;
  a.glue.dt2j = tp.dt2j; a.glue.bp_hi = tp.bp_hi; a.glue.bp_lo = tp.bp_lo; a.glue.bp_ld = BP_LD;
  return launch_chain(a, st);
}

extern "C" int humor_chain_debug(void* buf, size_t bytes) {
  chain_set_debug(static_cast<long long*>(buf), bytes);
  return HB_OK;
}

extern "C" size_t humor_rollout_workspace_bytes(int B, int S) { return carve(nullptr, B, S).total * sizeof(float); }

extern "C" int humor_rollout_fwd(const HbHumorWeights* w, int B, int S, const float* init_state, const float* z_seq,
                                 float* workspace, size_t workspace_bytes, float* world, float* prior_out,
                                 int64_t* launches, cudaStream_t st) {
  if (!w || B <= 0 || S <= 0 || !init_state || !z_seq || !workspace || !world) return HB_ERR_ARG;
  Tape tp = carve(workspace, B, S);
  if (workspace_bytes < tp.total * sizeof(float)) return HB_ERR_WORKSPACE;
  int64_t nl = 0;
  // zero the padded operand buffers once (pads must read as 0 in the K loops)
  HB_CUDA(cudaMemsetAsync(tp.xins, 0, (size_t)(S + 1) * B * XIN_LD * sizeof(float), st));
  HB_CUDA(cudaMemsetAsync(tp.raws, 0, (size_t)S * B * RAW_LD * sizeof(float), st));
  HB_CUDA(cudaMemsetAsync(tp.h1, 0, ((size_t)B * 1088 * 2 + (size_t)B * 576 + 256) * sizeof(float), st));
  const bool tc = w->use_umma && umma_available();      // tensor-core path (tcgen05 3xTF32) for every GEMM
  if (tc) {
    HB_CUDA(cudaMemsetAsync(tp.h1_lo, 0, ((size_t)B * 1088 * 2 + (size_t)B * 576) * sizeof(float), st));
    HB_CUDA(cudaMemsetAsync(tp.xin_hi, 0, (size_t)(S + 1) * B * XIN_LD * 2 * sizeof(float), st));
  }
  rollout_init_kernel<<<B, 128, 0, st>>>(B, S, init_state, z_seq, tp.xins, tc ? tp.xin_hi : nullptr, tc ? tp.xin_lo : nullptr, tp.Gs,
                                         tp.t2j, tp.h1, tp.h2, tp.h3, tc ? tp.h1_lo : nullptr, tc ? tp.h2_lo : nullptr,
                                         tc ? tp.h3_lo : nullptr);
  HB_LAUNCH_CHECK(); ++nl;
  // forward decoder chain on fp16 hi/lo planes (4 bytes per operand element; the tape the reverse pass reads is unchanged)
  const bool f16 = tc && w->use_umma == 2 && w->dec_w16_h[0] && w->dec_w16_l[0] && w->dec_w16_h[1] && w->dec_w16_l[1] &&
                   w->dec_w16_h[2] && w->dec_w16_l[2] && w->dec_w16_h[3] && w->dec_w16_l[3];
  if (f16)          // pads (input columns 416..447, hidden columns past the 48 z) must read as zero
    HB_CUDA(cudaMemsetAsync(tp.x16_h, 0, (size_t)((char*)(tp.h3_16l + (size_t)B * 576) - (char*)tp.x16_h), st));
  const int gb = cdiv(B, GLUE_WARPS);
  const bool persistent = tc && chain_ok(w, B);
  if (persistent) {                 // all S steps of the decoder chain in ONE launch
    if (f16) {                      // step 0's input row and z skip columns as fp16 planes; the glue writes those of the later steps
      chain16_pack_kernel<<<B, 128, 0, st>>>(B, tp.xins, tp.x16_h, tp.x16_l, tp.h1_16h, tp.h1_16l, tp.h2_16h, tp.h2_16l, tp.h3_16h, tp.h3_16l);
      HB_LAUNCH_CHECK(); ++nl;
    }
    HB_CUDA(chain_forward(w, tp, B, S, z_seq, world, f16, st));
    nl += 1;
  }
  for (int t = 0; t < S && !persistent; ++t) {
    const size_t r = (size_t)t * B;
    float* xin = tp.xins + r * XIN_LD;
    if (f16) {
      unsigned short* x16h = tp.x16_h + r * X16_LD;             // kept per step: the batched prior reads all of them afterwards
      unsigned short* x16l = tp.x16_l + r * X16_LD;
      chain16_pack_kernel<<<B, 128, 0, st>>>(B, xin, x16h, x16l, tp.h1_16h, tp.h1_16l, tp.h2_16h, tp.h2_16l, tp.h3_16h, tp.h3_16l);
      HB_LAUNCH_CHECK(); ++nl;
      HB_CUDA(launch_umma_gemm16(x16h, x16l, X16_LD, w->dec_w16_h[0], w->dec_w16_l[0], X16_LD, B, 1024, X16_LD, nullptr, 0,
                                 tp.h1_16h, tp.h1_16l, 1088, EPI_GN_RELU,
                                 epi_gn(w->dec_b[0], w->dec_g[0], w->dec_be[0], tp.dxh1 + r * 1024, 1024, tp.drs1 + r * 16, 1024, 64), st));
      HB_CUDA(launch_umma_gemm16(tp.h1_16h, tp.h1_16l, 1088, w->dec_w16_h[1], w->dec_w16_l[1], 1088, B, 1024, 1088, nullptr, 0,
                                 tp.h2_16h, tp.h2_16l, 1088, EPI_GN_RELU,
                                 epi_gn(w->dec_b[1], w->dec_g[1], w->dec_be[1], tp.dxh2 + r * 1024, 1024, tp.drs2 + r * 16, 1024, 64), st));
      HB_CUDA(launch_umma_gemm16(tp.h2_16h, tp.h2_16l, 1088, w->dec_w16_h[2], w->dec_w16_l[2], 1088, B, 512, 1088, nullptr, 0,
                                 tp.h3_16h, tp.h3_16l, 576, EPI_GN_RELU,
                                 epi_gn(w->dec_b[2], w->dec_g[2], w->dec_be[2], tp.dxh3 + r * 512, 512, tp.drs3 + r * 16, 512, 32), st));
      HB_CUDA(launch_umma_gemm16(tp.h3_16h, tp.h3_16l, 576, w->dec_w16_h[3], w->dec_w16_l[3], 576, B, 216, 576, tp.raws + r * RAW_LD, RAW_LD,
                                 nullptr, nullptr, 0, EPI_BIAS, epi_bias(w->dec_b[3]), st));
    } else if (tc) {
      HB_CUDA(launch_umma_gemm3(tp.xin_hi + r * XIN_LD, tp.xin_lo + r * XIN_LD, XIN_LD, w->dec_w_hi[0], w->dec_w_lo[0], 416, B, 1024, 416,
                                nullptr, tp.h1, tp.h1_lo, 1088, EPI_GN_RELU,
                                epi_gn(w->dec_b[0], w->dec_g[0], w->dec_be[0], tp.dxh1 + r * 1024, 1024, tp.drs1 + r * 16, 1024, 64), st));
      HB_CUDA(launch_umma_gemm3(tp.h1, tp.h1_lo, 1088, w->dec_w_hi[1], w->dec_w_lo[1], 1088, B, 1024, 1088, nullptr, tp.h2, tp.h2_lo, 1088,
                                EPI_GN_RELU,
                                epi_gn(w->dec_b[1], w->dec_g[1], w->dec_be[1], tp.dxh2 + r * 1024, 1024, tp.drs2 + r * 16, 1024, 64), st));
      HB_CUDA(launch_umma_gemm3(tp.h2, tp.h2_lo, 1088, w->dec_w_hi[2], w->dec_w_lo[2], 1088, B, 512, 1088, nullptr, tp.h3, tp.h3_lo, 576,
                                EPI_GN_RELU,
                                epi_gn(w->dec_b[2], w->dec_g[2], w->dec_be[2], tp.dxh3 + r * 512, 512, tp.drs3 + r * 16, 512, 32), st));
      HB_CUDA(launch_umma_gemm3(tp.h3, tp.h3_lo, 576, w->dec_w_hi[3], w->dec_w_lo[3], 576, B, 216, 576, tp.raws + r * RAW_LD, nullptr,
                                nullptr, RAW_LD, EPI_BIAS, epi_bias(w->dec_b[3]), st));
    } else {
      HB_CUDA(launch_gemm<EPI_GN_RELU>(xin, XIN_LD, w->dec_w[0], 416, tp.h1, 1088, B, 1024, 416,
                                       epi_gn(w->dec_b[0], w->dec_g[0], w->dec_be[0], tp.dxh1 + r * 1024, 1024, tp.drs1 + r * 16, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.h1, 1088, w->dec_w[1], 1088, tp.h2, 1088, B, 1024, 1088,
                                       epi_gn(w->dec_b[1], w->dec_g[1], w->dec_be[1], tp.dxh2 + r * 1024, 1024, tp.drs2 + r * 16, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.h2, 1088, w->dec_w[2], 1088, tp.h3, 576, B, 512, 1088,
                                       epi_gn(w->dec_b[2], w->dec_g[2], w->dec_be[2], tp.dxh3 + r * 512, 512, tp.drs3 + r * 16, 512, 32), st));
      HB_CUDA(launch_gemm<EPI_BIAS>(tp.h3, 576, w->dec_w[3], 576, tp.raws + r * RAW_LD, RAW_LD, B, 216, 576, epi_bias(w->dec_b[3]), st));
    }
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(gb); cfg.blockDim = dim3(GLUE_WARPS * 32); cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      HB_CUDA(cudaLaunchKernelEx(&cfg, glue_fwd_kernel, B, S, t, (const float*)xin, (const float*)(tp.raws + r * RAW_LD),
                                 (const float*)(tp.Gs + r * 12), (const float*)tp.t2j, z_seq, tp.xins + (r + B) * XIN_LD,
                                 tc ? tp.xin_hi + (r + B) * XIN_LD : (float*)nullptr, tc ? tp.xin_lo + (r + B) * XIN_LD : (float*)nullptr,
                                 world + r * WORLD_LD, tp.Gs + (r + B) * 12, tp.h1, tp.h2, tp.h3, tc ? tp.h1_lo : (float*)nullptr,
                                 tc ? tp.h2_lo : (float*)nullptr, tc ? tp.h3_lo : (float*)nullptr));
    }
    nl += 5;
  }
  if (prior_out) {
    const int M = S * B;
    bool p16 = f16;
    for (int l = 0; l < 5; ++l) p16 = p16 && w->pri_w16_h[l] && w->pri_w16_l[l];
    if (p16) {
      // batched prior on the fp16 hi/lo planes the pack kernel left for every step (K = 384 of the 448-half rows: state | pad)
      unsigned short* h16[2] = {tp.pa16_h, tp.pb16_h};
      unsigned short* l16[2] = {tp.pa16_l, tp.pb16_l};
      float* xh[4] = {tp.pxh1, tp.pxh2, tp.pxh3, tp.pxh4};
      float* rs[4] = {tp.prs1, tp.prs2, tp.prs3, tp.prs4};
      const unsigned short* a_h = tp.x16_h;
      const unsigned short* a_l = tp.x16_l;
      int lda = X16_LD, K = 384;
      for (int l = 0; l < 4; ++l) {
        HB_CUDA(launch_umma_gemm16(a_h, a_l, lda, w->pri_w16_h[l], w->pri_w16_l[l], K, M, 1024, K, nullptr, 0, h16[l & 1], l16[l & 1], 1024,
                                   EPI_GN_RELU, epi_gn(w->pri_b[l], w->pri_g[l], w->pri_be[l], xh[l], 1024, rs[l], 1024, 64), st));
        a_h = h16[l & 1]; a_l = l16[l & 1]; lda = 1024; K = 1024;
      }
      HB_CUDA(launch_umma_gemm16(a_h, a_l, 1024, w->pri_w16_h[4], w->pri_w16_l[4], 1024, M, 96, 1024, prior_out, 96, nullptr, nullptr, 0,
                                 EPI_BIAS, epi_bias(w->pri_b[4]), st));
      nl += 5;
    } else if (tc) {
      // batched prior on the 5th-gen tensor cores (3xTF32): activations travel as hi/lo planes
      // (the input planes xin_hi/xin_lo were written step by step by the glue kernel)
      float* hi[2] = {tp.pa, tp.pb};
      float* lo[2] = {tp.pa_lo, tp.pb_lo};
      float* xh[4] = {tp.pxh1, tp.pxh2, tp.pxh3, tp.pxh4};
      float* rs[4] = {tp.prs1, tp.prs2, tp.prs3, tp.prs4};
      const float* a_hi = tp.xin_hi;
      const float* a_lo = tp.xin_lo;
      int lda = XIN_LD, K = 352;
      for (int l = 0; l < 4; ++l) {
        HB_CUDA(launch_umma_gemm3(a_hi, a_lo, lda, w->pri_w_hi[l], w->pri_w_lo[l], K, M, 1024, K, nullptr, hi[l & 1], lo[l & 1], 1024,
                                  EPI_GN_RELU, epi_gn(w->pri_b[l], w->pri_g[l], w->pri_be[l], xh[l], 1024, rs[l], 1024, 64), st));
        a_hi = hi[l & 1]; a_lo = lo[l & 1]; lda = 1024; K = 1024;
      }
      HB_CUDA(launch_umma_gemm3(a_hi, a_lo, 1024, w->pri_w_hi[4], w->pri_w_lo[4], 1024, M, 96, 1024, prior_out, nullptr, nullptr, 96,
                                EPI_BIAS, epi_bias(w->pri_b[4]), st));
      nl += 5;
    } else {
      HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.xins, XIN_LD, w->pri_w[0], 352, tp.pa, 1024, M, 1024, 352,
                                       epi_gn(w->pri_b[0], w->pri_g[0], w->pri_be[0], tp.pxh1, 1024, tp.prs1, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.pa, 1024, w->pri_w[1], 1024, tp.pb, 1024, M, 1024, 1024,
                                       epi_gn(w->pri_b[1], w->pri_g[1], w->pri_be[1], tp.pxh2, 1024, tp.prs2, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.pb, 1024, w->pri_w[2], 1024, tp.pa, 1024, M, 1024, 1024,
                                       epi_gn(w->pri_b[2], w->pri_g[2], w->pri_be[2], tp.pxh3, 1024, tp.prs3, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.pa, 1024, w->pri_w[3], 1024, tp.pb, 1024, M, 1024, 1024,
                                       epi_gn(w->pri_b[3], w->pri_g[3], w->pri_be[3], tp.pxh4, 1024, tp.prs4, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_BIAS>(tp.pb, 1024, w->pri_w[4], 1024, prior_out, 96, M, 96, 1024, epi_bias(w->pri_b[4]), st));
      nl += 5;
    }
  }
  if (launches) *launches = nl;
  return HB_OK;
}

// Event recorded on the caller's stream right before the reverse decoder chain is launched: work without a consumer inside the
// closure (the dense, gradient-free LBS pass) is queued behind it on another stream and fills the SMs the latency-bound chain leaves
// idle (20 of 148 at 32 clusters x 4 CTAs) instead of holding the whole chip for 1 ms in the middle of the forward pass.
static cudaEvent_t g_chain_bwd_event = nullptr;
static bool g_chain_bwd_recorded = false;
extern "C" int humor_rollout_bwd_started_wait(cudaStream_t side) {
  if (!g_chain_bwd_event || !g_chain_bwd_recorded) return HB_ERR_ARG;
  HB_CUDA(cudaStreamWaitEvent(side, g_chain_bwd_event, 0));
  return HB_OK;
}

extern "C" int humor_rollout_bwd(const HbHumorWeights* w, int B, int S, float* workspace, size_t workspace_bytes,
                                 const float* d_world, const float* d_prior_out, float* d_init, float* d_z,
                                 int64_t* launches, cudaStream_t st) {
  if (!w || B <= 0 || S <= 0 || !workspace || !d_world || !d_init || !d_z) return HB_ERR_ARG;
  Tape tp = carve(workspace, B, S);
  if (workspace_bytes < tp.total * sizeof(float)) return HB_ERR_WORKSPACE;
  int64_t nl = 0;
  const int M = S * B;
  const bool tc = w->use_umma && umma_available();
  if (d_prior_out && tc) {
    HB_CUDA(launch_split_hilo(d_prior_out, tp.dpo_hi, tp.dpo_lo, (size_t)M * 96, st));
    float* hi[2] = {tp.pa, tp.pb};
    float* lo[2] = {tp.pa_lo, tp.pb_lo};
    float* xh[4] = {tp.pxh1, tp.pxh2, tp.pxh3, tp.pxh4};
    float* rs[4] = {tp.prs1, tp.prs2, tp.prs3, tp.prs4};
    const float* a_hi = tp.dpo_hi;
    const float* a_lo = tp.dpo_lo;
    int lda = 96, K = 96;
    for (int l = 4; l >= 1; --l) {            // reverse of Linear l, then GroupNorm/ReLU l-1
      const int o = (4 - l) & 1;
      HB_CUDA(launch_umma_gemm3(a_hi, a_lo, lda, w->pri_wt_hi[l], w->pri_wt_lo[l], K, M, 1024, K, nullptr, hi[o], lo[o], 1024,
                                EPI_GN_RELU_BWD, epi_gn(nullptr, w->pri_g[l - 1], w->pri_be[l - 1], xh[l - 1], 1024, rs[l - 1], 1024, 64), st));
      a_hi = hi[o]; a_lo = lo[o]; lda = 1024; K = 1024;
    }
    HB_CUDA(launch_umma_gemm3(a_hi, a_lo, 1024, w->pri_wt_hi[0], w->pri_wt_lo[0], 1024, M, 352, 1024, tp.dpx, nullptr, nullptr, 352,
                              EPI_BIAS, epi_bias(nullptr), st));
    nl += 6;
  } else if (d_prior_out) {
    HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(d_prior_out, 96, w->pri_wt[4], 96, tp.pa, 1024, M, 1024, 96,
                                         epi_gn(nullptr, w->pri_g[3], w->pri_be[3], tp.pxh4, 1024, tp.prs4, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.pa, 1024, w->pri_wt[3], 1024, tp.pb, 1024, M, 1024, 1024,
                                         epi_gn(nullptr, w->pri_g[2], w->pri_be[2], tp.pxh3, 1024, tp.prs3, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.pb, 1024, w->pri_wt[2], 1024, tp.pa, 1024, M, 1024, 1024,
                                         epi_gn(nullptr, w->pri_g[1], w->pri_be[1], tp.pxh2, 1024, tp.prs2, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.pa, 1024, w->pri_wt[1], 1024, tp.pb, 1024, M, 1024, 1024,
                                         epi_gn(nullptr, w->pri_g[0], w->pri_be[0], tp.pxh1, 1024, tp.prs1, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_BIAS>(tp.pb, 1024, w->pri_wt[0], 1024, tp.dpx, 352, M, 352, 1024, epi_bias(nullptr), st));
    nl += 5;
  } else {
    HB_CUDA(cudaMemsetAsync(tp.dpx, 0, (size_t)M * 352 * sizeof(float), st));
  }
  const int gb = cdiv(B, GLUE_WARPS);
  float* dGbuf[2] = {tp.dG0, tp.dG1};
  if (!g_chain_bwd_event) HB_CUDA(cudaEventCreateWithFlags(&g_chain_bwd_event, cudaEventDisableTiming));
  HB_CUDA(cudaEventRecord(g_chain_bwd_event, st));
  g_chain_bwd_recorded = true;
  if (tc && chain_ok(w, B)) {
    // all S reverse steps in ONE launch, then d z of every step as one batched GEMM over the operand planes it left
    HB_CUDA(chain_backward(w, tp, B, S, d_world, st));
    HB_CUDA(launch_umma_gemm3_bn(tp.bp_hi, tp.bp_lo, BP_LD, w->dec_wz_hi, w->dec_wz_lo, BP_LD, M, 48, BP_LD, tp.dz_all, nullptr, nullptr, 48,
                                 EPI_BIAS, epi_bias(nullptr), 64, st));
    chain_bwd_final_kernel<<<B, 128, 0, st>>>(B, S, tp.dxres, tp.da0, tp.dpx, tp.dt2j, tp.dz_all, d_init, d_z);
    HB_LAUNCH_CHECK();
    nl += 3;
    if (launches) *launches = nl;
    return HB_OK;
  }
  for (int t = S - 1; t >= 0; --t) {
    const size_t r = (size_t)t * B;
    const int have_next = (t + 1 < S);
    float* dGn = dGbuf[(t + 1) & 1];
    float* dGc = dGbuf[t & 1];
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(gb); cfg.blockDim = dim3(GLUE_WARPS * 32); cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      HB_CUDA(cudaLaunchKernelEx(&cfg, glue_bwd_kernel, B, S, t, have_next, (const float*)(tp.xins + r * XIN_LD),
                                 (const float*)(tp.raws + r * RAW_LD), (const float*)(tp.Gs + r * 12), (const float*)tp.t2j,
                                 d_world + r * WORLD_LD, (const float*)tp.da0, (const float*)(tp.dpx + (r + B) * 352),
                                 (const float*)tp.dh1, (const float*)tp.dh2, (const float*)tp.dh3,
                                 tc ? (const float*)tp.dh1_lo : (const float*)nullptr, tc ? (const float*)tp.dh2_lo : (const float*)nullptr,
                                 tc ? (const float*)tp.dh3_lo : (const float*)nullptr, tp.dxres, tp.dnsum, (const float*)dGn, dGc,
                                 tp.dt2j, tp.draw, tc ? tp.draw_hi : (float*)nullptr, tc ? tp.draw_lo : (float*)nullptr, d_z));
    }
    if (tc) {
      HB_CUDA(launch_umma_gemm3(tp.draw_hi, tp.draw_lo, RAW_LD, w->dec_wt_hi[3], w->dec_wt_lo[3], 224, B, 560, 224, nullptr, tp.dh3,
                                tp.dh3_lo, 576, EPI_GN_RELU_BWD,
                                epi_gn(nullptr, w->dec_g[2], w->dec_be[2], tp.dxh3 + r * 512, 512, tp.drs3 + r * 16, 512, 32), st));
      HB_CUDA(launch_umma_gemm3(tp.dh3, tp.dh3_lo, 576, w->dec_wt_hi[2], w->dec_wt_lo[2], 512, B, 1072, 512, nullptr, tp.dh2, tp.dh2_lo,
                                1088, EPI_GN_RELU_BWD,
                                epi_gn(nullptr, w->dec_g[1], w->dec_be[1], tp.dxh2 + r * 1024, 1024, tp.drs2 + r * 16, 1024, 64), st));
      HB_CUDA(launch_umma_gemm3(tp.dh2, tp.dh2_lo, 1088, w->dec_wt_hi[1], w->dec_wt_lo[1], 1024, B, 1072, 1024, nullptr, tp.dh1, tp.dh1_lo,
                                1088, EPI_GN_RELU_BWD,
                                epi_gn(nullptr, w->dec_g[0], w->dec_be[0], tp.dxh1 + r * 1024, 1024, tp.drs1 + r * 16, 1024, 64), st));
      HB_CUDA(launch_umma_gemm3(tp.dh1, tp.dh1_lo, 1088, w->dec_wt_hi[0], w->dec_wt_lo[0], 1024, B, 387, 1024, tp.da0, nullptr, nullptr,
                                XIN_LD, EPI_BIAS, epi_bias(nullptr), st));
    } else {
      HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.draw, RAW_LD, w->dec_wt[3], 224, tp.dh3, 576, B, 560, 224,
                                           epi_gn(nullptr, w->dec_g[2], w->dec_be[2], tp.dxh3 + r * 512, 512, tp.drs3 + r * 16, 512, 32), st));
      HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.dh3, 576, w->dec_wt[2], 512, tp.dh2, 1088, B, 1072, 512,
                                           epi_gn(nullptr, w->dec_g[1], w->dec_be[1], tp.dxh2 + r * 1024, 1024, tp.drs2 + r * 16, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.dh2, 1088, w->dec_wt[1], 1024, tp.dh1, 1088, B, 1072, 1024,
                                           epi_gn(nullptr, w->dec_g[0], w->dec_be[0], tp.dxh1 + r * 1024, 1024, tp.drs1 + r * 16, 1024, 64), st));
      HB_CUDA(launch_gemm<EPI_BIAS>(tp.dh1, 1088, w->dec_wt[0], 1024, tp.da0, XIN_LD, B, 387, 1024, epi_bias(nullptr), st));
    }
    nl += 5;
  }
  rollout_bwd_final_kernel<<<B, 128, 0, st>>>(B, S, tp.dxres, tp.da0, tp.dpx, tp.dh1, tp.dh2, tp.dh3, tc ? tp.dh1_lo : nullptr,
                                              tc ? tp.dh2_lo : nullptr, tc ? tp.dh3_lo : nullptr, tp.dt2j, d_init, d_z);
  HB_LAUNCH_CHECK(); ++nl;
  if (launches) *launches = nl;
  return HB_OK;
}
#endif  // HB_HOST_SHIM
