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
#include "gemm.cuh"
#include "rollout_glue.cuh"
#include "../../include/humor_b200.h"

namespace hb {

struct Tape {
  float *xins, *raws, *Gs, *t2j;
  float *dxh1, *dxh2, *dxh3, *drs1, *drs2, *drs3;
  float *pxh1, *pxh2, *pxh3, *pxh4, *prs1, *prs2, *prs3, *prs4;
  float *h1, *h2, *h3;                 // decoder transients [B][1088],[B][1088],[B][576]
  float *pa, *pb;                      // prior ping-pong [S*B][1024]
  float *dh3, *dh2, *dh1, *da0, *draw, *dxres, *dnsum, *dG0, *dG1, *dt2j, *dpx;
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
  t.total = off;
  return t;
}

// ------------------------------------------------------------------------------------------------
__global__ void rollout_init_kernel(int B, int S, const float* __restrict__ init, const float* __restrict__ z,
                                    float* xin0, float* G0, float* t2j, float* h1, float* h2, float* h3) {
  int b = blockIdx.x;
  float* x = xin0 + (size_t)b * XIN_LD;
  for (int i = threadIdx.x; i < XIN_LD; i += blockDim.x) {
    float v = 0.f;
    if (i < STATE_D) v = init[(size_t)b * STATE_D + i];
    else if (i < STATE_D + 48) v = z[((size_t)b * S) * 48 + (i - STATE_D)];
    x[i] = v;
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    float v = i < 48 ? z[((size_t)b * S) * 48 + i] : 0.f;
    h1[(size_t)b * 1088 + 1024 + i] = v;
    h2[(size_t)b * 1088 + 1024 + i] = v;
    h3[(size_t)b * 576 + 512 + i] = v;
  }
  if (threadIdx.x < 12) G0[(size_t)b * 12 + threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.f : 0.f;
  if (threadIdx.x < 3) t2j[b * 4 + threadIdx.x] = threadIdx.x < 2 ? -init[(size_t)b * STATE_D + 207 + threadIdx.x] : 0.f;
}

// one thread per sequence row
__global__ void glue_fwd_kernel(int B, int S, int t, const float* __restrict__ xin, const float* __restrict__ raw,
                                const float* __restrict__ G, const float* __restrict__ t2j, const float* __restrict__ z,
                                float* xnext, float* world, float* Gnext, float* h1, float* h2, float* h3) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* xn = xnext + (size_t)b * XIN_LD;
  glue_step_fwd(xin + (size_t)b * XIN_LD, raw + (size_t)b * RAW_LD, G + (size_t)b * 12, t2j + b * 4, xn,
                world + (size_t)b * WORLD_LD, Gnext + (size_t)b * 12);
  if (t + 1 < S) {
    const float* zt = z + ((size_t)b * S + (t + 1)) * 48;
    for (int i = 0; i < 48; ++i) {
      float v = zt[i];
      xn[STATE_D + i] = v;
      h1[(size_t)b * 1088 + 1024 + i] = v;
      h2[(size_t)b * 1088 + 1024 + i] = v;
      h3[(size_t)b * 576 + 512 + i] = v;
    }
  }
}

// reverse of one step.  have_next: grads from step t+1 exist (da0/dxres/dpx_next) and dz[:,t+1] is emitted.
__global__ void glue_bwd_kernel(int B, int S, int t, int have_next, const float* __restrict__ xin,
                                const float* __restrict__ raw, const float* __restrict__ G, const float* __restrict__ t2j,
                                const float* __restrict__ dworld, const float* __restrict__ da0,
                                const float* __restrict__ dpx_next, const float* __restrict__ dh1,
                                const float* __restrict__ dh2, const float* __restrict__ dh3, float* dxres, float* dnsum,
                                const float* __restrict__ dGn, float* dG, float* dt2j, float* draw, float* dz) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* dn = dnsum + (size_t)b * 340;
  float* dx = dxres + (size_t)b * 340;
  float dGnext[12];
  if (have_next) {
    const float* a0 = da0 + (size_t)b * XIN_LD;
    const float* px = dpx_next + (size_t)b * 352;
    for (int i = 0; i < STATE_D; ++i) dn[i] = dx[i] + a0[i] + px[i];
    float* dzt = dz + ((size_t)b * S + (t + 1)) * 48;
    for (int i = 0; i < 48; ++i)
      dzt[i] = a0[STATE_D + i] + dh1[(size_t)b * 1088 + 1024 + i] + dh2[(size_t)b * 1088 + 1024 + i] +
               dh3[(size_t)b * 576 + 512 + i];
    for (int i = 0; i < 12; ++i) dGnext[i] = dGn[(size_t)b * 12 + i];
  } else {
    for (int i = 0; i < STATE_D; ++i) dn[i] = 0.f;
    for (int i = 0; i < 12; ++i) dGnext[i] = 0.f;
    dt2j[b * 4 + 0] = dt2j[b * 4 + 1] = dt2j[b * 4 + 2] = 0.f;
  }
  float* dr = draw + (size_t)b * RAW_LD;
  glue_step_bwd(xin + (size_t)b * XIN_LD, raw + (size_t)b * RAW_LD, G + (size_t)b * 12, t2j + b * 4, dn,
                dworld + (size_t)b * WORLD_LD, dGnext, dx, dr, dG + (size_t)b * 12, dt2j + b * 4);
  for (int i = RAW_D; i < RAW_LD; ++i) dr[i] = 0.f;
}

__global__ void rollout_bwd_final_kernel(int B, int S, const float* __restrict__ dxres, const float* __restrict__ da0,
                                         const float* __restrict__ dpx0, const float* __restrict__ dh1,
                                         const float* __restrict__ dh2, const float* __restrict__ dh3,
                                         const float* __restrict__ dt2j, float* dinit, float* dz) {
  int b = blockIdx.x;
  for (int i = threadIdx.x; i < STATE_D; i += blockDim.x) {
    float v = dxres[(size_t)b * 340 + i] + da0[(size_t)b * XIN_LD + i] + dpx0[(size_t)b * 352 + i];
    if (i == 207 || i == 208) v -= dt2j[b * 4 + (i - 207)];     // t2j = -(joints0.x, joints0.y, 0)
    dinit[(size_t)b * STATE_D + i] = v;
  }
  for (int i = threadIdx.x; i < 48; i += blockDim.x)
    dz[((size_t)b * S) * 48 + i] = da0[(size_t)b * XIN_LD + STATE_D + i] + dh1[(size_t)b * 1088 + 1024 + i] +
                                   dh2[(size_t)b * 1088 + 1024 + i] + dh3[(size_t)b * 576 + 512 + i];
}

static GemmEpi epi_gn(const float* bias, const float* g, const float* be, float* xh, int ldxh, float* rs, int C, int gs) {
  GemmEpi e;
  e.bias = bias; e.gamma = g; e.beta = be; e.xhat = xh; e.rstd = rs; e.ldxh = ldxh; e.Cch = C; e.gsize = gs;
  return e;
}
static GemmEpi epi_bias(const float* bias) {
  GemmEpi e;
  e.bias = bias; e.gamma = e.beta = nullptr; e.xhat = e.rstd = nullptr; e.ldxh = 0; e.Cch = 0; e.gsize = 4;
  return e;
}

}  // namespace hb

using namespace hb;

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
  rollout_init_kernel<<<B, 128, 0, st>>>(B, S, init_state, z_seq, tp.xins, tp.Gs, tp.t2j, tp.h1, tp.h2, tp.h3);
  HB_LAUNCH_CHECK(); ++nl;
  const int gb = cdiv(B, 32);
  for (int t = 0; t < S; ++t) {
    const size_t r = (size_t)t * B;
    float* xin = tp.xins + r * XIN_LD;
    HB_CUDA(launch_gemm<EPI_GN_RELU>(xin, XIN_LD, w->dec_w[0], 416, tp.h1, 1088, B, 1024, 416,
                                     epi_gn(w->dec_b[0], w->dec_g[0], w->dec_be[0], tp.dxh1 + r * 1024, 1024, tp.drs1 + r * 16, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.h1, 1088, w->dec_w[1], 1088, tp.h2, 1088, B, 1024, 1088,
                                     epi_gn(w->dec_b[1], w->dec_g[1], w->dec_be[1], tp.dxh2 + r * 1024, 1024, tp.drs2 + r * 16, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_GN_RELU>(tp.h2, 1088, w->dec_w[2], 1088, tp.h3, 576, B, 512, 1088,
                                     epi_gn(w->dec_b[2], w->dec_g[2], w->dec_be[2], tp.dxh3 + r * 512, 512, tp.drs3 + r * 16, 512, 32), st));
    HB_CUDA(launch_gemm<EPI_BIAS>(tp.h3, 576, w->dec_w[3], 576, tp.raws + r * RAW_LD, RAW_LD, B, 216, 576, epi_bias(w->dec_b[3]), st));
    glue_fwd_kernel<<<gb, 32, 0, st>>>(B, S, t, xin, tp.raws + r * RAW_LD, tp.Gs + r * 12, tp.t2j, z_seq,
                                       tp.xins + (r + B) * XIN_LD, world + r * WORLD_LD, tp.Gs + (r + B) * 12, tp.h1, tp.h2, tp.h3);
    HB_LAUNCH_CHECK();
    nl += 5;
  }
  if (prior_out) {
    const int M = S * B;
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
  if (launches) *launches = nl;
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
  if (d_prior_out) {
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
  const int gb = cdiv(B, 32);
  float* dGbuf[2] = {tp.dG0, tp.dG1};
  for (int t = S - 1; t >= 0; --t) {
    const size_t r = (size_t)t * B;
    const int have_next = (t + 1 < S);
    float* dGn = dGbuf[(t + 1) & 1];
    float* dGc = dGbuf[t & 1];
    glue_bwd_kernel<<<gb, 32, 0, st>>>(B, S, t, have_next, tp.xins + r * XIN_LD, tp.raws + r * RAW_LD, tp.Gs + r * 12, tp.t2j,
                                       d_world + r * WORLD_LD, tp.da0, tp.dpx + (r + B) * 352, tp.dh1, tp.dh2, tp.dh3,
                                       tp.dxres, tp.dnsum, dGn, dGc, tp.dt2j, tp.draw, d_z);
    HB_LAUNCH_CHECK();
    HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.draw, RAW_LD, w->dec_wt[3], 224, tp.dh3, 576, B, 560, 224,
                                         epi_gn(nullptr, w->dec_g[2], w->dec_be[2], tp.dxh3 + r * 512, 512, tp.drs3 + r * 16, 512, 32), st));
    HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.dh3, 576, w->dec_wt[2], 512, tp.dh2, 1088, B, 1072, 512,
                                         epi_gn(nullptr, w->dec_g[1], w->dec_be[1], tp.dxh2 + r * 1024, 1024, tp.drs2 + r * 16, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_GN_RELU_BWD>(tp.dh2, 1088, w->dec_wt[1], 1024, tp.dh1, 1088, B, 1072, 1024,
                                         epi_gn(nullptr, w->dec_g[0], w->dec_be[0], tp.dxh1 + r * 1024, 1024, tp.drs1 + r * 16, 1024, 64), st));
    HB_CUDA(launch_gemm<EPI_BIAS>(tp.dh1, 1088, w->dec_wt[0], 1024, tp.da0, XIN_LD, B, 387, 1024, epi_bias(nullptr), st));
    nl += 5;
  }
  rollout_bwd_final_kernel<<<B, 128, 0, st>>>(B, S, tp.dxres, tp.da0, tp.dpx, tp.dh1, tp.dh2, tp.dh3, tp.dt2j, d_init, d_z);
  HB_LAUNCH_CHECK(); ++nl;
  if (launches) *launches = nl;
  return HB_OK;
}
