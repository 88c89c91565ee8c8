// Argument blocks of the persistent decoder-chain kernel (chain_persist.cuh): plain data shared by the kernel, its launcher
// (umma_gemm.cu) and the rollout driver (rollout.cu).
#pragma once
#include "gemm.cuh"

namespace hb {

constexpr int CH_CS = 4;            // CTAs per cluster = split-K ways of every GEMM tile
constexpr int CH_BN = 64;           // output-tile columns
constexpr int CH_MAX_MT = 4;        // 128-row tiles: B <= 512 sub-sequences per GPU
constexpr int CH_MAX_NT = 16;       // 64-column tiles per layer (N <= 1024)
constexpr int CH_NGEMM = 4;         // GEMM phases per step (the four decoder layers)
constexpr int CH_NMAPS = 8;         // operand planes (hi; lo = the same index in the lo array)
constexpr int CH_FLAGS = CH_NGEMM * CH_MAX_MT * CH_MAX_NT + CH_MAX_MT;     // tile flags, then one glue flag per row tile

// One GEMM phase:  C[B, N] (+)= A[B, K] * W[N, K]^T  with a fused epilogue.  A and C live in buffers that either hold one step
// (row step 0: overwritten every step) or all steps (row step B: the tape).
struct ChainGemm {
  int a_map, b_map;         // operand plane indices
  int a_col0;               // first column of the operand inside the A buffer
  int a_row_step;           // rows the A operand advances per step
  int nkb;                  // K / 32
  int ntn;                  // 64-column output tiles
  int N;                    // valid output columns
  int epi;                  // EPI_BIAS | EPI_GN_RELU | EPI_GN_RELU_BWD
  int gsize;                // GroupNorm group width (64 or 32)
  int dep_ntn;              // output tiles of the phase that produces A (k-block kb needs tile kb / 2; k-blocks past them: the glue)
  const float* bias; const float* gamma; const float* beta;
  float* xhat; int ldxh;    // GroupNorm tape [S*B][ldxh] (written forward, read in reverse)
  float* rstd;              // [S*B][16]
  float* C; float* C_hi; float* C_lo;   // fp32 result and/or its hi/lo planes (nullable)
  int ldc, c_col0, c_row_step;
  unsigned short* C16_h; unsigned short* C16_l; int ld16;   // fp16 hi / scaled-lo planes of the result (forward chain on 4-byte operand elements)
};

struct ChainGlue {
  // forward (reference: models/humor_model.py:961-1001)
  const float* z;           // [B][S][48]
  float* xins; float* xin_hi; float* xin_lo;    // [(S+1)*B][XIN_LD] step inputs and their planes
  float* raws;              // [S*B][RAW_LD] decoder outputs
  float* Gs;                // [(S+1)*B][12]
  const float* t2j;         // [B][4]
  float* world;             // [S*B][WORLD_LD]
  float *h1, *h1_lo, *h2, *h2_lo, *h3, *h3_lo;   // hidden-activation planes [B][1088 | 1088 | 576]: the glue writes the z skip columns
  // reverse
  const float* dworld;      // [S*B][WORLD_LD]
  const float* da0;         // [B][XIN_LD] d xin from the first decoder layer (written by the last GEMM phase)
  const float* dpx;         // [S*B][352] d xin from the batched prior
  float* dxres;             // [B][340]
  float* dG0; float* dG1;   // [B][12] ping-pong
  float* dt2j;              // [B][4]
  float* bp_hi; float* bp_lo; int bp_ld;   // per-step reverse operand planes [S*B][bp_ld]: d raw | d pre3 | d pre2 | d pre1
  // forward chain on fp16 hi / scaled-lo planes (ChainLaunch::f16): step inputs [S*B][x16_ld] and the z skip columns of the hidden planes
  unsigned short *x16_h, *x16_l; int x16_ld;
  unsigned short *h1_16h, *h1_16l, *h2_16h, *h2_16l, *h3_16h, *h3_16l;
};

struct ChainPlane { const void* hi; const void* lo; int rows, cols, ld, box_rows, half; };   // half: fp16 elements, 64-wide boxes

struct ChainLaunch {
  ChainPlane planes[CH_NMAPS];
  ChainGemm g[CH_NGEMM];
  ChainGlue glue;
  unsigned* flags;          // CH_FLAGS words, zeroed by the launcher
  int B, S, dir;            // dir 0: forward steps 0..S-1; 1: reverse steps S-1..0
  int f16;                  // forward only: operands are fp16 hi / scaled-lo planes (x = h + l * 2^-11; umma_gemm16.cuh), k-blocks of 64
};
constexpr int CH_DBG_EV = 16;       // humor_chain_debug: clock64 stamps of CTA 0, [step][phase 0..4][CH_DBG_EV]

}  // namespace hb
