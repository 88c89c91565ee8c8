// fp32 FFMA GEMM  C[M,N] = A[M,K] * B[N,K]^T  (both operands K-contiguous) with fused epilogues
// for the HuMoR MLP stack (models/humor_model.py:1206-1241: Linear -> GroupNorm(16) -> ReLU):
//   EPI_BIAS         C = acc + bias
//   EPI_GN_RELU      y = acc + bias ; per-row GroupNorm over groups of `gsize` channels ;
//                    C = relu(gamma*xhat+beta) ; also stores xhat and rstd for the backward
//   EPI_GN_RELU_BWD  acc is dL/dh ; columns < Cch: dL/dy of relu(GN(y)) (uses saved xhat, rstd) ;
//                    columns >= Cch (skip-connected z) pass through
// K must be a multiple of BK = 32 (operands are stored zero-padded), lda/ldb multiples of 4.
// This is the exact-fp32 path required by the 1e-5 parity bound on decoder states / prior log-prob.
#pragma once
#include "common.cuh"

namespace hb {

enum { EPI_BIAS = 0, EPI_GN_RELU = 1, EPI_GN_RELU_BWD = 2 };

struct GemmEpi {
  const float* bias;    // [N]   (EPI_BIAS, EPI_GN_RELU)
  const float* gamma;   // [Cch]
  const float* beta;    // [Cch]
  float* xhat;          // [M][ldxh]  written by EPI_GN_RELU, read by EPI_GN_RELU_BWD
  float* rstd;          // [M][16]
  int ldxh;
  int Cch;              // number of normalised channels
  int gsize;            // channels per group (64 or 32)
  // tcgen05 path only: the B operand (a weight matrix) is not written by the predecessor kernel of the stream, so its first tiles
  // may be fetched before the programmatic-dependent-launch wait (opt-in, HB_UMMA_PREFETCH_B=1; 0 = everything after the wait)
  int b_const = 0;
};

template <int BM, int BN, int BK, int NSM, int NSN, int EPI>
__global__ void __launch_bounds__((BM / (4 * NSM)) * (BN / (4 * NSN)))
gemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
               float* __restrict__ C, int ldc, int M, int N, int K, GemmEpi ep) {
  constexpr int TX = BN / (4 * NSN);
  constexpr int TY = BM / (4 * NSM);
  constexpr int NT = TX * TY;
  static_assert(TX == 16, "group reductions assume 16 threads along N");
  constexpr int A_V4 = BM * BK / 4;      // float4 loads per tile
  constexpr int B_V4 = BN * BK / 4;
  constexpr int A_PT = (A_V4 + NT - 1) / NT;
  constexpr int B_PT = (B_V4 + NT - 1) / NT;
  constexpr int KV = BK / 4;

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[NSM * 4][NSN * 4];
#pragma unroll
  for (int i = 0; i < NSM * 4; ++i)
#pragma unroll
    for (int j = 0; j < NSN * 4; ++j) acc[i][j] = 0.f;

  // global -> register -> (transposed) shared staging with a prefetch distance of TWO tiles: with one
  // CTA per SM (the sequential decoder steps only offer ~128 CTAs) a single tile of compute does not
  // cover the L2 latency, two do.
  float4 ra[2][A_PT], rb[2][B_PT];
  auto gload = [&](int k0, float4* pa, float4* pb) {
#pragma unroll
    for (int i = 0; i < A_PT; ++i) {
      int idx = tid + i * NT;
      int r = idx / KV, kv = idx % KV;
      pa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < A_V4 && m0 + r < M) pa[i] = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k0 + kv * 4);
    }
#pragma unroll
    for (int i = 0; i < B_PT; ++i) {
      int idx = tid + i * NT;
      int r = idx / KV, kv = idx % KV;
      pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_V4 && n0 + r < N) pb[i] = *reinterpret_cast<const float4*>(B + (size_t)(n0 + r) * ldb + k0 + kv * 4);
    }
  };
  auto sstore = [&](int buf, const float4* pa, const float4* pb) {
#pragma unroll
    for (int i = 0; i < A_PT; ++i) {
      int idx = tid + i * NT;
      if (idx < A_V4) {
        int r = idx / KV, kv = idx % KV;
        As[buf][kv * 4 + 0][r] = pa[i].x; As[buf][kv * 4 + 1][r] = pa[i].y;
        As[buf][kv * 4 + 2][r] = pa[i].z; As[buf][kv * 4 + 3][r] = pa[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_PT; ++i) {
      int idx = tid + i * NT;
      if (idx < B_V4) {
        int r = idx / KV, kv = idx % KV;
        Bs[buf][kv * 4 + 0][r] = pb[i].x; Bs[buf][kv * 4 + 1][r] = pb[i].y;
        Bs[buf][kv * 4 + 2][r] = pb[i].z; Bs[buf][kv * 4 + 3][r] = pb[i].w;
      }
    }
  };
  auto compute = [&](int buf) {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[NSM * 4], b[NSN * 4];
#pragma unroll
      for (int s = 0; s < NSM; ++s) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][s * (BM / NSM) + ty * 4]);
        a[s * 4 + 0] = v.x; a[s * 4 + 1] = v.y; a[s * 4 + 2] = v.z; a[s * 4 + 3] = v.w;
      }
#pragma unroll
      for (int s = 0; s < NSN; ++s) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][s * (BN / NSN) + tx * 4]);
        b[s * 4 + 0] = v.x; b[s * 4 + 1] = v.y; b[s * 4 + 2] = v.z; b[s * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < NSM * 4; ++i)
#pragma unroll
        for (int j = 0; j < NSN * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  };

  const int nk = K / BK;
  gload(0, ra[0], rb[0]);
  if (nk > 1) gload(BK, ra[1], rb[1]);
  sstore(0, ra[0], rb[0]);
  __syncthreads();
  if (nk > 2) gload(2 * BK, ra[0], rb[0]);
  // tile kt+1 sits in register stage (kt+1)&1; after it is stored, that stage prefetches tile kt+3
  for (int kt = 0; kt < nk; kt += 2) {
    compute(0);
    if (kt + 1 < nk) {
      sstore(1, ra[1], rb[1]);
      __syncthreads();
      if (kt + 3 < nk) gload((kt + 3) * BK, ra[1], rb[1]);
      compute(1);
      if (kt + 2 < nk) {
        sstore(0, ra[0], rb[0]);
        __syncthreads();
        if (kt + 4 < nk) gload((kt + 4) * BK, ra[0], rb[0]);
      }
    }
  }

  // ------------------------------------------------------------------ epilogue
  const int glanes = (EPI == EPI_BIAS) ? 1 : ep.gsize / 4;   // lanes (along tx) that share one GN group
#pragma unroll
  for (int i = 0; i < NSM * 4; ++i) {
    const int row = m0 + (i / 4) * (BM / NSM) + ty * 4 + (i % 4);
    const bool rok = row < M;
#pragma unroll
    for (int s = 0; s < NSN; ++s) {
      const int col = n0 + s * (BN / NSN) + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][s * 4 + j];
      if (EPI == EPI_BIAS) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (rok && col + j < N) C[(size_t)row * ldc + col + j] = v[j] + (ep.bias ? ep.bias[col + j] : 0.f);
      } else if (EPI == EPI_GN_RELU) {
        // N == Cch is a multiple of BN/NSN: whole groups live inside one 16-lane row segment
        const bool cok = col < N;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] += cok ? ep.bias[col + j] : 0.f; sum += v[j]; }
        for (int o = 1; o < glanes; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)ep.gsize;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { float d = v[j] - mean; sq += d * d; }
        for (int o = 1; o < glanes; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rs = rsqrtf(sq / (float)ep.gsize + 1e-5f);
        if (rok && cok) {
          float4 xh, h;
          float* xp = &xh.x; float* hp = &h.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            xp[j] = (v[j] - mean) * rs;
            hp[j] = fmaxf(fmaf(ep.gamma[col + j], xp[j], ep.beta[col + j]), 0.f);
          }
          *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = h;
          *reinterpret_cast<float4*>(ep.xhat + (size_t)row * ep.ldxh + col) = xh;
          if ((col % ep.gsize) == 0) ep.rstd[(size_t)row * 16 + col / ep.gsize] = rs;
        }
      } else {  // EPI_GN_RELU_BWD
        const bool cok = col < N;
        const bool isgn = col < ep.Cch;        // warp-uniform per segment: Cch % 64 == 0
        float u[4] = {0.f, 0.f, 0.f, 0.f}, xh[4] = {0.f, 0.f, 0.f, 0.f};
        float s1 = 0.f, s2 = 0.f;
        if (isgn && rok && cok) {
          float4 x4 = *reinterpret_cast<const float4*>(ep.xhat + (size_t)row * ep.ldxh + col);
          xh[0] = x4.x; xh[1] = x4.y; xh[2] = x4.z; xh[3] = x4.w;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float g = ep.gamma[col + j];
            bool on = fmaf(g, xh[j], ep.beta[col + j]) > 0.f;
            u[j] = on ? g * v[j] : 0.f;
            s1 += u[j];
            s2 += u[j] * xh[j];
          }
        }
        for (int o = 1; o < glanes; o <<= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (rok && cok) {
          if (isgn) {
            const float rs = ep.rstd[(size_t)row * 16 + col / ep.gsize];
            const float inv = 1.f / (float)ep.gsize;
            float4 o4;
            float* op = &o4.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = rs * (u[j] - s1 * inv - xh[j] * s2 * inv);
            *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = o4;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (col + j < N) C[(size_t)row * ldc + col + j] = v[j];
          }
        }
      }
    }
  }
}

#ifndef HB_HOST_SHIM   // launch syntax: device builds only (tests/host compiles the kernels with g++)
// Host-side dispatch: small-M tiles (32x64) for the sequential decoder steps, 128x128 for the
// batched prior.  Returns a cudaError_t.
template <int EPI>
static inline cudaError_t launch_gemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                                      int M, int N, int K, const GemmEpi& ep, cudaStream_t st) {
  if (M >= 1024) {
    dim3 grid(cdiv(N, 128), cdiv(M, 128));
    gemm_tn_kernel<128, 128, 16, 2, 2, EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, ep);
  } else {
    dim3 grid(cdiv(N, 64), cdiv(M, 32));
    gemm_tn_kernel<32, 64, 32, 1, 1, EPI><<<grid, 128, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, ep);
  }
  return cudaGetLastError();
}
#endif

}  // namespace hb
