// Test infrastructure: a minimal SIMT-on-CPU stand-in for <cuda_runtime.h>, found first on the include path when a
// kernel source is compiled with g++ for tests/host/*.cpp.  One std::thread per CUDA thread of a block, blocks run one
// after another (the blocks of a thread-block cluster together); __syncthreads() is a real barrier, __shared__ becomes a
// function-local static (shared by the threads of the running block; kernels launched as clusters must use dynamic shared
// memory).  Enough to execute the index arithmetic, tiling, tails and barrier structure of plain SIMT kernels on the host
// with the same source the GPU runs.  Not a product path: nothing in humor_b200/ includes this.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(8) uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }

typedef int cudaError_t;
typedef struct CUstream_st* cudaStream_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return 0; }

namespace shim {
// per-CTA state of the running launch.  Blocks normally run one after another; the blocks of a thread-block CLUSTER run
// concurrently (launch_cluster), each with its own context, so everything a kernel thread touches is reached through
// thread-local pointers.
struct Cta {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
  std::vector<uint32_t> warp_xchg;         // one slot per thread: shuffle exchange area
  std::unique_ptr<std::barrier<>> epi_bar; // named barrier of threads 64.. (bar.sync 1, nt - 64)
  std::unique_ptr<std::barrier<>> pair_bar[2];   // named barriers of warp pairs (2, 3) and (4, 5): bar.sync 2 + i, 64
};
inline thread_local uint3 t_idx{0, 0, 0};
inline thread_local uint3 t_bidx{0, 0, 0};
inline thread_local Cta* t_cta = nullptr;
inline thread_local unsigned t_crank = 0;          // rank of the thread's CTA inside its cluster
inline thread_local unsigned t_slot = 0;           // per-CTA emulation state slot (= rank; = block index when the whole grid runs at once)
inline thread_local std::barrier<>* t_cluster_bar = nullptr;
inline unsigned cluster_size = 1;
inline dim3 b_dim, g_dim;
inline thread_local int t_lin = 0;
alignas(1024) inline float dyn_smem_buf[64 * 1024];    // 256 KB: dynamic shared memory of the running block (plain SIMT kernels)
inline float* dyn_smem_f32() { return dyn_smem_buf; }
inline void sync_block() { t_cta->bar->arrive_and_wait(); }
inline void sync_warp() { t_cta->warp_bars[t_lin / 32]->arrive_and_wait(); }
inline void sync_cluster() { t_cluster_bar->arrive_and_wait(); }
inline void sync_pair(int pi) { t_cta->pair_bar[pi]->arrive_and_wait(); }

// Runs `body` once per thread of every block of the grid; `csize` consecutive blocks along x form a cluster and run together.
// concurrent = true: ALL clusters of a 1-D grid run at once (persistent kernels whose clusters wait on each other through
// global-memory flags); every CTA then has its own emulation state slot (t_slot = block index).
template <class F>
void launch_cluster(dim3 grid, dim3 block, unsigned csize, F body, bool concurrent = false) {
  g_dim = grid;
  b_dim = block;
  cluster_size = csize;
  const int nt = (int)(block.x * block.y * block.z);
  auto make_ctas = [&](std::vector<Cta>& ctas) {
    for (auto& c : ctas) {
      c.bar.reset(new std::barrier<>(nt));
      for (int w = 0; w < (nt + 31) / 32; ++w) c.warp_bars.emplace_back(new std::barrier<>(std::min(32, nt - 32 * w)));
      c.warp_xchg.assign(nt, 0u);
      if (nt > 64) c.epi_bar.reset(new std::barrier<>(nt - 64));
      if (nt >= 192) { c.pair_bar[0].reset(new std::barrier<>(64)); c.pair_bar[1].reset(new std::barrier<>(64)); }
    }
  };
  if (concurrent) {
    const unsigned nb = grid.x, ncl = (nb + csize - 1) / csize;
    std::vector<Cta> ctas(nb);
    make_ctas(ctas);
    std::vector<std::unique_ptr<std::barrier<>>> cbars;
    for (unsigned c = 0; c < ncl; ++c) cbars.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(csize, nb - c * csize) * nt));
    std::vector<std::thread> th;
    th.reserve((size_t)nb * nt);
    for (unsigned bx = 0; bx < nb; ++bx)
      for (int t = 0; t < nt; ++t)
        th.emplace_back([=, &body, &ctas, &cbars] {
          t_lin = t;
          t_idx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
          t_bidx = {bx, 0, 0};
          t_cta = &ctas[bx];
          t_crank = bx % csize;
          t_slot = bx;
          t_cluster_bar = cbars[bx / csize].get();
          body();
        });
    for (auto& x : th) x.join();
    cluster_size = 1;
    return;
  }
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx0 = 0; bx0 < grid.x; bx0 += csize) {
        const unsigned nc = std::min(csize, grid.x - bx0);
        std::vector<Cta> ctas(nc);
        make_ctas(ctas);
        std::barrier<> cbar((std::ptrdiff_t)nc * nt);
        std::vector<std::thread> th;
        th.reserve((size_t)nc * nt);
        for (unsigned r = 0; r < nc; ++r)
          for (int t = 0; t < nt; ++t)
            th.emplace_back([=, &body, &ctas, &cbar] {
              t_lin = t;
              t_idx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
              t_bidx = {bx0 + r, by, bz};
              t_cta = &ctas[r];
              t_crank = r;
              t_slot = r;
              t_cluster_bar = &cbar;
              body();
            });
        for (auto& x : th) x.join();
      }
  cluster_size = 1;
}
template <class F>
void launch(dim3 grid, dim3 block, F body) { launch_cluster(grid, block, 1u, body); }
}  // namespace shim

#define threadIdx (shim::t_idx)
#define blockIdx (shim::t_bidx)
#define blockDim (shim::b_dim)
#define gridDim (shim::g_dim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __constant__ static const
#define __launch_bounds__(...)
#define __syncthreads() shim::sync_block()
#define __syncwarp(...) shim::sync_warp()
#define __align__(n) __attribute__((aligned(n)))

// Separately rounded fp32 operations (compile the harness with -ffp-contract=off so that plain expressions are not
// fused either).
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __ldg(const float* p) { return *p; }
static inline float __ldcg(const float* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline void __stcs(float* p, float v) { *p = v; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  shim::t_cta->warp_xchg[shim::t_lin] = __float_as_uint(v);
  shim::sync_warp();
  const float r = __uint_as_float(shim::t_cta->warp_xchg[(shim::t_lin & ~31) | ((shim::t_lin ^ lane_mask) & 31)]);
  shim::sync_warp();
  return r;
}
static inline float __shfl_sync(unsigned, float v, int src) {
  shim::t_cta->warp_xchg[shim::t_lin] = __float_as_uint(v);
  shim::sync_warp();
  const float r = __uint_as_float(shim::t_cta->warp_xchg[(shim::t_lin & ~31) | (src & 31)]);
  shim::sync_warp();
  return r;
}
using std::max;
using std::min;
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline int __ldg(const unsigned* p) { return (int)*p; }
typedef unsigned long long cuuint64_t;
