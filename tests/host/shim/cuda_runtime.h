// Test infrastructure: a minimal SIMT-on-CPU stand-in for <cuda_runtime.h>, found first on the include path when a
// kernel source is compiled with g++ for tests/host/*.cpp.  One std::thread per CUDA thread of a block, blocks run one
// after another; __syncthreads() is a real barrier, __shared__ becomes a function-local static (shared by the threads
// of the running block).  Enough to execute the index arithmetic, tiling, tails and barrier structure of plain SIMT
// kernels on the host with the same source the GPU runs.  Not a product path: nothing in humor_b200/ includes this.
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
inline thread_local uint3 t_idx{0, 0, 0};
inline uint3 b_idx{0, 0, 0};
inline dim3 b_dim, g_dim;
inline std::barrier<>* block_bar = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
inline thread_local int t_lin = 0;
inline std::vector<uint32_t> warp_xchg;     // one slot per thread: shuffle exchange area
alignas(1024) inline float dyn_smem_buf[64 * 1024];    // 256 KB: dynamic shared memory of the running block
inline float* dyn_smem_f32() { return dyn_smem_buf; }
inline void sync_block() { block_bar->arrive_and_wait(); }
inline void sync_warp() { warp_bars[t_lin / 32]->arrive_and_wait(); }

// Runs `body` once per thread of every block of the grid.
template <class F>
void launch(dim3 grid, dim3 block, F body) {
  g_dim = grid;
  b_dim = block;
  const int nt = (int)(block.x * block.y * block.z);
  warp_xchg.assign(nt, 0u);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        b_idx = {bx, by, bz};
        std::barrier<> bar(nt);
        block_bar = &bar;
        warp_bars.clear();
        for (int w = 0; w < (nt + 31) / 32; ++w) warp_bars.emplace_back(new std::barrier<>(std::min(32, nt - 32 * w)));
        std::vector<std::thread> th;
        th.reserve(nt);
        for (int t = 0; t < nt; ++t)
          th.emplace_back([=, &body] {
            t_lin = t;
            t_idx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
            body();
          });
        for (auto& x : th) x.join();
      }
}
}  // namespace shim

#define threadIdx (shim::t_idx)
#define blockIdx (shim::b_idx)
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
static inline int __ldg(const int* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline void __stcs(float* p, float v) { *p = v; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  shim::warp_xchg[shim::t_lin] = __float_as_uint(v);
  shim::sync_warp();
  const float r = __uint_as_float(shim::warp_xchg[(shim::t_lin & ~31) | ((shim::t_lin ^ lane_mask) & 31)]);
  shim::sync_warp();
  return r;
}
static inline float __shfl_sync(unsigned, float v, int src) {
  shim::warp_xchg[shim::t_lin] = __float_as_uint(v);
  shim::sync_warp();
  const float r = __uint_as_float(shim::warp_xchg[(shim::t_lin & ~31) | (src & 31)]);
  shim::sync_warp();
  return r;
}
using std::max;
using std::min;
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline int __ldg(const unsigned* p) { return (int)*p; }
typedef unsigned long long cuuint64_t;
