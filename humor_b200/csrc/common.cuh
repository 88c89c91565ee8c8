#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef HB_OK          /* same values as include/humor_b200.h (sources that do not include the public header) */
#define HB_OK 0
#define HB_ERR_ARG 1001
#define HB_ERR_WORKSPACE 1002
#endif

// every launch is followed by this: the C-ABI returns the cudaError_t (non-zero) to the caller
#define HB_LAUNCH_CHECK()                           \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)
#define HB_CUDA(x)                                  \
  do {                                              \
    cudaError_t e__ = (x);                          \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)

// dynamic shared memory of a kernel as a float array.  tests/host compiles the kernels with g++ against a SIMT shim
// (HB_HOST_SHIM): there the block-shared buffer comes from the shim.
#ifdef HB_HOST_SHIM
#define HB_DYN_SMEM_F32(name) float* name = shim::dyn_smem_f32()
#else
#define HB_DYN_SMEM_F32(name) extern __shared__ __align__(16) float name[]
#endif

namespace hb {
// programmatic dependent launch (device) / nothing to order (tests/host CPU build of the kernels)
#ifdef HB_HOST_SHIM
static inline void pdl_launch_dependents() {}
static inline void pdl_wait() {}
#elif defined(__CUDACC__)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace hb
