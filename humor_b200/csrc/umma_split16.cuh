// x = h + l * 2^-11 with h = fp16(x), l = fp16((x - h) * 2^11): the operand split of the fp16 hi/lo tcgen05 GEMM (umma_gemm16.cuh),
// shared with the kernels that produce operand planes (rollout.cu).  Bit patterns, so that no fp16 type crosses a header.
#pragma once
#include <stdint.h>
#ifdef HB_HOST_SHIM
#include "tc_emul.h"
#else
#include <cuda_fp16.h>
#endif

namespace hb {
// one value -> (h, l) halves, as bit patterns
#ifdef HB_HOST_SHIM
static inline void split16(float x, unsigned short& h, unsigned short& l) {
  h = tcemu::f32_to_f16_bits(x);
  _Float16 hf; std::memcpy(&hf, &h, 2);
  l = tcemu::f32_to_f16_bits((x - (float)hf) * 2048.f);
}
#else
__device__ __forceinline__ void split16(float x, unsigned short& h, unsigned short& l) {
  const __half hh = __float2half_rn(x);
  h = __half_as_ushort(hh);
  l = __half_as_ushort(__float2half_rn((x - __half2float(hh)) * 2048.f));
}
#endif

}  // namespace hb
