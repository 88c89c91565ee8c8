// Test infrastructure: stand-in for the CUDA driver header when kernel sources are compiled for the host emulation
// (tests/host/shim/tc_emul.h).  Only the tensor-map handle is needed: here it simply describes the 2-D fp32 matrix a TMA
// box is cut from (the real one is an opaque 128-byte descriptor built by cuTensorMapEncodeTiled).
#pragma once
#include "cuda_runtime.h"
struct CUtensorMap {
  const float* base;              // element (row 0, col 0)
  unsigned long long rows, cols;  // extents; out-of-range box elements read as zero
  unsigned long long ld;          // row stride in floats
  unsigned box_cols, box_rows;    // box = {32 floats = 128 B (one swizzle span), box_rows}
  unsigned plain;                 // 0: SWIZZLE_128B (default); 1: SWIZZLE_NONE, box rows of box_cols floats back to back
  unsigned half;                  // 0: fp32 elements (default); 1: fp16 elements - base points at halves, ld / cols / box_cols count halves
};
#define __grid_constant__
