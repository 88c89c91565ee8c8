// Test infrastructure: a fake CUDA runtime (LD_PRELOADed) whose launches succeed and print the kernel symbol, so that the HOST
// dispatch of the C-ABI (which kernels, which grids) can be exercised on a machine without a GPU (tests/test_host_dispatch.py).
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cxxabi.h>
typedef int cudaError_t;
struct dim3 { unsigned x, y, z; };
static void report(const void* f, dim3 g, dim3 b, size_t sm) {
  Dl_info i; const char* n = "?";
  if (dladdr(f, &i) && i.dli_sname) n = i.dli_sname;
  int st; char* d = abi::__cxa_demangle(n, 0, 0, &st);
  const char* s = d ? d : n; char buf[80]; strncpy(buf, s, 79); buf[79] = 0; char* p = strchr(buf, '('); if (p) *p = 0;
  printf("LAUNCH %s grid=(%u,%u,%u) block=%u smem=%zu\n", buf, g.x, g.y, g.z, b.x, sm);
}
extern "C" {
cudaError_t cudaLaunchKernel(const void* f, dim3 g, dim3 b, void** a, size_t sm, void* st) { report(f, g, b, sm); return 0; }
struct cudaLaunchConfig_t { dim3 gridDim; dim3 blockDim; size_t dynamicSmemBytes; void* stream; void* attrs; unsigned numAttrs; };
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t* c, const void* f, void** a) { report(f, c->gridDim, c->blockDim, c->dynamicSmemBytes); return 0; }
cudaError_t cudaGetLastError() { return 0; }
cudaError_t cudaEventCreateWithFlags(void** e, unsigned) { static int dummy; *e = &dummy; return 0; }
cudaError_t cudaEventRecord(void*, void*) { return 0; }
cudaError_t cudaStreamWaitEvent(void*, void*, unsigned) { return 0; }
cudaError_t cudaPeekAtLastError() { return 0; }
cudaError_t cudaFuncSetAttribute(const void*, int, int) { return 0; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return 0; }
static int fake_encode(void*, int, unsigned, void*, const void*, const void*, const void*, const void*, int, int, int, int) { return 0; }
cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, unsigned long long, int* q) { *fn = (void*)fake_encode; if (q) *q = 0; return 0; }
cudaError_t cudaGetDriverEntryPointByVersion(const char*, void** fn, unsigned, unsigned long long, int* q) { *fn = (void*)fake_encode; if (q) *q = 0; return 0; }
}
