// Test infrastructure: a CUDA runtime stand-in (LD_PRELOADed into a child python process) that EXECUTES kernels on the CPU.
// The product library is linked once more with `-cudart shared` (same nvcc-compiled host code: argument checks, workspace
// carving, dispatch); every cudaLaunchKernel it issues is resolved here to the SAME kernel source compiled with g++ against
// the SIMT shim (tests/host/shim) and run with the launch's grid/block on host memory.  "Device pointers" are host pointers.
// Covers every SIMT kernel (lbs.cu, rot.cu, losses.cu, chamfer.cu, rollout.cu) and, on the functional tcgen05 / TMA / TMEM
// emulation of tests/host/shim/tc_emul.h, the single-CTA tcgen05 kernels (umma_gemm3_kernel<.,.,1>,
// lbs_fuseg_kernel) and the 4-CTA-cluster split-K instantiations of umma_gemm3_kernel (the CTAs of a cluster
// run concurrently, DSMEM stores and cluster barriers are emulated); cuTensorMapEncodeTiled is emulated too, so the library's
// own descriptor code runs.
// Not a product path: nothing in humor_b200/ references it; the product rejects CPU tensors unless a test patches that out.
#include <cxxabi.h>
#include <dlfcn.h>

#include <cstddef>
#include <cstdio>
#include <map>
#include <string>

#define hb hb_emu          // the emulated kernels live in their own namespace: the library's host stubs keep `hb::`
#include "../../../humor_b200/csrc/lbs.cu"
#include "../../../humor_b200/csrc/rot.cu"
#include "../../../humor_b200/csrc/losses.cu"
#include "../../../humor_b200/csrc/chamfer.cu"
#include "../../../humor_b200/csrc/rollout.cu"
#include "../../../humor_b200/csrc/umma_gemm.cu"     // tcgen05 kernels on tests/host/shim/tc_emul.h + split_hilo_kernel
#undef hb

namespace {
using Thunk = void (*)(dim3, dim3, void**);
template <class T> T& arg(void** a, int i) { return *static_cast<T*>(a[i]); }
typedef const float* cf;
typedef const int* ci;
#define A(T, i) arg<T>(a, i)
unsigned g_cluster_x = 1;        // cluster dimension of the launch being executed (cudaLaunchKernelExC attribute)
#define RUN(...) shim::launch_cluster(g, b, g_cluster_x, [&] { __VA_ARGS__; })
const std::map<std::string, Thunk>& registry() {
  static const std::map<std::string, Thunk> r = {
      {"hb::lbs_pose_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_pose_kernel(A(HbLbsModel, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(float*, 7), A(float*, 8), A(float*, 9), A(int, 10), A(float*, 11), A(float*, 12), A(float, 13), A(int, 14))); }},
      {"hb::lbs_pose_warp_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_pose_warp_kernel(A(HbLbsModel, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(float*, 7), A(float*, 8), A(float*, 9), A(int, 10), A(float*, 11), A(float*, 12), A(float, 13), A(int, 14))); }},
      {"hb::lbs_pose_bwd_warp_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_pose_bwd_warp_kernel(A(HbLbsModel, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(cf, 7), A(cf, 8), A(cf, 9), A(int, 10), A(float*, 11), A(float*, 12), A(float*, 13), A(float*, 14))); }},
      {"hb::lbs_pose_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_pose_bwd_kernel(A(HbLbsModel, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(cf, 7), A(cf, 8), A(cf, 9), A(int, 10), A(float*, 11), A(float*, 12), A(float*, 13), A(float*, 14))); }},
      {"hb::lbs_skin_fwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_skin_fwd_kernel(A(HbLbsModel, 0), A(int, 1), A(cf, 2), A(cf, 3), A(cf, 4), A(ci, 5), A(int, 6), A(float*, 7), A(size_t, 8))); }},
      {"hb::lbs_skin_apply_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_skin_apply_kernel(A(HbLbsModel, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(float*, 6))); }},
      {"hb::lbs_shape_rows_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_shape_rows_kernel(A(HbLbsModel, 0), A(int, 1), A(cf, 2), A(float, 3), A(float*, 4))); }},
      {"hb::lbs_gather_extra_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_gather_extra_kernel(A(HbLbsModel, 0), A(int, 1), A(cf, 2), A(float*, 3))); }},
      {"hb::lbs_skin_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::lbs_skin_bwd_kernel(A(HbLbsModel, 0), A(int, 1), A(cf, 2), A(cf, 3), A(ci, 4), A(int, 5), A(cf, 6), A(size_t, 7), A(float*, 8), A(float*, 9), A(float*, 10), A(int, 11), A(ci, 12), A(int, 13), A(cf, 14), A(size_t, 15))); }},
      {"hb::rodrigues_fwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::rodrigues_fwd_kernel(A(int, 0), A(cf, 1), A(float*, 2))); }},
      {"hb::rodrigues_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::rodrigues_bwd_kernel(A(int, 0), A(cf, 1), A(cf, 2), A(float*, 3))); }},
      {"hb::mat2aa_fwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::mat2aa_fwd_kernel(A(int, 0), A(cf, 1), A(float*, 2))); }},
      {"hb::mat2aa_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::mat2aa_bwd_kernel(A(int, 0), A(cf, 1), A(cf, 2), A(float*, 3))); }},
      {"hb::cam2prior_fwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::cam2prior_fwd_kernel(A(int, 0), A(cf, 1), A(cf, 2), A(int, 3), A(cf, 4), A(int, 5), A(cf, 6), A(int, 7), A(float*, 8), A(float*, 9), A(float*, 10))); }},
      {"hb::cam2prior_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::cam2prior_bwd_kernel(A(int, 0), A(cf, 1), A(cf, 2), A(int, 3), A(cf, 4), A(int, 5), A(cf, 6), A(int, 7), A(cf, 8), A(cf, 9), A(cf, 10), A(float*, 11), A(float*, 12), A(float*, 13), A(float*, 14))); }},
      {"hb::rollout_outputs_fwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::rollout_outputs_fwd_kernel(A(hb_emu::RollOut, 0))); }},
      {"hb::rollout_outputs_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::rollout_outputs_bwd_kernel(A(hb_emu::RollOutBwd, 0))); }},
      {"hb::fit_losses_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::fit_losses_kernel(A(HbFitArgs, 0))); }},
      {"hb::fit_reduce1_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::fit_reduce1_kernel(A(HbFitArgs, 0))); }},
      {"hb::fit_reduce_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::fit_reduce_kernel(A(HbFitArgs, 0))); }},
      {"hb::gmm_nll_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::gmm_nll_kernel(A(int, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(cf, 7), A(float*, 8), A(float*, 9))); }},
      {"hb::gmm_pass1_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::gmm_pass1_kernel(A(int, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(cf, 7), A(float*, 8), A(float*, 9))); }},
      {"hb::gmm_pass2_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::gmm_pass2_kernel(A(int, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(float*, 6))); }},
      {"hb::gmm_reduce_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::gmm_reduce_kernel(A(int, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(float*, 5), A(float*, 6))); }},
      {"hb::rollout_init_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::rollout_init_kernel(A(int, 0), A(int, 1), A(cf, 2), A(cf, 3), A(float*, 4), A(float*, 5), A(float*, 6), A(float*, 7), A(float*, 8), A(float*, 9), A(float*, 10), A(float*, 11), A(float*, 12), A(float*, 13), A(float*, 14))); }},
      {"hb::glue_fwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::glue_fwd_kernel(A(int, 0), A(int, 1), A(int, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(cf, 7), A(float*, 8), A(float*, 9), A(float*, 10), A(float*, 11), A(float*, 12), A(float*, 13), A(float*, 14), A(float*, 15), A(float*, 16), A(float*, 17), A(float*, 18))); }},
      {"hb::glue_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::glue_bwd_kernel(A(int, 0), A(int, 1), A(int, 2), A(int, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(cf, 7), A(cf, 8), A(cf, 9), A(cf, 10), A(cf, 11), A(cf, 12), A(cf, 13), A(cf, 14), A(cf, 15), A(cf, 16), A(float*, 17), A(float*, 18), A(cf, 19), A(float*, 20), A(float*, 21), A(float*, 22), A(float*, 23), A(float*, 24), A(float*, 25))); }},
      {"hb::chain16_pack_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::chain16_pack_kernel(A(int, 0), A(cf, 1), A(unsigned short*, 2), A(unsigned short*, 3), A(unsigned short*, 4), A(unsigned short*, 5), A(unsigned short*, 6), A(unsigned short*, 7), A(unsigned short*, 8), A(unsigned short*, 9))); }},
      {"hb::rollout_bwd_final_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::rollout_bwd_final_kernel(A(int, 0), A(int, 1), A(cf, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(cf, 7), A(cf, 8), A(cf, 9), A(cf, 10), A(cf, 11), A(float*, 12), A(float*, 13))); }},
#define GEMM_THUNK(BM, BN, BK, NSM, NSN, EPI)                                                                                   \
      {"hb::gemm_tn_kernel<" #BM ", " #BN ", " #BK ", " #NSM ", " #NSN ", " #EPI ">", [](dim3 g, dim3 b, void** a) {            \
         RUN((hb_emu::gemm_tn_kernel<BM, BN, BK, NSM, NSN, EPI>(A(cf, 0), A(int, 1), A(cf, 2), A(int, 3), A(float*, 4), A(int, 5), A(int, 6), A(int, 7), A(int, 8), A(hb_emu::GemmEpi, 9)))); }},
      GEMM_THUNK(128, 128, 16, 2, 2, 0) GEMM_THUNK(128, 128, 16, 2, 2, 1) GEMM_THUNK(128, 128, 16, 2, 2, 2)
      GEMM_THUNK(32, 64, 32, 1, 1, 0) GEMM_THUNK(32, 64, 32, 1, 1, 1) GEMM_THUNK(32, 64, 32, 1, 1, 2)
#define MAP(i) (*static_cast<CUtensorMap*>(a[i]))     /* the first bytes of the 128-byte driver object hold the emulated map */
#define UMMA_THUNK(BN, EPI)                                                                                                     \
      {"hb::umma_gemm3_kernel<" #BN ", " #EPI ", 1>", [](dim3 g, dim3 b, void** a) {                                            \
         tcemu::reset();                                                                                                        \
         RUN((hb_emu::umma_gemm3_kernel<BN, EPI, 1>(MAP(0), MAP(1), MAP(2), MAP(3), A(int, 4), A(int, 5), A(int, 6), A(float*, 7), A(float*, 8), A(float*, 9), A(int, 10), A(hb_emu::GemmEpi, 11)))); }},
      UMMA_THUNK(64, 0) UMMA_THUNK(64, 1) UMMA_THUNK(64, 2) UMMA_THUNK(128, 0) UMMA_THUNK(128, 1) UMMA_THUNK(128, 2)
#define UMMAP_THUNK(EPI)   /* persistent 128x128 tiles, two epilogue groups: CTAs are independent, run one after another */             \
      {"hb::umma_gemm3p_kernel<" #EPI ">", [](dim3 g, dim3 b, void** a) {                                                       \
         tcemu::reset();                                                                                                        \
         RUN((hb_emu::umma_gemm3p_kernel<EPI>(MAP(0), MAP(1), MAP(2), MAP(3), A(int, 4), A(int, 5), A(int, 6), A(float*, 7), A(float*, 8), A(float*, 9), A(int, 10), A(hb_emu::GemmEpi, 11)))); }},
      UMMAP_THUNK(0) UMMAP_THUNK(1) UMMAP_THUNK(2)
#define UMMA_SPLITK_THUNK(EPI)   /* split-K over a 4-CTA cluster: the four CTAs run concurrently, partials meet through DSMEM stores */      \
      {"hb::umma_gemm3_kernel<64, " #EPI ", 4>", [](dim3 g, dim3 b, void** a) {                                                 \
         tcemu::reset();                                                                                                        \
         if (g_cluster_x != 4) { std::fprintf(stderr, "cudart_emul: split-K GEMM launched without its 4-CTA cluster\n"); std::abort(); }  \
         RUN((hb_emu::umma_gemm3_kernel<64, EPI, 4>(MAP(0), MAP(1), MAP(2), MAP(3), A(int, 4), A(int, 5), A(int, 6), A(float*, 7), A(float*, 8), A(float*, 9), A(int, 10), A(hb_emu::GemmEpi, 11)))); }},
      UMMA_SPLITK_THUNK(0) UMMA_SPLITK_THUNK(1) UMMA_SPLITK_THUNK(2)
#define UMMA16_THUNK(BN, EPI, KS)                                                                                                \
      {"hb::umma_gemm16_kernel<" #BN ", " #EPI ", " #KS ">", [](dim3 g, dim3 b, void** a) {                                     \
         tcemu::reset();                                                                                                        \
         if (g_cluster_x != KS) { std::fprintf(stderr, "cudart_emul: umma_gemm16 cluster size mismatch\n"); std::abort(); }     \
         RUN((hb_emu::umma_gemm16_kernel<BN, EPI, KS>(MAP(0), MAP(1), MAP(2), MAP(3), A(int, 4), A(int, 5), A(int, 6), A(float*, 7), A(int, 8), A(unsigned short*, 9), A(unsigned short*, 10), A(int, 11), A(hb_emu::GemmEpi, 12)))); }},
      UMMA16_THUNK(64, 0, 1) UMMA16_THUNK(64, 0, 4) UMMA16_THUNK(128, 0, 1) UMMA16_THUNK(64, 1, 1) UMMA16_THUNK(64, 1, 4) UMMA16_THUNK(128, 1, 1)
      {"hb::split16_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::split16_kernel(A(cf, 0), A(unsigned short*, 1), A(unsigned short*, 2), A(size_t, 3))); }},
      {"hb::lbs_fuseg_kernel", [](dim3 g, dim3 b, void** a) { tcemu::reset(); RUN(hb_emu::lbs_fuseg_kernel(MAP(0), MAP(1), MAP(2), MAP(3), MAP(4), MAP(5), MAP(6), MAP(7), MAP(8), A(int, 9), A(hb_emu::LbsFusegArgs, 10))); }},
      {"hb::feat_f16_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::feat_f16_kernel(A(cf, 0), A(int, 1), A(int, 2), A(int, 3), A(int, 4), A(int, 5), A(unsigned short*, 6), A(unsigned short*, 7), A(int, 8))); }},
      // persistent decoder chain: the WHOLE grid runs at once (clusters wait on each other through global-memory flags); the
      // launcher's parameter block holds 128-byte driver tensor maps, the emulated kernel's block the emulated ones
      {"hb::chain_kernel", [](dim3 g, dim3 b, void** a) {
         tcemu::reset();
         if (g_cluster_x != (unsigned)hb_emu::CH_CS || g.x > (unsigned)tcemu::MAX_CTAS) { std::fprintf(stderr, "cudart_emul: chain_kernel grid %u / cluster %u not emulated\n", g.x, g_cluster_x); std::abort(); }
         const unsigned char* raw = static_cast<const unsigned char*>(a[0]);
         static hb_emu::ChainParams q;
         for (int i = 0; i < hb_emu::CH_NMAPS; ++i) {
           std::memcpy(&q.map_hi[i], raw + 128 * i, sizeof(CUtensorMap));
           std::memcpy(&q.map_lo[i], raw + 128 * (hb_emu::CH_NMAPS + i), sizeof(CUtensorMap));
         }
         std::memcpy(reinterpret_cast<unsigned char*>(&q) + offsetof(hb_emu::ChainParams, g), raw + 2 * hb_emu::CH_NMAPS * 128,
                     offsetof(hb_emu::ChainParams, f16) + sizeof(int) - offsetof(hb_emu::ChainParams, g));     // everything but the debug pointer
         shim::launch_cluster(g, b, g_cluster_x, [&] { hb_emu::chain_kernel(q); }, true); }},
      {"hb::chain_bwd_final_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::chain_bwd_final_kernel(A(int, 0), A(int, 1), A(cf, 2), A(cf, 3), A(cf, 4), A(cf, 5), A(cf, 6), A(float*, 7), A(float*, 8))); }},
      {"hb::split_hilo_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::split_hilo_kernel(A(cf, 0), A(float*, 1), A(float*, 2), A(size_t, 3))); }},
      {"hb::chamfer_nn_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::chamfer_nn_kernel(A(int, 0), A(cf, 1), A(int, 2), A(cf, 3), A(float*, 4), A(int*, 5))); }},
      {"hb::chamfer_bwd_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::chamfer_bwd_kernel(A(int, 0), A(cf, 1), A(int, 2), A(cf, 3), A(cf, 4), A(ci, 5), A(cf, 6), A(ci, 7), A(float*, 8), A(float*, 9))); }},
      {"hb::chamfer_fill_zero_kernel", [](dim3 g, dim3 b, void** a) { RUN(hb_emu::chamfer_fill_zero_kernel(A(size_t, 0), A(float*, 1), A(int*, 2))); }},
  };
  return r;
}
long long g_launches = 0;
int run_kernel(const void* f, dim3 g, dim3 b, void** args) {
  Dl_info i;
  std::string name = "?";
  if (dladdr(f, &i) && i.dli_sname) {
    int st = 0;
    char* d = abi::__cxa_demangle(i.dli_sname, nullptr, nullptr, &st);
    name = d ? d : i.dli_sname;
    std::free(d);
  }
  const size_t paren = name.find('(');
  if (paren != std::string::npos) name.resize(paren);
  if (name.rfind("void ", 0) == 0) name = name.substr(5);
  const auto it = registry().find(name);
  if (it == registry().end()) {
    std::fprintf(stderr, "cudart_emul: kernel '%s' has no CPU emulation registered\n", name.c_str());
    return 98;                                   // cudaErrorInvalidDeviceFunction
  }
  if (getenv("HB_EMUL_TRACE")) std::fprintf(stderr, "EMUL %s grid=(%u,%u,%u) block=%u\n", name.c_str(), g.x, g.y, g.z, b.x);
  it->second(g, b, args);
  ++g_launches;
  return 0;
}
}  // namespace

extern "C" {
int cudaLaunchKernel(const void* f, dim3 g, dim3 b, void** a, size_t, void*) { g_cluster_x = 1; return run_kernel(f, g, b, a); }
struct EmulLaunchAttr { int id; int pad; unsigned val[16]; };                       // cudaLaunchAttribute: id, padding to 8, 64-byte value
struct EmulLaunchConfig { dim3 grid; dim3 block; size_t dynamicSmemBytes; void* stream; EmulLaunchAttr* attrs; unsigned numAttrs; };   // cudaLaunchConfig_t
int cudaLaunchKernelExC(const EmulLaunchConfig* c, const void* f, void** a) {
  g_cluster_x = 1;
  for (unsigned i = 0; i < c->numAttrs; ++i)
    if (c->attrs[i].id == 4 /* cudaLaunchAttributeClusterDimension */) {
      if (c->attrs[i].val[1] != 1 || c->attrs[i].val[2] != 1) { std::fprintf(stderr, "cudart_emul: only x-clusters are emulated\n"); return 1; }
      g_cluster_x = c->attrs[i].val[0];
    }
  const int rc = run_kernel(f, c->grid, c->block, a);
  g_cluster_x = 1;
  return rc;
}
int cudaPeekAtLastError() { return 0; }
// the shim header already has a static cudaGetLastError for the kernel sources: export the runtime symbol under an asm label
int emul_get_last_error() __asm__("cudaGetLastError");
int emul_get_last_error() { return 0; }
int cudaFuncSetAttribute(const void*, int, int) { return 0; }
int cudaOccupancyMaxActiveClusters(int* n, const void*, const void*) { *n = tcemu::MAX_CTAS / 4; return 0; }
int cudaGetDevice(int* d) { *d = 0; return 0; }
int cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return 0; }
int cudaMemsetAsync(void* p, int v, size_t n, void*) { std::memset(p, v, n); return 0; }
// events: execution is synchronous and in program order, so ordering primitives have nothing to do
int cudaEventCreateWithFlags(void** e, unsigned) { static int dummy; *e = &dummy; return 0; }
int cudaEventRecord(void*, void*) { return 0; }
int cudaStreamWaitEvent(void*, void*, unsigned) { return 0; }
// cuTensorMapEncodeTiled: 2-D fp32 maps only; the emulated description is stored in the caller's (128-byte) CUtensorMap
static int emul_encode_tiled(void* m, int dtype, unsigned rank, void* base, const unsigned long long* gdim, const unsigned long long* gstr,
                             const unsigned* box, const unsigned*, int, int swizzle, int, int) {
  const bool half = dtype == 6;   /* CU_TENSOR_MAP_DATA_TYPE_FLOAT16 */
  const unsigned esz = half ? 2 : 4;
  if (rank != 2 || (dtype != 7 /* FLOAT32 */ && !half) || (swizzle != 3 /* SWIZZLE_128B */ && swizzle != 0 /* NONE */) ||
      (half && swizzle != 3) || (gstr[0] % 16) || ((box[0] * esz) % 16) || box[0] > 256 || box[1] > 256)
    return 1;
  CUtensorMap e{static_cast<const float*>(base), gdim[1], gdim[0], gstr[0] / esz, box[0], box[1], swizzle == 0 ? 1u : 0u, half ? 1u : 0u};
  std::memcpy(m, &e, sizeof(e));
  return 0;
}
int cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, int* q) {
  const bool ok = std::strcmp(name, "cuTensorMapEncodeTiled") == 0;
  *fn = ok ? (void*)emul_encode_tiled : nullptr;
  if (q) *q = ok ? 0 : 1;
  return 0;
}
long long hb_emul_launches() { return g_launches; }
}
