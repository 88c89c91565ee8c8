// Test infrastructure: lbs_skin_group_kernel (humor_b200/csrc/lbs_skin_group.cuh) executed on the CPU through the SIMT shim
// with the grid/block shape humor_lbs_fwd uses.
//   g++ -O1 -std=c++20 -pthread -shared -fPIC -Itests/host/shim -DHB_HOST_SHIM lbs_skin_host.cpp
#include "../../humor_b200/csrc/lbs_skin_group.cuh"
using namespace hb;
extern "C" void h_lbs_skin_group(int num_verts, int num_groups, const int* g_start, const int* g_joint, const float* g_w,
                                 int nframes, int v3_ld, const float* vposed, const float* A, const float* trans, float* out,
                                 int groups_per_block) {
  HbLbsModel m{};
  m.num_verts = num_verts;
  m.num_groups = num_groups;
  m.g_start = g_start; m.g_joint = g_joint; m.g_w = g_w;
  shim::launch(dim3(cdiv(num_groups, groups_per_block), cdiv(nframes, SG_FT)), dim3(SG_WARPS * 32),
               [&] { lbs_skin_group_kernel(m, nframes, v3_ld, vposed, A, trans, out, groups_per_block); });
}
