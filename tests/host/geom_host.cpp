// CPU build of the device geometry (humor_b200/csrc/geom.cuh) for tests: g++ -shared, loaded by ctypes.
#include "../../humor_b200/csrc/geom.cuh"
using namespace hb;
extern "C" {
void h_rodrigues_fwd(int n, const float* r, float* R) { for (int i = 0; i < n; ++i) rodrigues_fwd(r + 3 * i, R + 9 * i); }
void h_rodrigues_bwd(int n, const float* r, const float* dR, float* dr) {
  for (int i = 0; i < n; ++i) { dr[3*i]=dr[3*i+1]=dr[3*i+2]=0.f; rodrigues_bwd(r + 3 * i, dR + 9 * i, dr + 3 * i); } }
void h_mat2aa_fwd(int n, const float* R, float* aa) { for (int i = 0; i < n; ++i) mat2aa_fwd(R + 9 * i, aa + 3 * i); }
void h_mat2aa_bwd(int n, const float* R, const float* daa, float* dR) {
  for (int i = 0; i < n; ++i) { for (int k = 0; k < 9; ++k) dR[9*i+k] = 0.f; mat2aa_bwd(R + 9 * i, daa + 3 * i, dR + 9 * i); } }
void h_w2a_fwd(int n, const float* R, float* Ra) { for (int i = 0; i < n; ++i) w2a_fwd(R + 9 * i, Ra + 9 * i); }
void h_w2a_bwd(int n, const float* R, const float* dRa, float* dR) {
  for (int i = 0; i < n; ++i) { for (int k = 0; k < 9; ++k) dR[9*i+k] = 0.f; w2a_bwd(R + 9 * i, dRa + 9 * i, dR + 9 * i); } }
}

#include "../../humor_b200/csrc/rollout_glue.cuh"
#include <vector>
#include <cstring>
extern "C" {
// Chained glue-only rollout (the MLP is replaced by given raw outputs) used to validate glue_step_{fwd,bwd}.
// xins [S+1][B][416], raws [S][B][224], Gs [S+1][B][12], t2j [B][3], worlds [S][B][348]
void h_glue_rollout_fwd(int B, int S, float* xins, const float* raws, float* Gs, const float* t2j, float* worlds) {
  for (int b = 0; b < B; ++b) { float* g = Gs + b * 12; for (int i = 0; i < 12; ++i) g[i] = 0.f; g[0] = g[4] = g[8] = 1.f; }
  for (int t = 0; t < S; ++t)
    for (int b = 0; b < B; ++b)
      glue_step_fwd(xins + ((size_t)t * B + b) * XIN_LD, raws + ((size_t)t * B + b) * RAW_LD, Gs + ((size_t)t * B + b) * 12,
                    t2j + 3 * b, xins + ((size_t)(t + 1) * B + b) * XIN_LD, worlds + ((size_t)t * B + b) * WORLD_LD,
                    Gs + ((size_t)(t + 1) * B + b) * 12);
}
// dworlds [S][B][348]; outputs dxin0 [B][339], draws [S][B][216], dt2j [B][3]
void h_glue_rollout_bwd(int B, int S, const float* xins, const float* raws, const float* Gs, const float* t2j,
                        const float* dworlds, float* dxin0, float* draws, float* dt2j) {
  for (int b = 0; b < B; ++b) {
    float dn[339], dG[12], dGp[12], dx[339];
    std::memset(dn, 0, sizeof(dn)); std::memset(dG, 0, sizeof(dG));
    dt2j[3*b] = dt2j[3*b+1] = dt2j[3*b+2] = 0.f;
    for (int t = S - 1; t >= 0; --t) {
      glue_step_bwd(xins + ((size_t)t * B + b) * XIN_LD, raws + ((size_t)t * B + b) * RAW_LD, Gs + ((size_t)t * B + b) * 12,
                    t2j + 3 * b, dn, dworlds + ((size_t)t * B + b) * WORLD_LD, dG, dx, draws + ((size_t)t * B + b) * RAW_D, dGp, dt2j + 3 * b);
      std::memcpy(dn, dx, sizeof(dn)); std::memcpy(dG, dGp, sizeof(dG));
    }
    std::memcpy(dxin0 + 339 * b, dn, sizeof(dn));
  }
}
}

#include "../../humor_b200/csrc/lbs_chain.cuh"
extern "C" {
void h_lbs_chain_fwd(int n, const float* pose, const float* Jrest, const int* parents, float* feat, float* A, float* Jp) {
  for (int i = 0; i < n; ++i) lbs_chain_fwd(pose + 66 * i, Jrest + 156 * i, parents, feat + 189 * i, A + 624 * i, Jp + 156 * i);
}
void h_lbs_chain_bwd(int n, const float* pose, const float* Jrest, const int* parents, const float* dA, const float* dJp,
                     const float* dfeat, float* dpose, float* dJrest) {
  for (int i = 0; i < n; ++i)
    lbs_chain_bwd(pose + 66 * i, Jrest + 156 * i, parents, dA + 624 * i, dJp + 156 * i, dfeat + 189 * i, dpose + 66 * i, dJrest + 156 * i);
}
}
