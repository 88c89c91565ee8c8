// Dense skinning pass, second form: lane = frame, thread-private groups of 8 consecutive vertices.
//
//   out[n][v] = sum_j W[v][j] * (A[n][j] . [v_posed[n][v]; 1]) + trans[n]          (A: 3x4 rest->posed transforms)
//
// The first form (lbs_skin_apply_kernel: lane = vertex) gathers <= 4 transforms of 48 B per (vertex, frame) pair from
// shared memory: 192 B/pair, every pair, half of it bank conflicts (ncu: 12.7 M wavefronts per 512-frame slab, 6.5 M of
// them conflicts; profiles/r01g_lbs_slab_kernels_set_full.ncu-rep) - the shared-memory port is the bound.  Here a warp is
// 32 frames and walks vertex GROUPS: for a group the host has tabulated the union of the joints its 8 vertices are
// skinned to and, per joint, the 8 weights (zeros where a vertex is not influenced).  A thread loads a joint's
// transform of ITS frame once (3 x LDS.128, conflict-free: frame stride 628 floats = 20 mod 32 banks) and applies it to
// every vertex of the group that uses it; joint index and weights are warp-uniform (plain broadcast loads, uniform
// branches).  Shared-memory traffic drops from 192 B/pair to 48 B x |union| / 8 per pair (SMPL-like meshes: ~5 joints
// per group -> 30 B/pair; the benchmark's synthetic asset with one random joint per vertex: ~11 -> 66 B/pair).
// v_posed rows are read straight from L2 (96 contiguous bytes per thread and group), results written with 8-byte
// stores (an output row is 82 680 B: 8- but not 16-byte aligned).
#pragma once
#include "common.cuh"
#include "../../include/humor_b200.h"

namespace hb {

constexpr int SG_G = 8;            // vertices per group
constexpr int SG_AS = 628;         // floats per frame of the staged transforms (52*12 + 4)
constexpr int SG_FT = 32;          // frames per block = lanes of a warp
constexpr int SG_WARPS = 8;
constexpr int SG_SMEM = SG_FT * SG_AS * (int)sizeof(float);

// grid = (ceil(num_groups / groups_per_block), ceil(nframes / 32)), block = 256
__global__ void __launch_bounds__(SG_WARPS * 32, 2)
lbs_skin_group_kernel(HbLbsModel m, int nframes, int v3_ld, const float* __restrict__ vposed, const float* __restrict__ A,
                      const float* __restrict__ trans, float* __restrict__ out, int groups_per_block) {
  HB_DYN_SMEM_F32(As);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f0 = blockIdx.y * SG_FT;
  const int nf = min(SG_FT, nframes - f0);
  // stage the transforms of the block's frames: 156 float4 per frame, coalesced
  for (int i = tid; i < SG_FT * 156; i += SG_WARPS * 32) {
    const int f = i / 156, c = i - f * 156;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < nf) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(f0 + f) * 624) + c);
    *reinterpret_cast<float4*>(As + f * SG_AS + 4 * c) = v;
  }
  __syncthreads();
  const bool fok = lane < nf;
  const int frame = f0 + (fok ? lane : 0);
  const float t0 = trans[(size_t)frame * 3], t1 = trans[(size_t)frame * 3 + 1], t2 = trans[(size_t)frame * 3 + 2];
  const float* Al = As + lane * SG_AS;
  const int g_end = min(m.num_groups, (int)(blockIdx.x + 1) * groups_per_block);
  for (int g = blockIdx.x * groups_per_block + warp; g < g_end; g += SG_WARPS) {
    float p[3 * SG_G], acc[3 * SG_G];
    const float4* src = reinterpret_cast<const float4*>(vposed + (size_t)frame * v3_ld + (size_t)g * (3 * SG_G));
#pragma unroll
    for (int q = 0; q < 3 * SG_G / 4; ++q) {
      const float4 v = src[q];
      p[4 * q] = v.x; p[4 * q + 1] = v.y; p[4 * q + 2] = v.z; p[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < 3 * SG_G; ++i) acc[i] = 0.f;
    const int e0 = m.g_start[g], e1 = m.g_start[g + 1];
    // one entry of look-ahead on the (warp-uniform) joint index and weight row: their L1/L2 latency hides under the FMAs
    int jn = 0;
    float4 wan = make_float4(0.f, 0.f, 0.f, 0.f), wbn = wan;
    if (e0 < e1) {
      jn = m.g_joint[e0];
      wan = __ldg(reinterpret_cast<const float4*>(m.g_w + (size_t)e0 * SG_G));
      wbn = __ldg(reinterpret_cast<const float4*>(m.g_w + (size_t)e0 * SG_G) + 1);
    }
    for (int e = e0; e < e1; ++e) {                            // joints of this group (warp-uniform trip count)
      const float4* a = reinterpret_cast<const float4*>(Al + jn);
      const float4 r0 = a[0], r1 = a[1], r2 = a[2];
      const float w[SG_G] = {wan.x, wan.y, wan.z, wan.w, wbn.x, wbn.y, wbn.z, wbn.w};
      if (e + 1 < e1) {
        jn = m.g_joint[e + 1];
        wan = __ldg(reinterpret_cast<const float4*>(m.g_w + (size_t)(e + 1) * SG_G));
        wbn = __ldg(reinterpret_cast<const float4*>(m.g_w + (size_t)(e + 1) * SG_G) + 1);
      }
#pragma unroll
      for (int i = 0; i < SG_G; ++i) {
        if (w[i] != 0.f) {                                     // warp-uniform: weights depend on the vertex only
          const float px = p[3 * i], py = p[3 * i + 1], pz = p[3 * i + 2];
          acc[3 * i] = fmaf(w[i], fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))), acc[3 * i]);
          acc[3 * i + 1] = fmaf(w[i], fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))), acc[3 * i + 1]);
          acc[3 * i + 2] = fmaf(w[i], fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))), acc[3 * i + 2]);
        }
      }
    }
    if (fok) {
      const int nvalid = min(SG_G, m.num_verts - g * SG_G) * 3;           // floats of this group inside the mesh (even)
      float2* dst = reinterpret_cast<float2*>(out + ((size_t)frame * m.num_verts + (size_t)g * SG_G) * 3);
#pragma unroll
      for (int q = 0; q < 3 * SG_G / 2; ++q) {
        // floats 2q, 2q+1 of the group: coordinate (2q) % 3 and (2q+1) % 3 of vertices (2q)/3 and (2q+1)/3
        const int c0 = (2 * q) % 3, c1 = (2 * q + 1) % 3;
        const float a0 = acc[2 * q] + (c0 == 0 ? t0 : (c0 == 1 ? t1 : t2));
        const float a1 = acc[2 * q + 1] + (c1 == 0 ? t0 : (c1 == 1 ? t1 : t2));
        if (2 * q + 1 < nvalid) dst[q] = make_float2(a0, a1);
        else if (2 * q < nvalid) out[((size_t)frame * m.num_verts + (size_t)g * SG_G) * 3 + 2 * q] = a0;
      }
    }
  }
}

}  // namespace hb
