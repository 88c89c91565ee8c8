// The sequential decoder chain of the HuMoR rollout as ONE persistent kernel per direction
// (reference loop: models/humor_model.py:870-1001 roll_out; :445-498 decode; :696-772 apply_world2local_trans).
//
// The launch-per-layer chain (rollout.cu: 4 GEMM + 1 glue launch per step and direction, x 59 steps) spent most of its
// 13-30 us per launch on launch latency, prologues (barrier init, TMEM allocation), an un-overlapped leader epilogue and the
// kernel tail.  Here 32 clusters of 4 CTAs stay resident for all S steps of one direction:
//   * every GEMM phase is tiled 128 rows x 64 columns; the 4 CTAs of a cluster split K (as umma_gemm3_kernel<64,.,4>: TMA ->
//     3-stage mbarrier ring -> tcgen05.mma kind::tf32, 3xTF32 split operands, K-chunks of 128 promoted from two ping-pong
//     TMEM buffers to fp32 registers);
//   * the split-K partials are REDUCE-SCATTERED through distributed shared memory: epilogue warp q of every CTA owns TMEM
//     lanes 32q..32q+31 and sends those 32 rows to CTA q of the cluster, so each CTA finalises 32 rows x 64 columns with all of
//     its 128 epilogue threads (4 threads per row, GroupNorm statistics by two shuffles) - the epilogue and the DSMEM
//     traffic are spread over 4 SMs instead of serialised on a leader; hand-shake by remote mbarrier arrivals (no cluster
//     barrier: the TMA and MMA warps never stall on the exchange and run ahead into the next tile);
//   * phases are ordered by DATA-FLOW FLAGS in global memory instead of grid barriers: a finished 32x64 slab bumps its tile's
//     counter (release), the TMA producer of a consuming tile polls exactly the counters of the k-blocks it is about to load
//     (acquire) - weights (B operand) are requested before that wait;
//   * the per-row glue (delta composition, canonicalisation, world transform; glue_warp.cuh) runs on the epilogue warps of
//     all CTAs, one warp per sub-sequence, between the last and the first GEMM phase of consecutive steps;
//   * reverse direction: the z-skip columns of the transposed GEMMs are NOT on the recurrence; every step's operand planes
//     (d raw | d pre3 | d pre2 | d pre1) are kept and d z for all steps comes from ONE batched GEMM afterwards (rollout.cu).
// Any number of resident clusters >= 1 is correct (tiles are strided over clusters; dependencies only point to earlier
// phases), so the kernel also runs on the CPU emulation of tests/host with two clusters.
#pragma once
#include "chain_args.cuh"
#include "glue_warp.cuh"
#include "umma_gemm.cuh"

namespace hb {

constexpr int CH_STAGES = 3;
constexpr int CH_A_TILE = UM_BM * 128;                         // bytes of one operand plane tile
constexpr int CH_B_TILE = CH_BN * 128;
constexpr int CH_STAGE = 2 * CH_A_TILE + 2 * CH_B_TILE;        // 48 KB: A_hi | A_lo | B_hi | B_lo
constexpr int CH_XROWS = UM_BM / CH_CS;                        // rows a CTA finalises
constexpr int CH_XLD = 68;                                     // floats per row of a partial slab (64 + pad: conflict-free v4 stores)
constexpr int CH_XBUF = CH_CS * CH_XROWS * CH_XLD * 4;         // one slab per source CTA
constexpr int CH_SLAB = CH_XROWS * CH_XLD * 4;                 // one 32-row partial slab: 8 704 bytes
constexpr int CH_GLUE_W = CH_SLAB / 4;                         // floats of an epilogue warp's staging area: its outgoing slab, and the glue's
static_assert(CH_GLUE_W >= GLUE_BWD_SMEM, "the glue's row arrays live in the slab staging area");   // row arrays between tiles
constexpr int CH_GLUE = 4 * CH_SLAB;
constexpr int CH_PF_W = 340 + 216 + 348 + 12 + 352;            // reverse glue: prefetched xin | raw | d world | G | d xin (prior) of the next step
constexpr int CH_PF = 2 * CH_PF_W * 4;                         // one set per warp pair
constexpr int CH_BARS = 256;
constexpr int CH_SMEM = CH_STAGES * CH_STAGE + CH_XBUF + CH_GLUE + CH_PF + CH_BARS + 1024 /*align slack*/;
static_assert(CH_SMEM <= 232448, "shared memory of one CTA");
constexpr int CHAIN_THREADS = 192;

struct alignas(64) ChainParams {
  CUtensorMap map_hi[CH_NMAPS];
  CUtensorMap map_lo[CH_NMAPS];
  ChainGemm g[CH_NGEMM];
  ChainGlue glue;
  unsigned* flags;
  int B, S, dir;
  int f16;                 // forward only: operand planes are fp16 hi / scaled-lo (k-blocks of 64 halves, kind::f16 MMAs, D1 | D2 accumulators)
  long long* dbg;          // optional clock64 stamps of CTA 0 (humor_chain_debug), else nullptr
};
// stamp event `ev` of (step u, phase ph: GEMM phases 0..3, glue 4) - one thread of CTA 0 only
#define CH_STAMP(u, ph, ev) do { if (p.dbg && blockIdx.x == 0) p.dbg[((u) * 5 + (ph)) * CH_DBG_EV + (ev)] = hb_clock64(); } while (0)

#ifdef HB_HOST_SHIM
static inline long long hb_clock64() { return 0; }
#else
__device__ __forceinline__ long long hb_clock64() { return clock64(); }
#endif

#ifdef HB_HOST_SHIM
using tcemu::cluster_id_x; using tcemu::cluster_nid_x; using tcemu::mbar_arrive_remote; using tcemu::mbar_wait_cluster;
using tcemu::flag_wait_ge; using tcemu::flag_add_release; using tcemu::fence_proxy_async; using tcemu::epi_bar_sync;
using tcemu::st_async_v4; using tcemu::fence_gpu; using tcemu::bulk_s2c;
#else
// 16-byte store into another CTA's shared memory that completes 16 transaction bytes on an mbarrier of THAT CTA when it lands:
// data and signal travel together, no release fence / separate arrival on the critical path
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_mbar) : "memory");
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// 1-D bulk copy from this CTA's shared memory into a peer's (or its own) through the cluster's shared-memory network, completing
// `bytes` on an mbarrier of the destination CTA: ONE copy per 8.7 KB slab instead of 512 st.async of 16 bytes
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
// arrive on an mbarrier of another CTA of the cluster (address from mapa); orders this thread's earlier DSMEM stores
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// data-flow flags in global memory: monotonic counters, bumped with release, polled with acquire (bounded: a protocol bug traps)
__device__ __forceinline__ void flag_wait_ge(const unsigned* p, unsigned target) {
  // (measured on the B200, profiles/r02d-e: polling with ld.relaxed + one fence.acq_rel afterwards hands over 0.2 us LATER
  // than acquire loads - the fence costs more than the per-iteration L1 invalidation saves)
  const long long t0 = clock64();
  while (true) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= target) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void flag_add_release(unsigned* p, unsigned v) {
  asm volatile("fence.acq_rel.gpu;\n\tred.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy global writes -> visible to (and ordered before) TMA reads issued after the flag hand-off
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
#endif

__device__ __forceinline__ unsigned* chain_tile_flag(unsigned* flags, int gi, int mt, int nt) {
  return flags + (gi * CH_MAX_MT + mt) * CH_MAX_NT + nt;
}
__device__ __forceinline__ unsigned* chain_glue_flag(unsigned* flags, int mt) { return flags + CH_NGEMM * CH_MAX_MT * CH_MAX_NT + mt; }

__global__ void __launch_bounds__(CHAIN_THREADS, 1)
chain_kernel(const __grid_constant__ ChainParams p) {
  HB_DYN_SMEM(smem_raw);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t xbuf = base + CH_STAGES * CH_STAGE;
  const uint32_t glue_s = xbuf + CH_XBUF;
  const uint32_t pf_s = glue_s + CH_GLUE;
  const uint32_t bars = pf_s + CH_PF;
  const uint32_t full0 = bars, empty0 = bars + 8 * CH_STAGES, tfull0 = bars + 16 * CH_STAGES, tempty0 = tfull0 + 16,
                 xfull = tempty0 + 16, xfree = xfull + 8, tptr = xfree + 8, gbar0 = tptr + 8;
  float* glue_f = reinterpret_cast<float*>(smem_raw + (glue_s - raw));
  float* pf_f = reinterpret_cast<float*>(smem_raw + (pf_s - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t krank = cluster_ctarank();
  const int cid = (int)cluster_id_x(), ncl = (int)cluster_nid_x();
  const int cta = blockIdx.x, nctas = gridDim.x;
  const int B = p.B, S = p.S, dir = p.dir;
  const int MT = (B + UM_BM - 1) / UM_BM;
  unsigned* const flags = p.flags;
  const bool f16 = p.f16 != 0;
  const int bk = f16 ? 64 : UM_BK;                             // operand elements per k-block (always 128-byte rows)
  const int chunk = f16 ? 2 : UM_CHUNK;                        // k-blocks (K = 128) accumulated in TMEM before promotion

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < CH_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
    mbar_init(xfull, 1);                   // the receiver's own arrive.expect_tx; the partial slabs arrive as transaction bytes
    mbar_init(xfree, CH_CS);               // one arrival per CTA of the cluster once its slab has been consumed
    for (int i = 0; i < 4; ++i) mbar_init(gbar0 + 8 * i, 1);   // reverse glue: tape rows of the next step prefetched into the staging arrays
    mbar_fence_init();
    mbar_expect_tx(xfull, CH_CS * CH_SLAB);                 // tile 0: four slabs of 32 rows x 68 floats
  }
  if (warp == 1) {
    tmem_alloc(tptr, (uint32_t)(4 * CH_BN));               // two ping-pong buffers x (D1 | D2): the fp16 form keeps the cross terms apart
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_u32(tptr);
  cluster_sync_all();                      // peers' barriers are initialised before anyone arrives on them remotely

  // rows of row tile mt (the last one may be ragged)
  auto rows_of = [&](int mt) { return min(UM_BM, B - mt * UM_BM); };
  // glue completions an A operand / a z tail written by the glue needs before step u of this direction
  auto glue_need = [&](int mt, int u) { return (unsigned)(rows_of(mt) * (dir ? u + 1 : u)); };

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = 0; u < S; ++u) {
        const int t = dir ? S - 1 - u : u;
        for (int gi = 0; gi < CH_NGEMM; ++gi) {
          const ChainGemm& g = p.g[gi];
          const int nkb_per = (g.nkb + CH_CS - 1) / CH_CS;
          const int kb0 = (int)krank * nkb_per, kb1 = min(g.nkb, kb0 + nkb_per);
          const CUtensorMap* a_hi = &p.map_hi[g.a_map]; const CUtensorMap* a_lo = &p.map_lo[g.a_map];
          const CUtensorMap* b_hi = &p.map_hi[g.b_map]; const CUtensorMap* b_lo = &p.map_lo[g.b_map];
          for (int tile = cid; tile < g.ntn * MT; tile += ncl) {
            const int mt = tile / g.ntn, nt = tile % g.ntn;
            const int m0 = mt * UM_BM, n0 = nt * CH_BN;
            int dep_ok = -2;                                    // producer tile already seen complete (-1: the glue)
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
              const int s = it % CH_STAGES;
              const uint32_t ph = (it / CH_STAGES) & 1;
              const uint32_t st = base + s * CH_STAGE;
              mbar_wait(empty0 + 8 * s, ph ^ 1);
              mbar_expect_tx(full0 + 8 * s, CH_STAGE);
              tma_load_2d(st + 2 * CH_A_TILE, b_hi, full0 + 8 * s, kb * bk, n0);               // weights: no dependency
              tma_load_2d(st + 2 * CH_A_TILE + CH_B_TILE, b_lo, full0 + 8 * s, kb * bk, n0);
              const int pt = (kb * bk) / CH_BN;                 // producer tile that holds this k-block's columns
              const int ptile = (gi > 0 && pt < g.dep_ntn) ? pt : -1;
              if (kb == kb0) CH_STAMP(u, gi, 0);
              if (ptile != dep_ok) {
                if (ptile >= 0) flag_wait_ge(chain_tile_flag(flags, gi - 1, mt, ptile), (unsigned)(CH_CS * (u + 1)));
                else { const unsigned need = glue_need(mt, u); if (need) flag_wait_ge(chain_glue_flag(flags, mt), need); }
                fence_proxy_async();
                dep_ok = ptile;
              }
              if (kb == kb0) CH_STAMP(u, gi, 1);
              if (kb == kb1 - 1) CH_STAMP(u, gi, 2);
              tma_load_2d(st, a_hi, full0 + 8 * s, g.a_col0 + kb * bk, t * g.a_row_step + m0);
              tma_load_2d(st + CH_A_TILE, a_lo, full0 + 8 * s, g.a_col0 + kb * bk, t * g.a_row_step + m0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CH_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      uint32_t it = 0, ck = 0;
      for (int u = 0; u < S; ++u)
        for (int gi = 0; gi < CH_NGEMM; ++gi) {
          const ChainGemm& g = p.g[gi];
          const int nkb_per = (g.nkb + CH_CS - 1) / CH_CS;
          const int kb0 = (int)krank * nkb_per, kb1 = min(g.nkb, kb0 + nkb_per);
          for (int tile = cid; tile < g.ntn * MT; tile += ncl) {
            for (int c0 = kb0; c0 < kb1; c0 += chunk, ++ck) {
              const int buf = ck & 1;
              mbar_wait(tempty0 + 8 * buf, ((ck >> 1) & 1) ^ 1);           // the epilogue has drained this TMEM buffer
              tc_fence_after();
              const uint32_t tacc = tmem_base + buf * 2 * CH_BN;            // D1 (and, fp16 form, D2 = the cross terms at + CH_BN)
              const int c1 = min(kb1, c0 + chunk);
              for (int kb = c0; kb < c1; ++kb, ++it) {
                const int s = it % CH_STAGES;
                mbar_wait(full0 + 8 * s, (it / CH_STAGES) & 1);
                tc_fence_after();
                if (kb == kb0) CH_STAMP(u, gi, 3);
                if (kb == kb1 - 1) CH_STAMP(u, gi, 4);
                const uint32_t st = base + s * CH_STAGE;
                if (!f16) {
#pragma unroll
                  for (int k = 0; k < UM_BK / 8; ++k) {
                    const uint64_t ah = umma_desc_sw128(st + k * 32);
                    const uint64_t al = umma_desc_sw128(st + CH_A_TILE + k * 32);
                    const uint64_t bh = umma_desc_sw128(st + 2 * CH_A_TILE + k * 32);
                    const uint64_t bl = umma_desc_sw128(st + 2 * CH_A_TILE + CH_B_TILE + k * 32);
                    umma_tf32(tacc, ah, bh, idesc, (kb != c0) || (k != 0));
                    umma_tf32(tacc, al, bh, idesc, 1);
                    umma_tf32(tacc, ah, bl, idesc, 1);
                  }
                } else {
                  // x = h + l 2^-11: h.h into D1, l.h + h.l into D2 (umma_gemm16.cuh); one UMMA consumes K = 16 halves = 32 bytes
                  constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(CH_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t ah = umma_desc_sw128(st + k * 32);
                    const uint64_t al = umma_desc_sw128(st + CH_A_TILE + k * 32);
                    const uint64_t bh = umma_desc_sw128(st + 2 * CH_A_TILE + k * 32);
                    const uint64_t bl = umma_desc_sw128(st + 2 * CH_A_TILE + CH_B_TILE + k * 32);
                    const uint32_t first = (kb != c0) || (k != 0);
                    umma_f16(tacc, ah, bh, idesc16, first);
                    umma_f16(tacc + CH_BN, al, bh, idesc16, first);
                    umma_f16(tacc + CH_BN, ah, bl, idesc16, 1);
                  }
                }
                umma_commit(empty0 + 8 * s);
              }
              umma_commit(tfull0 + 8 * buf);
            }
          }
        }
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue + glue warps
    const int q = warp & 3;                                    // TMEM lane quadrant of this warp = destination CTA of its rows
    const int ew = warp - 2;                                   // 0..3
    const int et = threadIdx.x - 64;                           // 0..127
    const int fr = et >> 2, cq = et & 3;                       // finalise: row inside this CTA's slab, 16-column quarter
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ck = 0, tl = 0;

    // Glue rows: with at least two epilogue warps per sub-sequence in the grid (B = 256 on 128 CTAs) a row is split over a warp
    // PAIR (glue_*_pair: joint rotations | everything else), else one warp takes a row.
    const bool pairs = 2 * nctas >= B;
    const int pi = ew >> 1, role = ew & 1;                      // pair of this warp, its role inside the pair
    const int b_first = pairs ? pi * nctas + cta : ew * nctas + cta;
    const int b_step = pairs ? 2 * nctas : 4 * nctas;
    const bool lead = lane == 0 && (!pairs || role == 0);     // the lane that waits / releases / prefetches for this warp's rows
    float* const gs = glue_f + (pairs ? 2 * pi : ew) * CH_GLUE_W;      // a pair shares the staging area of its first warp
    // Reverse glue: xin / raw / d world / G rows of a step come from the forward tape in HBM (~2 us of load latency on the
    // recurrence, measured: profiles/r02d).  With one row per warp pair and step they are known a whole step ahead: the lead lane
    // bulk-copies the NEXT step's rows into the pair's prefetch set as soon as this step's row is done.
    const bool prefetch = dir && pairs && b_first < B && b_first + b_step >= B;
    const uint32_t gbar = gbar0 + 8 * pi;
    float* const pf = pf_f + pi * CH_PF_W;
    const uint32_t pf_a = pf_s + (uint32_t)(pi * CH_PF_W) * 4u;
    auto prefetch_rows = [&](int t) {                            // lead lane only
      const ChainGlue& gl = p.glue;
      const size_t r = (size_t)t * B + b_first;
      fence_proxy_async();                                       // generic reads of the prefetch set -> async-proxy writes
      const bool px = t + 1 < S;                                 // the last step has no successor: no d xin row from the prior
      mbar_expect_tx(gbar, (CH_PF_W - (px ? 0 : 352)) * 4);
      bulk_g2s(pf_a, gl.xins + r * XIN_LD, 340 * 4, gbar);
      bulk_g2s(pf_a + 340 * 4, gl.raws + r * RAW_LD, RAW_D * 4, gbar);
      bulk_g2s(pf_a + 556 * 4, gl.dworld + r * WORLD_LD, WORLD_LD * 4, gbar);
      bulk_g2s(pf_a + 904 * 4, gl.Gs + r * 12, 12 * 4, gbar);
      if (px) bulk_g2s(pf_a + 916 * 4, gl.dpx + (r + B) * 352, 352 * 4, gbar);   // the batched prior's d xin of step t+1
    };
    if (prefetch && lead) prefetch_rows(S - 1);
    auto run_glue = [&](int u, int t) {
      const ChainGlue& gl = p.glue;
      const ChainGemm& gp = p.g[CH_NGEMM - 1];                  // the phase that feeds the glue in either direction
      for (int b = b_first; b < B; b += b_step) {
        const int mt = b / UM_BM;
        const unsigned need = (unsigned)(CH_CS * (dir ? u : u + 1));
        if (lead && ew == 0) CH_STAMP(u, 4, 0);
        if (lead && need) {
          for (int nt = 0; nt < gp.ntn; ++nt) flag_wait_ge(chain_tile_flag(flags, CH_NGEMM - 1, mt, nt), need);
        }
        if (lead && prefetch) mbar_wait(gbar, u & 1);           // landed long ago
        if (tl > 0) mbar_wait_cluster(xfree, (tl - 1) & 1);     // the staging area doubles as the glue's row arrays: every peer has
        if (pairs) glue_pair_sync(pi); else __syncwarp();       // consumed the last slab, so its bulk copy has read it
        if (lead && ew == 0) CH_STAMP(u, 4, 1);
        const size_t r = (size_t)t * B + b;
        if (!dir) {
          GlueFwdRow io;
          io.xr = gl.xins + r * XIN_LD; io.rr = gl.raws + r * RAW_LD; io.G = gl.Gs + r * 12; io.t2j = gl.t2j + b * 4;
          io.zt = (t + 1 < S) ? gl.z + ((size_t)b * S + (t + 1)) * 48 : nullptr;
          io.xn = gl.xins + (r + B) * XIN_LD; io.xn_hi = gl.xin_hi + (r + B) * XIN_LD; io.xn_lo = gl.xin_lo + (r + B) * XIN_LD;
          io.wo = gl.world + r * WORLD_LD; io.gn = gl.Gs + (r + B) * 12;
          io.h1 = gl.h1 + (size_t)b * 1088 + 1024; io.h1_lo = gl.h1_lo + (size_t)b * 1088 + 1024;
          io.h2 = gl.h2 + (size_t)b * 1088 + 1024; io.h2_lo = gl.h2_lo + (size_t)b * 1088 + 1024;
          io.h3 = gl.h3 + (size_t)b * 576 + 512; io.h3_lo = gl.h3_lo + (size_t)b * 576 + 512;
          io.xn16_h = io.xn16_l = io.h1_16h = io.h1_16l = io.h2_16h = io.h2_16l = io.h3_16h = io.h3_16l = nullptr;
          if (f16) {                                            // operand planes of the next step as fp16 hi / scaled lo; no fp32 planes
            io.xn_hi = io.xn_lo = nullptr;
            io.xn16_h = gl.x16_h + (r + B) * gl.x16_ld; io.xn16_l = gl.x16_l + (r + B) * gl.x16_ld;
            io.h1_16h = gl.h1_16h + (size_t)b * 1088 + 1024; io.h1_16l = gl.h1_16l + (size_t)b * 1088 + 1024;
            io.h2_16h = gl.h2_16h + (size_t)b * 1088 + 1024; io.h2_16l = gl.h2_16l + (size_t)b * 1088 + 1024;
            io.h3_16h = gl.h3_16h + (size_t)b * 576 + 512; io.h3_16l = gl.h3_16l + (size_t)b * 576 + 512;
          }
          if (pairs) glue_fwd_pair<true>(io, role, lane, pi, gs, gs + 340, gs + 556, gs + 896);
          else glue_fwd_warp<true>(io, lane, gs, gs + 340, gs + 556, gs + 896);
        } else {
          GlueBwdRow io;
          io.xr = gl.xins + r * XIN_LD; io.rr = gl.raws + r * RAW_LD; io.wr = gl.dworld + r * WORLD_LD;
          io.G = prefetch ? pf + 904 : gl.Gs + r * 12; io.staged = prefetch ? 1 : 0;
          io.t2j = gl.t2j + b * 4; io.have_next = u > 0;
          io.a0 = gl.da0 + (size_t)b * XIN_LD; io.px = prefetch ? pf + 916 : gl.dpx + (r + B) * 352; io.xs = gl.dxres + (size_t)b * 340;
          io.dGn = ((t + 1) & 1 ? gl.dG1 : gl.dG0) + (size_t)b * 12; io.dG = (t & 1 ? gl.dG1 : gl.dG0) + (size_t)b * 12;
          io.dt2j = gl.dt2j + b * 4; io.dzt = nullptr;
          io.dh1 = io.dh1_lo = io.dh2 = io.dh2_lo = io.dh3 = io.dh3_lo = nullptr;
          io.draw = nullptr; io.draw_hi = gl.bp_hi + r * gl.bp_ld; io.draw_lo = gl.bp_lo + r * gl.bp_ld;
          if (pairs) glue_bwd_pair<true>(io, role, lane, pi, prefetch ? pf : gs, prefetch ? pf + 340 : gs + 340, gs + 556, prefetch ? pf + 556 : gs + 896,
                                         gs + 1244, gs + 1584);
          else glue_bwd_warp<true>(io, lane, gs, gs + 340, gs + 556, gs + 896, gs + 1244, gs + 1584);
        }
        if (lead && ew == 0) CH_STAMP(u, 4, 2);
        fence_proxy_async();                                    // this lane's rows -> TMA reads of the consuming GEMM phase
        if (pairs) glue_pair_sync(pi); else __syncwarp();
        if (lead) flag_add_release(chain_glue_flag(flags, mt), 1u);
        if (lead && ew == 0) CH_STAMP(u, 4, 3);
        if (lead && prefetch && u + 1 < S) prefetch_rows(t - 1);
      }
    };

    for (int u = 0; u < S; ++u) {
      const int t = dir ? S - 1 - u : u;
      if (dir) run_glue(u, t);
      for (int gi = 0; gi < CH_NGEMM; ++gi) {
        const ChainGemm& g = p.g[gi];
        const int nkb_per = (g.nkb + CH_CS - 1) / CH_CS;
        const int kb0 = (int)krank * nkb_per, kb1 = min(g.nkb, kb0 + nkb_per);
        for (int tile = cid; tile < g.ntn * MT; tile += ncl, ++tl) {
          const int mt = tile / g.ntn, nt = tile % g.ntn;
          const int m0 = mt * UM_BM, n0 = nt * CH_BN;
          float acc[CH_BN];
#pragma unroll
          for (int j = 0; j < CH_BN; ++j) acc[j] = 0.f;
          for (int c0 = kb0; c0 < kb1; c0 += chunk, ++ck) {
            const int buf = ck & 1;
            mbar_wait(tfull0 + 8 * buf, (ck >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < CH_BN; cc += 32) {
              float tv[32];
              tmem_ld32(trow + buf * 2 * CH_BN + cc, tv);
              if (f16) {                                          // + 2^-11 x the cross terms
                float tw[32];
                tmem_ld32(trow + buf * 2 * CH_BN + CH_BN + cc, tw);
#pragma unroll
                for (int j = 0; j < 32; ++j) tv[j] = fmaf(tw[j], 0.00048828125f, tv[j]);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[cc + j] += tv[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
          }
          // ---- reduce-scatter of the split-K partials: my quadrant's 32 rows go to CTA q
          if (et == 0) CH_STAMP(u, gi, 5);
          if (tl > 0) mbar_wait_cluster(xfree, (tl - 1) & 1);   // every CTA of the cluster has consumed its previous slab
          if (et == 0) CH_STAMP(u, gi, 6);
          {
            // park the warp's 32 x 64 block in its staging area (lane = row, conflict-free 272-byte stride), then ONE bulk copy
            // carries the slab to CTA q and completes its bytes on that CTA's barrier
            const uint32_t stg = glue_s + (uint32_t)ew * CH_SLAB;
#pragma unroll
            for (int j = 0; j < CH_BN; j += 4) st_shared_v4(stg + (uint32_t)lane * (CH_XLD * 4) + j * 4, acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            fence_proxy_async();                                // generic writes of the slab -> the bulk copy's reads
            __syncwarp();
            if (lane == 0) bulk_s2c(map_to_cta(xbuf + (uint32_t)krank * CH_SLAB, (uint32_t)q), stg, CH_SLAB, map_to_cta(xfull, (uint32_t)q));
          }
          if (et == 0) CH_STAMP(u, gi, 7);
          // ---- finalise rows krank*32 .. +31 of the tile: 4 threads per row, 16 columns each.  Everything the epilogue needs from
          // global memory is requested NOW, under the exchange
          const int row = m0 + (int)krank * CH_XROWS + fr;
          const bool rok = row < B;
          const int col = n0 + cq * 16;
          const size_t trw = (size_t)t * B + row;               // row of the per-step tapes
          float cb[16], cg[16], ce[16], xh[16];                 // bias | gamma | beta | saved x-hat (reverse)
          float rs_saved = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { cb[j] = 0.f; cg[j] = 0.f; ce[j] = 0.f; xh[j] = 0.f; }
          if (g.epi == EPI_BIAS) {
            if (g.bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j) cb[j] = (col + j < g.N) ? g.bias[col + j] : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 g4 = *reinterpret_cast<const float4*>(g.gamma + col + j);
              const float4 e4 = *reinterpret_cast<const float4*>(g.beta + col + j);
              cg[j] = g4.x; cg[j + 1] = g4.y; cg[j + 2] = g4.z; cg[j + 3] = g4.w;
              ce[j] = e4.x; ce[j + 1] = e4.y; ce[j + 2] = e4.z; ce[j + 3] = e4.w;
              if (g.epi == EPI_GN_RELU) {
                const float4 b4 = *reinterpret_cast<const float4*>(g.bias + col + j);
                cb[j] = b4.x; cb[j + 1] = b4.y; cb[j + 2] = b4.z; cb[j + 3] = b4.w;
              } else if (rok) {
                const float4 x4 = *reinterpret_cast<const float4*>(g.xhat + trw * g.ldxh + col + j);
                xh[j] = x4.x; xh[j + 1] = x4.y; xh[j + 2] = x4.z; xh[j + 3] = x4.w;
              }
            }
            if (g.epi == EPI_GN_RELU_BWD && rok) rs_saved = g.rstd[trw * 16 + col / g.gsize];
          }
          mbar_wait_cluster(xfull, tl & 1);
          if (et == 0) {
            mbar_expect_tx(xfull, CH_CS * CH_SLAB);                // arm the next tile (peers send it only after xfree below)
            CH_STAMP(u, gi, 8);
          }
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll
          for (int src = 0; src < CH_CS; ++src) {
            const uint32_t a = xbuf + (uint32_t)((src * CH_XROWS + fr) * CH_XLD + cq * 16) * 4u;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 x = ld_shared_v4(a + j * 4);
              v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
            }
          }
          const int gl_lanes = (g.gsize == 64) ? 3 : 1;         // xor-shuffle masks that span one GroupNorm group
          auto group_sum = [&](float x) {
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            if (gl_lanes == 3) x += __shfl_xor_sync(0xffffffffu, x, 2);
            return x;
          };
          if (g.epi == EPI_BIAS) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += cb[j];
          } else if (g.epi == EPI_GN_RELU) {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] += cb[j]; sum += v[j]; }
            const float inv = 1.f / (float)g.gsize;
            const float mean = group_sum(sum) * inv;
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float d = v[j] - mean; sq += d * d; }
            const float rs = rsqrtf(group_sum(sq) * inv + 1e-5f);
            if (rok && (cq & gl_lanes) == 0) g.rstd[trw * 16 + col / g.gsize] = rs;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 xh;
              float* xp = &xh.x;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                xp[e] = (v[j + e] - mean) * rs;
                v[j + e] = fmaxf(fmaf(cg[j + e], xp[e], ce[j + e]), 0.f);
              }
              if (rok) *reinterpret_cast<float4*>(g.xhat + trw * g.ldxh + col + j) = xh;
            }
          } else {                                              // EPI_GN_RELU_BWD: every column of these tiles is normalised
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float gm = cg[j];
              const bool on = fmaf(gm, xh[j], ce[j]) > 0.f;
              const float uu = on ? gm * v[j] : 0.f;
              v[j] = uu;
              s1 += uu;
              s2 += uu * xh[j];
            }
            s1 = group_sum(s1); s2 = group_sum(s2);
            const float rs = rs_saved;
            const float inv = 1.f / (float)g.gsize;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = rs * (v[j] - s1 * inv - xh[j] * s2 * inv);
          }
          if (rok && g.C16_h) {                                 // fp16 hi / scaled-lo planes of the next layer: 32 bytes per plane and thread
            unsigned short hh[16], ll[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split16(v[j], hh[j], ll[j]);
            uint4* dh = reinterpret_cast<uint4*>(g.C16_h + (size_t)row * g.ld16 + col);
            uint4* dl = reinterpret_cast<uint4*>(g.C16_l + (size_t)row * g.ld16 + col);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const unsigned short* a = hh + 8 * j; const unsigned short* b = ll + 8 * j;
              dh[j] = make_uint4(a[0] | (a[1] << 16), a[2] | (a[3] << 16), a[4] | (a[5] << 16), a[6] | (a[7] << 16));
              dl[j] = make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4] | (b[5] << 16), b[6] | (b[7] << 16));
            }
          }
          if (rok) {
            const size_t crow = (size_t)t * g.c_row_step + row;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int c = col + j;
              if (c + 3 < g.N) {
                if (g.C) *reinterpret_cast<float4*>(g.C + crow * g.ldc + g.c_col0 + c) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (g.C_hi) {
                  const float4 h = make_float4(tf32_hi(v[j]), tf32_hi(v[j + 1]), tf32_hi(v[j + 2]), tf32_hi(v[j + 3]));
                  *reinterpret_cast<float4*>(g.C_hi + crow * g.ldc + g.c_col0 + c) = h;
                  *reinterpret_cast<float4*>(g.C_lo + crow * g.ldc + g.c_col0 + c) = make_float4(v[j] - h.x, v[j + 1] - h.y, v[j + 2] - h.z, v[j + 3] - h.w);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (c + e < g.N) {
                    if (g.C) g.C[crow * g.ldc + g.c_col0 + c + e] = v[j + e];
                    if (g.C_hi) { const float h = tf32_hi(v[j + e]); g.C_hi[crow * g.ldc + g.c_col0 + c + e] = h; g.C_lo[crow * g.ldc + g.c_col0 + c + e] = v[j + e] - h; }
                  }
              }
            }
          }
          if (et == 0) CH_STAMP(u, gi, 9);
          fence_proxy_async();                                  // the slab -> TMA reads of the next phase
          fence_gpu();                                          // every writer waits for ITS stores (in parallel), not one thread for all
          epi_bar_sync();                                       // all 128 epilogue threads: slab read and written
          if (et == 0) {
            CH_STAMP(u, gi, 10);
            flag_add_release(chain_tile_flag(flags, gi, mt, nt), 1u);     // consumers first ...
            CH_STAMP(u, gi, 11);
#pragma unroll
            for (uint32_t rk = 0; rk < (uint32_t)CH_CS; ++rk) mbar_arrive_remote(map_to_cta(xfree, rk));   // ... then the peers' next slab
          }
        }
      }
      if (!dir) run_glue(u, t);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // no CTA leaves while a peer may still write to its shared memory / barriers
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)(4 * CH_BN));
}

}  // namespace hb
