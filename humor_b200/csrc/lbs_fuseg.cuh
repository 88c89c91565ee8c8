// Dense SMPL+H forward, skin form 3: blend GEMM + skinning in ONE persistent tcgen05 kernel with a lane = frame epilogue.
//
//   v_posed[frame, 3v+d] = v_template[3v+d] + feat[frame, :224] . blend_t[3v+d, :224]      (blend form 1: 3xTF32; form 5: fp16 hi/lo, see below)
//   out[frame, v, :]     = sum_j W[v][j] (A[frame, j] . [v_posed[frame, v]; 1]) + trans[frame]
//
// What the two measured predecessors taught (profiles/r01g):
//   * lbs_skin_apply_kernel (lane = vertex) gathers <= 4 transforms of 48 B per (vertex, frame) from shared memory - the
//     shared-memory port is its bound; a thread = frame kernel walking vertex by vertex (round 1, retired) re-loaded transforms
//     from global memory whenever the joint of a weight slot changed.
//   * the TMEM accumulator layout IS lane = frame (M = frames): a thread that reads its row with tcgen05.ld holds
//     consecutive vertices of ONE frame, which is exactly the operand of group skinning (vertices sorted by joint set): a
//     transform is fetched once per (frame, joint, group of 8 vertices) and joint index / weights are warp-uniform.
// Plan of one CTA (one per SM, 608 threads):
//   tiles   128 frames x 192 columns (= 64 vertices = 8 groups), walked ROW-major in one contiguous chunk per CTA: ~88
//           consecutive column tiles of the same 128 frames, so the frames' transforms stay on the SM
//   warp 0  TMA producer of the operand ring: 2 entries of 40 KB = one A plane (128 x 32 floats) + one B plane (192 x 32);
//           a k-block takes two entries (hi planes, lo planes); gated by the ring only, L2 evict_last hint (every CTA re-reads
//           the planes while 1.3 GB of output stream through the L2 evict-first)
//   warp 1  tcgen05.mma issuer; a tile's k-blocks accumulate into ONE of two 192-column TMEM buffers (K = 224: no
//           promotion chunks needed), so tile i+1's MMAs run under tile i's epilogue
//   warp 18 TMA producer, gated by the epilogue's progress, of (a) the 3x4 transforms of the joints the tile is skinned to, as
//           [128 frames][12 floats] boxes cut from A[N][52*12] into 13 shared-memory slots under the host's static schedule
//           (body_model.fuseg_tables): a joint keeps its slot while consecutive tiles need it, so a tile loads ~1 new slot (6 KB)
//           instead of ~6; (b) the tile's skinning RECORD (group offsets, per entry slot | joint | 8 weights: HbLbsModel.ft_rec)
//           with one bulk copy into one of two record buffers - the epilogue reads joint lists and weights at shared-memory
//           latency (read straight from global memory they missed the 28 KB of L1 this kernel leaves 3 times out of 4 and made
//           up 41 % of all stall cycles: profiles/r02g_fuseg35_set_full_details.txt); (c) with LbsFusegArgs.vs (one shape per
//           >= 32 frames) the rows of the shaped template [sequences][3V] that the tile's frames belong to (<= 5 x 768 B)
//   warps 2..17 epilogue: TMEM lane quadrant q = warp % 4, column quarter (warp - 2) / 4 -> 2 groups each.  Per group:
//           tcgen05.ld 24 columns (8 vertices of the thread's frame; the template rides in the GEMM, column 205 of the planes,
//           and the 2^-10 scale-back / the root translation sit in the transforms the pose kernel wrote for this pass), skin
//           with the group's joint list (3 x LDS.128 per joint from the slot, conflict-free: 48-byte frame stride), park
//           the 24 floats in a per-warp staging tile (6 x STS.128, chunks rotated by one for rows 4..7 mod 8: conflict-
//           free at a dense 96-byte row) and write 2 2/3 frame rows of 96 bytes per store instruction with all 32 lanes,
//           reading the tile as contiguous 256-byte pieces (a lane = frame store would touch 32 different lines per
//           instruction).  With LbsFusegArgs.vs the shaped template of the frame's sequence is added to the accumulators first.
//   fp16x3  (blend form 5, the default) EVERY column as fp16 hi + lo planes, three products per k-block (h.h + l.h + h.l) into the one accumulator:
//           the lo planes are UNSCALED (l = fp16(x - h)); what that costs is an absolute floor of 3e-8 on tiny operands, i.e.
//           ~1e-9 m after the 2^-10 scale-back - irrelevant here, and it keeps one accumulator per tile (a scaled lo part would
//           need a second one: 768 TMEM columns).  fp32-level accuracy (as three TF32 passes) at 320 instead of 560 KB per tile.
// Barrier protocol (all mbarriers, phases counted per use):
//   full[s]/empty[s]   operand ring (TMA complete_tx / tcgen05.commit)
//   tfull[b]/tempty[b] TMEM buffer b = tile parity (tcgen05.commit / one arrive per epilogue warp)
//   ttf[b]             transforms + record of tile parity b have landed (arrive.expect_tx by the producer, complete_tx by TMA).
//                      The producer issues tile i's transform loads only after tempty says the epilogue is done with tile
//                      i-2 (slots tile i-1 uses are never chosen by the schedule), and after tile i-1 when the frames change
//                      (then every slot is reloaded).
// Same TMA / UMMA descriptor forms and TMEM protocol as umma_gemm3_kernel; the default dense forward since round 2 (measured on
// the B200: profiles/r02g_* -> r03g_*; what bounds it now - the SM's L1 / shared-memory data pipe - is in DESIGN.md 4.2).  Executed on the CPU through tests/host/shim/tc_emul.h (tests/test_host_tc.py).
#pragma once
#include "umma_gemm.cuh"
#include "umma_launch.cuh"

namespace hb {

constexpr int FG_BN = 192;                          // columns per tile = 64 vertices
constexpr int FG_GPT = 8;                           // vertex groups per tile
constexpr int FG_G = 8;                             // vertices per group
constexpr int FG_GC = 3 * FG_G;                     // columns per group
constexpr int FG_RING = 2;                          // operand entries (40 KB: hi or lo planes of one k-block) in flight; a third entry bought nothing
                                                    // (840 vs 842 us, profiles/r03d vs r03c): the kernel is bound by L2 <-> SM bytes, not by load latency
constexpr int FG_A_PLANE = UM_BM * 128;             // bytes: 128 rows x 128 B
constexpr int FG_B_PLANE = FG_BN * 128;
constexpr int FG_ENTRY = FG_A_PLANE + FG_B_PLANE;   // 40 KB
constexpr int FG_NSLOT = 13;                        // body_model.FG_NSLOT
constexpr int FG_SLOT = UM_BM * 48;                 // [128 frames][12 floats]
constexpr int FG_EPI_WARPS = 16;                    // 4 per TMEM lane quadrant: two vertex groups of the tile each
constexpr int FG_STAGE_W = 32 * FG_GC * 4;          // bytes per epilogue warp: 32 frame rows x 96 B, dense
constexpr int FG_REC_HEAD = 64;                     // body_model.FG_REC_*: 9 group offsets + padding,
constexpr int FG_REC_ENTRY = 48;                    //   then { slot byte offset | joint * 12 | 0 | 0 | 8 weights } per entry
constexpr int FG_REC_MAX = FG_REC_HEAD + FG_REC_ENTRY * 96;
constexpr int FG_OFF_SLOTS = FG_RING * FG_ENTRY;
constexpr int FG_OFF_STAGE = FG_OFF_SLOTS + FG_NSLOT * FG_SLOT;
constexpr int FG_OFF_REC = FG_OFF_STAGE + FG_EPI_WARPS * FG_STAGE_W;
constexpr int FG_VS_ROWS = 5;                       // shaped-template rows (sequences) a 128-frame tile can span: frames_per_beta >= 32
constexpr int FG_VS_ROW = FG_BN * 4;                // bytes of one row inside a column tile
constexpr int FG_VS_BUF = FG_VS_ROWS * FG_VS_ROW;
constexpr int FG_OFF_VS = FG_OFF_REC + 2 * FG_REC_MAX;
constexpr int FG_OFF_BARS = FG_OFF_VS + 2 * FG_VS_BUF;
constexpr int FG_SMEM = FG_OFF_BARS + 128 + 1024;   // + barriers + 1024-byte alignment slack = 229 120 B
constexpr int FG_THREADS = 96 + 32 * FG_EPI_WARPS;   // + TMA warp (operands), MMA warp, TMA warp (transforms + records)
constexpr int FG_TAB = 4 + 2 * FG_NSLOT;            // ints per column tile of ft_tab
static_assert(FG_SMEM <= 232448, "lbs_fuseg_kernel: shared memory");

#ifndef HB_HOST_SHIM
// tcgen05.ld of one vertex group: 24 consecutive columns of the thread's TMEM lane, as three naturally aligned 8-column loads
// (group offsets are multiples of 24, i.e. of 8 but not of 16)
__device__ __forceinline__ void tmem_ld24(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[8 * c]), "=r"(r[8 * c + 1]), "=r"(r[8 * c + 2]), "=r"(r[8 * c + 3]), "=r"(r[8 * c + 4]), "=r"(r[8 * c + 5]),
                   "=r"(r[8 * c + 6]), "=r"(r[8 * c + 7])
                 : "r"(taddr + 8u * c) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// tcgen05.ld of half a vertex group: 12 consecutive columns as three 4-column loads (half-group offsets are multiples of 12)
__device__ __forceinline__ void tmem_ld12(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[4 * c]), "=r"(r[4 * c + 1]), "=r"(r[4 * c + 2]), "=r"(r[4 * c + 3]) : "r"(taddr + 4u * c) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void stcs2(float* p, float x, float y) { __stcs(reinterpret_cast<float2*>(p), make_float2(x, y)); }
__device__ __forceinline__ void stcs1(float* p, float x) { __stcs(p, x); }
#define HB_EMU_GUARD_ACQ(addr, bytes)
#define HB_EMU_GUARD_REL(addr)
#define HB_OPAQUE(x) asm volatile("" : "+r"(x))      // the value lives in a register from here on: not rematerialised
#else
using tcemu::tmem_ld24; using tcemu::tmem_ld12;
static inline void stcs2(float* p, float x, float y) { p[0] = x; p[1] = y; }
static inline void stcs1(float* p, float x) { p[0] = x; }
// tests/host: tell the emulation which shared-memory ranges are being read, so that a TMA write into them aborts
#define HB_EMU_GUARD_ACQ(addr, bytes) tcemu::guard_acquire(addr, bytes)
#define HB_EMU_GUARD_REL(addr) tcemu::guard_release(addr)
#define HB_OPAQUE(x)
#endif

__global__ void __launch_bounds__(FG_THREADS, 1)
lbs_fuseg_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmA16,
                 const __grid_constant__ CUtensorMap tmB16, const __grid_constant__ CUtensorMap tmA16l,
                 const __grid_constant__ CUtensorMap tmB16l, int K, LbsFusegArgs a) {
  HB_DYN_SMEM(smem_raw);
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + FG_OFF_BARS;
  const uint32_t full0 = bars, empty0 = bars + 8 * FG_RING, tfull0 = empty0 + 8 * FG_RING, tempty0 = tfull0 + 16, ttf0 = tempty0 + 16,
                 tptr = ttf0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = a.nrt * a.nct;
  const int nkb = K / UM_BK;
  // one contiguous chunk of the row-major tile list per CTA (column tile fastest: the frames change at most twice per chunk)
  const int t_begin = (int)((long long)ntiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < FG_RING; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, FG_EPI_WARPS); mbar_init(ttf0 + 8 * b, 1); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 512u);                                     // 2 x 192 columns (allocations are powers of two)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ld_shared_u32(tptr);

  if (warp == 0) {
    if (lane == 0) {
      // ---- operand planes: gated by the ring only, so the first k-blocks of tile i+1 land while tile i is still being multiplied
      const uint64_t keep = l2_policy_evict_last();             // every CTA re-reads the planes; the output streams through the L2
      int g = 0;                                                // ring entries issued
      for (int t = t_begin; t < t_end; ++t) {
        const int r = t / a.nct, c = t - r * a.nct;
        const int m0 = r * UM_BM, n0 = c * FG_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          for (int pl = 0; pl < 2; ++pl, ++g) {                 // hi entry, then lo entry
            const int s = g % FG_RING;
            mbar_wait(empty0 + 8 * s, ((g / FG_RING) & 1) ^ 1);
            const uint32_t st = base + s * FG_ENTRY;
            mbar_expect_tx(full0 + 8 * s, FG_ENTRY);
            tma_load_2d_hint(st, pl ? &tmA_lo : &tmA_hi, full0 + 8 * s, kb * UM_BK, m0, keep);
            tma_load_2d_hint(st + FG_A_PLANE, pl ? &tmB_lo : &tmB_hi, full0 + 8 * s, kb * UM_BK, n0, keep);
          }
        }
        for (int kb = 0; kb < a.nkb16; ++kb) {                  // fp16 k-blocks: 64 halves = the same 128-byte rows, same entry
          for (int pl = 0; pl < 2; ++pl, ++g) {                 // hi entry, then the lo entry
            const int s = g % FG_RING;
            mbar_wait(empty0 + 8 * s, ((g / FG_RING) & 1) ^ 1);
            const uint32_t st = base + s * FG_ENTRY;
            mbar_expect_tx(full0 + 8 * s, FG_ENTRY);
            if (a.dbg & 4) {
              tma_load_2d(st, pl ? &tmA16l : &tmA16, full0 + 8 * s, kb * 64, m0);
              tma_load_2d(st + FG_A_PLANE, pl ? &tmB16l : &tmB16, full0 + 8 * s, kb * 64, n0);
            } else {
              tma_load_2d_hint(st, pl ? &tmA16l : &tmA16, full0 + 8 * s, kb * 64, m0, keep);
              tma_load_2d_hint(st + FG_A_PLANE, pl ? &tmB16l : &tmB16, full0 + 8 * s, kb * 64, n0, keep);
            }
          }
        }
      }
    }
  } else if (warp == FG_EPI_WARPS + 2) {
    if (lane == 0) {
      // ---- skinning transforms of a tile's joints -> shared-memory slots, and the tile's record -> record buffer tc & 1: gated by
      // the epilogue's progress (tempty), on a warp of their own so that the operand stream above never waits for them
      int tc = 0, prev_r = -1;                                  // tiles started, row tile of the previous tile
      for (int t = t_begin; t < t_end; ++t, ++tc) {
        const int r = t / a.nct, c = t - r * a.nct;
        const int m0 = r * UM_BM;
        const bool fresh = r != prev_r;                         // first tile of the CTA, or other frames: reload every slot
        prev_r = r;
        if (fresh && tc >= 1) mbar_wait(tempty0 + 8 * ((tc - 1) & 1), ((tc - 1) >> 1) & 1);   // epilogue done with tile tc-1
        if (tc >= 2) mbar_wait(tempty0 + 8 * (tc & 1), ((tc >> 1) & 1) ^ 1);                   // ... with tile tc-2
        const int* tab = a.ft_tab + (size_t)c * FG_TAB;
        const int nl = fresh ? tab[0] : tab[1];
        const int* ent = tab + 4 + (fresh ? 0 : FG_NSLOT);
        const uint32_t tb = ttf0 + 8 * (tc & 1);
        const uint32_t recb = (uint32_t)tab[2];
        // shaped-template rows of the sequences this tile's frames belong to (a.vs; see LbsFusegArgs)
        int s0 = 0, nvs = 0;
        if (a.vs) {
          s0 = m0 / a.fpb;
          nvs = min(m0 + UM_BM - 1, a.N - 1) / a.fpb - s0 + 1;
        }
        mbar_expect_tx(tb, (uint32_t)nl * FG_SLOT + recb + (uint32_t)nvs * FG_VS_ROW);
        bulk_g2s(base + FG_OFF_REC + (uint32_t)(tc & 1) * FG_REC_MAX, a.ft_rec + (size_t)c * a.ft_rec_stride, recb, tb);
        for (int i = 0; i < nvs; ++i)
          bulk_g2s(base + FG_OFF_VS + (uint32_t)(tc & 1) * FG_VS_BUF + (uint32_t)i * FG_VS_ROW, a.vs + (size_t)(s0 + i) * a.vs_ld + (size_t)c * FG_BN,
                   FG_VS_ROW, tb);
        for (int i = 0; i < nl; ++i) {
          const int e = ent[i];
          tma_load_2d(base + FG_OFF_SLOTS + (uint32_t)(e >> 16) * FG_SLOT, &tmT, tb, e & 0xffff, m0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FG_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      int g = 0, tc = 0;
      for (int t = t_begin; t < t_end; ++t, ++tc) {
        const int buf = tc & 1;
        mbar_wait(tempty0 + 8 * buf, ((tc >> 1) & 1) ^ 1);      // epilogue has drained this TMEM buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * FG_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          const int gh = g++;
          const int sh = gh % FG_RING;
          mbar_wait(full0 + 8 * sh, (gh / FG_RING) & 1);
          tc_fence_after();
          const uint32_t sth = base + sh * FG_ENTRY;
#pragma unroll
          for (int k = 0; k < UM_BK / 8; ++k)                   // hi . hi
            umma_tf32(tacc, umma_desc_sw128(sth + k * 32), umma_desc_sw128(sth + FG_A_PLANE + k * 32), idesc, (kb != 0) || (k != 0));
          const int gl = g++;
          const int sl = gl % FG_RING;
          mbar_wait(full0 + 8 * sl, (gl / FG_RING) & 1);
          tc_fence_after();
          const uint32_t stl = base + sl * FG_ENTRY;
#pragma unroll
          for (int k = 0; k < UM_BK / 8; ++k) {                 // lo . hi + hi . lo
            umma_tf32(tacc, umma_desc_sw128(stl + k * 32), umma_desc_sw128(sth + FG_A_PLANE + k * 32), idesc, 1);
            umma_tf32(tacc, umma_desc_sw128(sth + k * 32), umma_desc_sw128(stl + FG_A_PLANE + k * 32), idesc, 1);
          }
          umma_commit(empty0 + 8 * sh);
          umma_commit(empty0 + 8 * sl);
        }
        for (int kb = 0; kb < a.nkb16; ++kb) {                  // fp16 planes: 4 MMAs of K = 16 per product and 64-wide k-block
          constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(FG_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
          const int gh = g++;
          const int sh = gh % FG_RING;
          mbar_wait(full0 + 8 * sh, (gh / FG_RING) & 1);
          tc_fence_after();
          const uint32_t sth = base + sh * FG_ENTRY;
#pragma unroll
          for (int k = 0; k < 4; ++k)                           // h . h (the tile's very first MMA overwrites the accumulator)
            umma_f16(tacc, umma_desc_sw128(sth + k * 32), umma_desc_sw128(sth + FG_A_PLANE + k * 32), idesc16, (nkb != 0) || (kb != 0) || (k != 0));
          const int gl = g++;                                   // + l . h + h . l: x = h + l with UNSCALED fp16 lo planes, one accumulator
          const int sl = gl % FG_RING;
          mbar_wait(full0 + 8 * sl, (gl / FG_RING) & 1);
          tc_fence_after();
          const uint32_t stl = base + sl * FG_ENTRY;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(tacc, umma_desc_sw128(stl + k * 32), umma_desc_sw128(sth + FG_A_PLANE + k * 32), idesc16, 1);
            umma_f16(tacc, umma_desc_sw128(sth + k * 32), umma_desc_sw128(stl + FG_A_PLANE + k * 32), idesc16, 1);
          }
          umma_commit(empty0 + 8 * sh);
          umma_commit(empty0 + 8 * sl);
        }
        umma_commit(tfull0 + 8 * buf);
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;                                     // TMEM lane quadrant of this warp
    const int h4 = ew >> 2;                                     // column quarter: groups 2*h4, 2*h4+1 of the tile
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = base + FG_OFF_STAGE + (uint32_t)ew * FG_STAGE_W;             // this warp's staging tile: [32 frames][96 B]
    const uint32_t tsl = base + FG_OFF_SLOTS + (uint32_t)(q * 32 + lane) * 48u;       // this thread's frame inside a slot
    // staging tile, written lane = frame: the six 16-byte chunks of row r sit at ((chunk + ((r >> 2) & 1)) % 6) * 16, which
    // makes eight consecutive rows of a dense 96-byte stride hit eight different bank quads
    const uint32_t rot = (uint32_t)(lane >> 2) & 1u;
    const uint32_t sw0 = stg + (uint32_t)lane * 96u + rot * 16u;                      // chunks 0..4 at sw0 + 16 k
    const uint32_t sw5 = stg + (uint32_t)lane * 96u + (rot ? 0u : 80u);               // chunk 5
    // store phase: a lane reads float2 number P = 32 it + lane of the tile's PHYSICAL 32 x 12 (contiguous 256 bytes per instruction:
    // conflict-free), which is row P / 12, chunk ((P % 12) / 2 - rot(row)) mod 6 of the group; three iterations cover eight rows
    // exactly, so a lane needs three (shared offset, global offset) pairs and adds 8 rows per round
    const int rowf = a.num_verts * 3;                           // floats per output frame
    uint32_t sr3[3];
    int go3[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int P3 = 32 * i + lane, r3 = P3 / 12, pos = P3 - 12 * r3;
      int k = (pos >> 1) - ((r3 >> 2) & 1);
      if (k < 0) k += 6;
      sr3[i] = stg + (uint32_t)P3 * 8u;
      go3[i] = r3 * rowf + 4 * k + 2 * (pos & 1);
    }
    // (the compiler otherwise re-derives these shared-memory addresses from %cgaid inside the joint loop: an S2R on the critical
    // path of every transform load, profiles/r03f)
    uint32_t tsl_ = tsl, sw0_ = sw0, sw5_ = sw5;
    HB_OPAQUE(tsl_); HB_OPAQUE(sw0_); HB_OPAQUE(sw5_); HB_OPAQUE(sr3[0]); HB_OPAQUE(sr3[1]); HB_OPAQUE(sr3[2]);
    int tc = 0;
    for (int t = t_begin; t < t_end; ++t, ++tc) {
      const int buf = tc & 1;
      const int r = t / a.nct, c = t - r * a.nct;
      const int f0 = r * UM_BM + q * 32;
      const int fr = min(f0 + lane, a.N - 1);                   // rows past N: computed on a valid frame, never stored
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      if (a.trans) {                                            // kernel-uniform: the translation is not folded into A
        t0 = __ldg(a.trans + (size_t)fr * 3); t1 = __ldg(a.trans + (size_t)fr * 3 + 1); t2 = __ldg(a.trans + (size_t)fr * 3 + 2);
      }
      const float* Arow = a.A + (size_t)fr * 624;
      const uint32_t rec = base + FG_OFF_REC + (uint32_t)buf * FG_REC_MAX;
      // this frame's row of the shaped template inside the tile's buffer (lanes of a warp: at most two different rows)
      const uint32_t vsr = a.vs ? base + FG_OFF_VS + (uint32_t)buf * FG_VS_BUF + (uint32_t)(fr / a.fpb - (r * UM_BM) / a.fpb) * FG_VS_ROW : 0u;
      mbar_wait(tfull0 + 8 * buf, (tc >> 1) & 1);
      mbar_wait(ttf0 + 8 * buf, (tc >> 1) & 1);
      tc_fence_after();
#ifdef HB_HOST_SHIM
      if (lane == 0) {                                          // (emulation only) the slots and the record this tile reads
        const int* tab = a.ft_tab + (size_t)c * FG_TAB;
        for (int i = 0; i < tab[0]; ++i) HB_EMU_GUARD_ACQ(base + FG_OFF_SLOTS + (uint32_t)(tab[4 + i] >> 16) * FG_SLOT, FG_SLOT);
        HB_EMU_GUARD_ACQ(rec, FG_REC_MAX);
      }
#endif
      // Two groups of 8 vertices per warp.  Sixteen epilogue warps (four per scheduler) hide the latency of the shared-memory
      // loads: 96 FMAs per (group, joint) against 5 warp-uniform LDS + 3 LDS.128 of the transform.  (Round 1 ran 8 warps at 135
      // registers with look-ahead registers and a branch per vertex and joint: 42 % of the issue slots,
      // profiles/r02a_fuseg35_set_full_details.txt; a half-group form at 16 warps doubled the loop overhead instead: r02f_*.)
#pragma unroll 1
      for (int gg = 0; gg < ((a.dbg & 2) ? 0 : 2); ++gg) {
        const int gi = h4 * 2 + gg;
        const int g = c * FG_GPT + gi;
        if (g >= a.num_groups) break;                           // warp-uniform
        float p[FG_GC], acc[FG_GC];
        tmem_ld24(trow + buf * FG_BN + gi * FG_GC, p);
        const int nv3 = min(FG_G, a.num_verts - g * FG_G) * 3;  // floats of this group inside the mesh
        if (a.vs) {                                             // kernel-uniform: + template + shape blend of the frame's sequence
#pragma unroll
          for (int k = 0; k < FG_GC / 4; ++k) {
            const float4 v = ld_shared_v4(vsr + (uint32_t)gi * (FG_GC * 4) + 16u * k);
            p[4 * k] += v.x; p[4 * k + 1] += v.y; p[4 * k + 2] += v.z; p[4 * k + 3] += v.w;
          }
        }
#pragma unroll
        for (int i = 0; i < FG_GC; ++i) acc[i] = 0.f;
        uint32_t ea = rec + FG_REC_HEAD + ld_shared_u32(rec + 4u * gi) * FG_REC_ENTRY;
        const uint32_t ee = rec + FG_REC_HEAD + ld_shared_u32(rec + 4u * gi + 4u) * FG_REC_ENTRY;
        int so = (int)ld_shared_u32(ea);                        // warp-uniform, fetched one entry ahead (the read behind the last
#pragma unroll 1                                                // entry stays inside the record buffer and is never used)
        for (; ea < ee; ea += FG_REC_ENTRY) {
          const float4 wa = ld_shared_v4(ea + 16u), wb = ld_shared_v4(ea + 32u);
          float4 r0, r1, r2;
          if (so >= 0) {
            r0 = ld_shared_v4(tsl_ + (uint32_t)so); r1 = ld_shared_v4(tsl_ + (uint32_t)so + 16u); r2 = ld_shared_v4(tsl_ + (uint32_t)so + 32u);
          } else {                                              // joint without a slot in this tile (rare): from L1/L2
            const float4* ap = reinterpret_cast<const float4*>(Arow + ld_shared_u32(ea + 4u));
            r0 = __ldg(ap); r1 = __ldg(ap + 1); r2 = __ldg(ap + 2);
          }
          so = (int)ld_shared_u32(ea + FG_REC_ENTRY);
          const float w[FG_G] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int i = 0; i < FG_G; ++i) {
            const float px = p[3 * i], py = p[3 * i + 1], pz = p[3 * i + 2];
            acc[3 * i] = fmaf(w[i], fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))), acc[3 * i]);
            acc[3 * i + 1] = fmaf(w[i], fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))), acc[3 * i + 1]);
            acc[3 * i + 2] = fmaf(w[i], fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))), acc[3 * i + 2]);
          }
        }
        if (a.trans) {
#pragma unroll
          for (int i = 0; i < FG_GC; ++i) acc[i] += (i % 3) == 0 ? t0 : ((i % 3) == 1 ? t1 : t2);
        }
        // park the group (lane = frame), then 2 2/3 frame rows of 96 bytes per store instruction
#pragma unroll
        for (int k = 0; k < 5; ++k) st_shared_v4(sw0_ + 16u * k, acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
        st_shared_v4(sw5_, acc[20], acc[21], acc[22], acc[23]);
        __syncwarp();
        float* tb = a.out + (size_t)f0 * rowf + (size_t)g * FG_GC;
        if (a.dbg & 1) continue;
        if (f0 + 32 <= a.N && nv3 == FG_GC) {                   // warp-uniform: every row and every column is stored
#pragma unroll
          for (int m = 0; m < 4; ++m) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const float2 v = ld_shared_v2(sr3[i] + 768u * m);
              stcs2(tb + (size_t)(8 * m) * rowf + go3[i], v.x, v.y);
            }
          }
        } else {                                                // ragged last row tile / the mesh's last, partial group
#pragma unroll 1
          for (int L = lane; L < 32 * 12; L += 32) {           // (offsets recomputed: no dynamic index into sr3 / go3)
            const int row = L / 12, j = L - 12 * row;
            int ch = (j >> 1) + ((row >> 2) & 1);
            if (ch >= 6) ch -= 6;
            if (f0 + row < a.N) {
              const float2 v = ld_shared_v2(stg + (uint32_t)(row * 96 + ch * 16 + (j & 1) * 8));
              float* dst = tb + (size_t)row * rowf + 2 * j;
              if (2 * j + 1 < nv3) stcs2(dst, v.x, v.y);
              else if (2 * j < nv3) stcs1(dst, v.x);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
#ifdef HB_HOST_SHIM
      if (lane == 0) {
        const int* tab = a.ft_tab + (size_t)c * FG_TAB;
        for (int i = 0; i < tab[0]; ++i) HB_EMU_GUARD_REL(base + FG_OFF_SLOTS + (uint32_t)(tab[4 + i] >> 16) * FG_SLOT);
        HB_EMU_GUARD_REL(rec);
      }
#endif
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace hb
