// Test infrastructure: functional CPU emulation of the Blackwell primitives the tcgen05 kernels use, so that the SAME
// kernel source (humor_b200/csrc/lbs_blend.cuh, lbs_fused.cuh) runs under the SIMT shim (cuda_runtime.h in this directory):
//   * a shared-memory window with 32-bit addresses (the dynamic buffer deliberately starts at address 16, so the kernels'
//     1024-byte align-up matters),
//   * mbarriers: pending-arrival count + transaction bytes + phase parity (init / arrive / arrive.expect_tx / complete_tx /
//     try_wait.parity; a wait that does not complete within 30 s aborts: protocol bugs fail, they do not hang),
//   * TMA 2-D tile loads with SWIZZLE_128B (16-byte chunk index XOR (address bits 7..9)), out-of-range elements = 0,
//   * tcgen05.mma kind::tf32, M = 128, K = 8 per instruction, operands read through K-major SWIZZLE_128B descriptors
//     (start address, SBO) exactly as issued by the kernels, operands truncated to tf32, fp32 accumulation into TMEM,
//   * TMEM as [128 lanes][512 columns] fp32, tcgen05.ld 32x32b.x32 / x8, alloc/dealloc, tcgen05.commit -> mbarrier arrive,
//   * un-swizzled and fp16 TMA boxes, tcgen05.mma kind::f16, read guards on shared-memory ranges (lbs_fuseg_kernel),
//   * thread-block clusters: one state (shared-memory window, TMEM, mbarriers) per CTA of the running cluster, the CTAs run
//     concurrently (shim::launch_cluster); %cluster_ctarank, barrier.cluster, mapa + st.shared::cluster (DSMEM stores).
// What it checks: tile/index arithmetic, descriptor offsets, barrier protocol (phases, buffer reuse), epilogues.  What it
// cannot check: that the hardware interprets descriptors / swizzles the way modelled here (lbs_fused_kernel, which IS
// verified on the B200, runs through the same emulation as a cross-check of the model), timing, real asynchrony.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>

#include "cuda.h"

namespace tcemu {

constexpr uint32_t WINDOW_BYTES = 256 * 1024;
constexpr uint32_t DYN_OFFSET = 16;
constexpr int MAX_CLUSTER = 8;
constexpr int MAX_CTAS = 16;                           // emulation state slots: a cluster, or a whole (small) concurrent grid
constexpr uint32_t RANK_SHIFT = 24;                    // mapa result: (rank + 1) << 24 | offset in that CTA's window
struct Bar { int expected = 0, pending = 0; long long tx = 0; int phase = 0; };
struct Guard { uint32_t lo, hi; int count; };
// one instance per CTA of the running cluster (rank = shim::t_crank; single-CTA launches use rank 0)
struct CtaState {
  alignas(1024) uint8_t smem[WINDOW_BYTES];
  float tmem[128][512];
  std::map<uint32_t, Bar> bars;
  std::map<uint32_t, Guard> guards;
};
inline CtaState g_cta[MAX_CTAS];
inline std::mutex g_mu;
inline long long g_mma_count = 0, g_tma_count = 0;
inline CtaState& cur() { return g_cta[shim::t_slot]; }
#define g_smem (tcemu::cur().smem)
#define g_tmem (tcemu::cur().tmem)
#define g_bars (tcemu::cur().bars)
#define g_guards (tcemu::cur().guards)

inline uint8_t* dyn_smem() { return cur().smem + DYN_OFFSET; }
inline uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const uint8_t*>(p) - cur().smem); }
inline uint32_t swz128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }
inline float tf32_trunc(float x) { uint32_t u; std::memcpy(&u, &x, 4); u &= 0xffffe000u; std::memcpy(&x, &u, 4); return x; }

inline void reset() {
  std::lock_guard<std::mutex> l(g_mu);
  for (auto& c : g_cta) {
    c.bars.clear();
    c.guards.clear();
    std::memset(c.smem, 0xff, sizeof(c.smem));            // NaN patterns: reading a byte nobody wrote shows
    for (auto& r : c.tmem) for (auto& v : r) v = __builtin_nanf("");
  }
  g_mma_count = g_tma_count = 0;
}
inline void complete_if_done(Bar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1; b.pending = b.expected; }
}
inline void mbar_init(uint32_t bar, int count) {
  std::lock_guard<std::mutex> l(g_mu);
  Bar b; b.expected = b.pending = count;
  g_bars[bar] = b;
}
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  std::lock_guard<std::mutex> l(g_mu);
  Bar& b = g_bars.at(bar);
  b.tx += bytes; b.pending -= 1;
  if (b.pending < 0) { std::fprintf(stderr, "tcemu: mbarrier %u over-arrived\n", bar); std::abort(); }
  complete_if_done(b);
}
inline void mbar_arrive(uint32_t bar) {
  std::lock_guard<std::mutex> l(g_mu);
  Bar& b = g_bars.at(bar);
  b.pending -= 1;
  if (b.pending < 0) { std::fprintf(stderr, "tcemu: mbarrier %u over-arrived\n", bar); std::abort(); }
  complete_if_done(b);
}
inline void complete_tx(uint32_t bar, uint32_t bytes) {
  std::lock_guard<std::mutex> l(g_mu);
  Bar& b = g_bars.at(bar);
  b.tx -= bytes;
  complete_if_done(b);
}
inline void mbar_wait(uint32_t bar, uint32_t parity) {
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    {
      std::lock_guard<std::mutex> l(g_mu);
      if ((uint32_t)g_bars.at(bar).phase != (parity & 1u)) return;      // the phase with this parity has completed
    }
    std::this_thread::yield();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) {
      std::fprintf(stderr, "tcemu: wait on mbarrier %u (parity %u) did not complete: protocol deadlock\n", bar, parity);
      std::abort();
    }
  }
}
// cp.async.bulk.tensor.2d ... mbarrier::complete_tx::bytes with a {32 floats, box_rows} box and SWIZZLE_128B
// shared-memory ranges a kernel declares "being read" (HB_EMU_GUARD_* in the kernel source, no-ops on the device): a TMA
// write into one of them is a protocol bug the synchronous emulation would otherwise hide
inline unsigned long long l2_policy_evict_last() { return 0ull; }
inline void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y);
inline void tma_load_2d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y, unsigned long long) { tma_load_2d(dst, m, bar, x, y); }
inline void guard_acquire(uint32_t addr, uint32_t bytes) {
  std::lock_guard<std::mutex> l(g_mu);
  Guard& g = g_guards[addr];
  g.lo = addr; g.hi = addr + bytes; g.count += 1;
}
inline void guard_release(uint32_t addr) {
  std::lock_guard<std::mutex> l(g_mu);
  Guard& g = g_guards.at(addr);
  if (--g.count < 0) { std::fprintf(stderr, "tcemu: guard %u released twice\n", addr); std::abort(); }
}
inline void guard_check_write(uint32_t dst, uint32_t bytes) {
  std::lock_guard<std::mutex> l(g_mu);
  for (const auto& kv : g_guards)
    if (kv.second.count > 0 && dst < kv.second.hi && kv.second.lo < dst + bytes) {
      std::fprintf(stderr, "tcemu: TMA write [%u, %u) into a range still being read [%u, %u)\n", dst, dst + bytes, kv.second.lo, kv.second.hi);
      std::abort();
    }
}
inline void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
  if (m->plain) {                 // SWIZZLE_NONE: box rows back to back, 16-byte granules
    const uint32_t rowb = m->box_cols * 4u;
    if (dst % 128u || rowb % 16u) { std::fprintf(stderr, "tcemu: plain TMA box: destination %u / row of %u bytes not aligned\n", dst, rowb); std::abort(); }
    guard_check_write(dst, m->box_rows * rowb);
    for (unsigned r = 0; r < m->box_rows; ++r)
      for (unsigned c = 0; c < m->box_cols; ++c) {
        const long long gr = (long long)y + r, gc = (long long)x + c;
        const float v = (gr >= 0 && gc >= 0 && (unsigned long long)gr < m->rows && (unsigned long long)gc < m->cols) ? m->base[gr * m->ld + gc] : 0.f;
        std::memcpy(g_smem + dst + r * rowb + c * 4u, &v, 4);
      }
    { std::lock_guard<std::mutex> l(g_mu); ++g_tma_count; }
    complete_tx(bar, m->box_rows * rowb);
    return;
  }
  if (dst % 1024u) { std::fprintf(stderr, "tcemu: TMA destination %u is not 1024-byte aligned\n", dst); std::abort(); }
  if (m->half) {                  // fp16 elements, SWIZZLE_128B: 64 halves per box row
    if (m->box_cols != 64) { std::fprintf(stderr, "tcemu: fp16 box must be one 128-byte swizzle span (64 halves) wide\n"); std::abort(); }
    const uint16_t* hb = reinterpret_cast<const uint16_t*>(m->base);
    for (unsigned r = 0; r < m->box_rows; ++r)
      for (unsigned c = 0; c < 64; ++c) {
        const long long gr = (long long)y + r, gc = (long long)x + c;
        const uint16_t v = (gr >= 0 && gc >= 0 && (unsigned long long)gr < m->rows && (unsigned long long)gc < m->cols) ? hb[gr * m->ld + gc] : (uint16_t)0;
        const uint32_t a = swz128(dst + r * 128u + c * 2u);
        std::memcpy(g_smem + a, &v, 2);
      }
    { std::lock_guard<std::mutex> l(g_mu); ++g_tma_count; }
    complete_tx(bar, m->box_rows * 128u);
    return;
  }
  if (m->box_cols != 32) { std::fprintf(stderr, "tcemu: box must be one 128-byte swizzle span wide\n"); std::abort(); }
  for (unsigned r = 0; r < m->box_rows; ++r)
    for (unsigned c = 0; c < 32; ++c) {
      const long long gr = (long long)y + r, gc = (long long)x + c;
      const float v = (gr >= 0 && gc >= 0 && (unsigned long long)gr < m->rows && (unsigned long long)gc < m->cols) ? m->base[gr * m->ld + gc] : 0.f;
      const uint32_t a = swz128(dst + r * 128u + c * 4u);
      std::memcpy(g_smem + a, &v, 4);
    }
  { std::lock_guard<std::mutex> l(g_mu); ++g_tma_count; }
  complete_tx(bar, m->box_rows * 128u);
}
// tcgen05.mma.cta_group::1.kind::tf32, M = 128 (idesc bits 24..28), N = idesc bits 17..22 << 3, K = 8 (32 bytes)
inline void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  if (M != 128 || N < 8 || N > 256 || (N % 16)) { std::fprintf(stderr, "tcemu: unsupported UMMA shape %dx%d\n", M, N); std::abort(); }
  if (((adesc >> 61) & 7) != 2 || ((bdesc >> 61) & 7) != 2) { std::fprintf(stderr, "tcemu: descriptors must be SWIZZLE_128B\n"); std::abort(); }
  const uint32_t a0 = (uint32_t)(adesc & 0x3FFF) << 4, b0 = (uint32_t)(bdesc & 0x3FFF) << 4;
  const uint32_t asbo = (uint32_t)((adesc >> 32) & 0x3FFF) << 4, bsbo = (uint32_t)((bdesc >> 32) & 0x3FFF) << 4;
  static thread_local float A[128][8], B[256][8];
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 8; ++k) {
      float v; std::memcpy(&v, g_smem + swz128(a0 + (m / 8) * asbo + (m % 8) * 128u + k * 4u), 4);
      A[m][k] = tf32_trunc(v);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 8; ++k) {
      float v; std::memcpy(&v, g_smem + swz128(b0 + (n / 8) * bsbo + (n % 8) * 128u + k * 4u), 4);
      B[n][k] = tf32_trunc(v);
    }
  const uint32_t lane0 = tmem_d >> 16, col0 = tmem_d & 0xFFFFu;
  if (lane0 != 0 || col0 + (uint32_t)N > 512u) { std::fprintf(stderr, "tcemu: accumulator outside TMEM\n"); std::abort(); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = accum ? g_tmem[m][col0 + n] : 0.f;
      for (int k = 0; k < 8; ++k) s += A[m][k] * B[n][k];
      g_tmem[m][col0 + n] = s;
    }
  std::lock_guard<std::mutex> l(g_mu);
  ++g_mma_count;
}
// tcgen05.mma.cta_group::1.kind::f16 with fp16 operands (idesc a/b format 0), M = 128, K = 16 per instruction (32 bytes), fp32
// accumulation into the same TMEM columns the tf32 MMAs use
inline void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  if (M != 128 || N < 8 || N > 256 || (N % 16)) { std::fprintf(stderr, "tcemu: unsupported UMMA shape %dx%d\n", M, N); std::abort(); }
  if (((idesc >> 7) & 7) != 0 || ((idesc >> 10) & 7) != 0 || ((idesc >> 4) & 3) != 1) { std::fprintf(stderr, "tcemu: kind::f16 expects F16 x F16 -> F32\n"); std::abort(); }
  if (((adesc >> 61) & 7) != 2 || ((bdesc >> 61) & 7) != 2) { std::fprintf(stderr, "tcemu: descriptors must be SWIZZLE_128B\n"); std::abort(); }
  const uint32_t a0 = (uint32_t)(adesc & 0x3FFF) << 4, b0 = (uint32_t)(bdesc & 0x3FFF) << 4;
  const uint32_t asbo = (uint32_t)((adesc >> 32) & 0x3FFF) << 4, bsbo = (uint32_t)((bdesc >> 32) & 0x3FFF) << 4;
  static thread_local float A[128][16], B[256][16];
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 16; ++k) {
      _Float16 v; std::memcpy(&v, g_smem + swz128(a0 + (m / 8) * asbo + (m % 8) * 128u + k * 2u), 2);
      A[m][k] = (float)v;
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k) {
      _Float16 v; std::memcpy(&v, g_smem + swz128(b0 + (n / 8) * bsbo + (n % 8) * 128u + k * 2u), 2);
      B[n][k] = (float)v;
    }
  const uint32_t lane0 = tmem_d >> 16, col0 = tmem_d & 0xFFFFu;
  if (lane0 != 0 || col0 + (uint32_t)N > 512u) { std::fprintf(stderr, "tcemu: accumulator outside TMEM\n"); std::abort(); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = accum ? g_tmem[m][col0 + n] : 0.f;
      for (int k = 0; k < 16; ++k) s += A[m][k] * B[n][k];
      g_tmem[m][col0 + n] = s;
    }
  std::lock_guard<std::mutex> l(g_mu);
  ++g_mma_count;
}
inline uint16_t f32_to_f16_bits(float x) { const _Float16 h = (_Float16)x; uint16_t b; std::memcpy(&b, &h, 2); return b; }
inline void umma_commit(uint32_t bar) { mbar_arrive(bar); }      // the MMAs above ran synchronously
inline void tmem_ld32(uint32_t taddr, float* v) {
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31u), col = taddr & 0xFFFFu;
  if (lane >= 128u || col + 32u > 512u) { std::fprintf(stderr, "tcemu: tcgen05.ld outside TMEM\n"); std::abort(); }
  // a warp may only touch its own lane quadrant (warp id % 4)
  if ((taddr >> 16) != ((threadIdx.x >> 5) & 3u) * 32u) { std::fprintf(stderr, "tcemu: warp reads a foreign TMEM quadrant\n"); std::abort(); }
  for (int j = 0; j < 32; ++j) v[j] = g_tmem[lane][col + j];
}
inline void tmem_ld24(uint32_t taddr, float* v) {              // 32x32b.x16 + .x8: 24 consecutive columns
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31u), col = taddr & 0xFFFFu;
  if (lane >= 128u || col + 24u > 512u) { std::fprintf(stderr, "tcemu: tcgen05.ld outside TMEM\n"); std::abort(); }
  if ((taddr >> 16) != ((threadIdx.x >> 5) & 3u) * 32u) { std::fprintf(stderr, "tcemu: warp reads a foreign TMEM quadrant\n"); std::abort(); }
  for (int j = 0; j < 24; ++j) v[j] = g_tmem[lane][col + j];
}
inline void tmem_ld12(uint32_t taddr, float* v) {              // 32x32b.x4 three times: 12 consecutive columns
  const uint32_t lane = (taddr >> 16) + (threadIdx.x & 31u), col = taddr & 0xFFFFu;
  if (lane >= 128u || col + 12u > 512u || (col & 3u)) { std::fprintf(stderr, "tcemu: tcgen05.ld outside TMEM / misaligned\n"); std::abort(); }
  if ((taddr >> 16) != ((threadIdx.x >> 5) & 3u) * 32u) { std::fprintf(stderr, "tcemu: warp reads a foreign TMEM quadrant\n"); std::abort(); }
  for (int j = 0; j < 12; ++j) v[j] = g_tmem[lane][col + j];
}
inline void tmem_alloc(uint32_t dst, uint32_t ncols) {
  if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) { std::fprintf(stderr, "tcemu: bad TMEM allocation %u\n", ncols); std::abort(); }
  const uint32_t base = 0;
  std::memcpy(g_smem + dst, &base, 4);
}
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline uint32_t ld_shared_u32(uint32_t addr) { uint32_t v; std::memcpy(&v, g_smem + addr, 4); return v; }
inline float4 ld_shared_v4(uint32_t addr) { float4 v; std::memcpy(&v, g_smem + addr, 16); return v; }
inline float2 ld_shared_v2(uint32_t addr) {
  if (addr & 7u) { std::fprintf(stderr, "tcemu: bad ld.shared.v2 address %u\n", addr); std::abort(); }
  float2 v; std::memcpy(&v, g_smem + addr, 8); return v;
}
// thread-block clusters: the CTAs of a cluster run concurrently (shim::launch_cluster), each on its own CtaState.  What the
// split-K GEMM uses: %cluster_ctarank, barrier.cluster (all threads of all CTAs), mapa + st.shared::cluster (DSMEM stores).
inline uint32_t cluster_ctarank() { return shim::t_crank; }
inline void cluster_sync_all() { shim::sync_cluster(); }
inline uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  const uint32_t slot = shim::t_slot - shim::t_crank + rank;      // the cluster's CTAs occupy consecutive state slots
  if (rank >= shim::cluster_size || slot >= (uint32_t)MAX_CTAS || local >= WINDOW_BYTES) { std::fprintf(stderr, "tcemu: mapa to CTA %u of a %u-CTA cluster / offset %u\n", rank, shim::cluster_size, local); std::abort(); }
  return ((slot + 1u) << RANK_SHIFT) | local;
}
inline void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  const uint32_t r = addr >> RANK_SHIFT, off = addr & ((1u << RANK_SHIFT) - 1u);
  if (r == 0 || r > (uint32_t)MAX_CTAS || off + 16u > WINDOW_BYTES || (off & 15u)) { std::fprintf(stderr, "tcemu: bad st.shared::cluster address %u\n", addr); std::abort(); }
  const float v[4] = {a, b, c, d};
  std::memcpy(g_cta[r - 1].smem + off, v, 16);
}
// ---- persistent decoder chain (chain_persist.cuh): cluster ids, remote mbarrier arrivals, data-flow flags, named barrier
inline uint32_t cluster_id_x() { return blockIdx.x / shim::cluster_size; }
inline uint32_t cluster_nid_x() { return gridDim.x / shim::cluster_size; }
inline void mbar_arrive_remote(uint32_t addr) {                  // addr from map_to_cta
  const uint32_t r = addr >> RANK_SHIFT, off = addr & ((1u << RANK_SHIFT) - 1u);
  if (r == 0 || r > (uint32_t)MAX_CTAS) { std::fprintf(stderr, "tcemu: bad remote mbarrier address %u\n", addr); std::abort(); }
  std::lock_guard<std::mutex> l(g_mu);
  Bar& b = g_cta[r - 1].bars.at(off);
  b.pending -= 1;
  if (b.pending < 0) { std::fprintf(stderr, "tcemu: mbarrier %u of CTA slot %u over-arrived\n", off, r - 1); std::abort(); }
  complete_if_done(b);
}
inline void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
// st.async ... mbarrier::complete_tx::bytes.v4: 16 bytes into a peer's window + 16 transaction bytes on the peer's barrier
inline void st_async_v4(uint32_t addr, float a, float b, float c, float d, uint32_t mbar) {
  st_cluster_v4(addr, a, b, c, d);
  const uint32_t r = mbar >> RANK_SHIFT, off = mbar & ((1u << RANK_SHIFT) - 1u);
  if (r == 0 || r > (uint32_t)MAX_CTAS || r != (addr >> RANK_SHIFT)) { std::fprintf(stderr, "tcemu: st.async barrier %u is not in the destination CTA of %u\n", mbar, addr); std::abort(); }
  std::lock_guard<std::mutex> l(g_mu);
  Bar& bb = g_cta[r - 1].bars.at(off);
  bb.tx -= 16;
  complete_if_done(bb);
}
inline void fence_gpu() {}
inline void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  if ((addr & 15u) || addr + 16u > WINDOW_BYTES) { std::fprintf(stderr, "tcemu: bad st.shared.v4 address %u\n", addr); std::abort(); }
  const float v[4] = {a, b, c, d};
  std::memcpy(g_smem + addr, v, 16);
}
// cp.async.bulk.shared::cluster.shared::cta: `bytes` from this CTA's window into a peer's, completing them on the peer's barrier
inline void bulk_s2c(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t mbar) {
  const uint32_t r = dst >> RANK_SHIFT, off = dst & ((1u << RANK_SHIFT) - 1u), rb = mbar >> RANK_SHIFT, ob = mbar & ((1u << RANK_SHIFT) - 1u);
  if (r == 0 || r > (uint32_t)MAX_CTAS || rb != r || (off & 15u) || (src & 15u) || (bytes & 15u) || off + bytes > WINDOW_BYTES || src + bytes > WINDOW_BYTES) {
    std::fprintf(stderr, "tcemu: bad bulk shared->cluster copy %u -> %u (%u bytes, barrier %u)\n", src, dst, bytes, mbar);
    std::abort();
  }
  std::memcpy(g_cta[r - 1].smem + off, g_smem + src, bytes);
  std::lock_guard<std::mutex> l(g_mu);
  Bar& bb = g_cta[r - 1].bars.at(ob);
  bb.tx -= bytes;
  complete_if_done(bb);
}
// cp.async.bulk global -> shared (1-D) with mbarrier::complete_tx
inline void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  if ((dst & 15u) || (bytes & 15u) || (reinterpret_cast<uintptr_t>(src) & 15u) || dst + bytes > WINDOW_BYTES) {
    std::fprintf(stderr, "tcemu: bulk copy %p -> %u (%u bytes) is not 16-byte aligned / in range\n", src, dst, bytes);
    std::abort();
  }
  guard_check_write(dst, bytes);
  std::memcpy(g_smem + dst, src, bytes);
  complete_tx(bar, bytes);
}
inline void flag_wait_ge(const unsigned* p, unsigned target) {
  const auto t0 = std::chrono::steady_clock::now();
  while (__atomic_load_n(p, __ATOMIC_ACQUIRE) < target) {
    std::this_thread::yield();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
      std::fprintf(stderr, "tcemu: data-flow flag %p stuck at %u < %u: protocol deadlock\n", (const void*)p, *p, target);
      std::abort();
    }
  }
}
inline void flag_add_release(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline void fence_proxy_async() {}
inline void epi_bar_sync() { shim::t_cta->epi_bar->arrive_and_wait(); }

}  // namespace tcemu
