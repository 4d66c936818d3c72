// Test infrastructure: functional stand-ins for the tcgen05 / TMEM / TMA / mbarrier wrappers of leopard_b200/csrc/lp_common.cuh,
// on top of cuda_emu.h, so that the SOURCE of the tensor-core kernels can run on the CPU (tests/test_attention_kernels_emulated.py,
// tests/test_gemm_kernel_emulated.py).  What is modelled — and only as far as the kernels rely on it:
//   mbarrier   arrival count + transaction bytes + phase parity; try_wait(parity) is true once the phase of that parity has
//              completed (a fresh barrier passes parity 1); a wait that makes no progress for EMU_DEADLOCK_MS aborts the run
//   TMA        2-D / 3-D tiled bf16 loads with 128-byte swizzle (16-byte chunk index XOR row-within-8), zero fill out of
//              bounds; 2-D tile stores (swizzle 0 / 64 / 128 bytes, clipped at the tensor's extents)
//   UMMA       kind::f16 bf16 x bf16 -> fp32, K = 16 per instruction, executed synchronously in issue order; A from shared
//              memory (K-major) or from TMEM (packed bf16 pairs), B K-major or MN-major, both through the 64-bit descriptors
//              (start, LBO, SBO) with the swizzle applied to the absolute address bits; cta_group::2: M = 256 over the two
//              CTAs of a pair (A rows and B rows split between their shared memories, D rows between their TMEMs)
//   TMEM       128 lanes x 512 columns of 32-bit cells per CTA; 32x32b loads / stores address lane = base lane + lane id
//   cluster    clusters of 2 or 4 CTAs run concurrently (emu_tc::launch_clusters): %cluster_ctarank, mapa, remote mbarrier
//              arrive, barrier.cluster, masked multicast commits, cta_group::2 multicast TMA loads (bytes counted on each
//              destination's pair leader)
//   named barriers  bar.sync / bar.arrive id, n
// The layouts are the author's reading of the hardware, the same reading the GPU-verified kernels are built on; the emulator
// checks a kernel's LOGIC (protocol, indexing, masks, operand roles), not the hardware.
#pragma once
#include "cuda_emu.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#ifndef EMU_DEADLOCK_MS
#define EMU_DEADLOCK_MS 20000
#endif

#undef __shared__
#define __shared__

struct CUtensorMap {
  const uint8_t* base;
  uint64_t dims[3];      // elements, innermost first
  uint64_t strides[2];   // bytes, for dims 1 and 2
  uint32_t box[3];
  int elem_bytes = 2;
  int swizzle = 128;     // bytes: 0, 64 or 128
};

namespace lp {
typedef __nv_bfloat16 bf16;
// kernels that declare `extern __shared__ uint8_t smem_raw[];` at block scope inside namespace lp name lp::smem_raw: the
// shared memory of CTA 0 (single-CTA launches).  Cluster launches give every CTA its own buffer (emu_tc::cta().smem) and
// the test rewrites that declaration into `uint8_t* smem_raw = emu_tc::cta().smem;`.
alignas(1024) uint8_t smem_raw[232448];

namespace emu_tc {
struct Mbar { int expected = 0, pending = 0; long tx = 0; int phase = 0; };
struct NamedBar { int count = 0; long gen = 0; };
struct Cta {
  uint8_t* smem = nullptr;
  uint32_t (*tmem)[512] = nullptr;                 // [128][512]
  std::unordered_map<uint32_t, Mbar> bars;
  NamedBar nbar[16];
  std::unique_ptr<std::barrier<>> named;           // the 128 row threads of lp_attn_bwd.cu
  uint32_t rank = 0;
};
inline std::mutex g_mu;
inline Cta g_cta[4];
alignas(1024) inline uint8_t g_smem_peers[3][232448];
inline uint32_t g_tmem_store[4][128][512];
inline thread_local Cta* t_cta = &g_cta[0];
inline std::unique_ptr<std::barrier<>> g_cluster_bar;
inline Cta& cta() { return *t_cta; }

inline void reset_cta(int r) {
  Cta& c = g_cta[r];
  c.smem = r == 0 ? smem_raw : g_smem_peers[r - 1];
  c.tmem = g_tmem_store[r];
  c.bars.clear();
  for (auto& b : c.nbar) b = NamedBar();
  c.named = std::make_unique<std::barrier<>>(128);
  c.rank = (uint32_t)r;
  std::memset(c.tmem, 0xff, sizeof(g_tmem_store[r]));   // NaN pattern: an accumulator read before it is written shows up
}
inline void reset() {   // single-CTA launches: per-block hook of emu::launch
  std::lock_guard<std::mutex> l(g_mu);
  reset_cta(0);
}
inline void maybe_complete(Mbar& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1;
    b.pending = b.expected;
  }
}
inline float bf_bits(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }
inline uint32_t swz64(uint32_t a) { return a ^ (((a >> 7) & 3u) << 4); }
inline uint32_t swz(uint32_t a, int mode) { return mode == 128 ? swz128(a) : mode == 64 ? swz64(a) : a; }
// shared::cluster addresses: bits [28, 30) = 1 + CTA rank, 0 = the executing CTA
inline Cta& cta_of(uint32_t addr) { return (addr >> 28) ? g_cta[(addr >> 28) - 1] : cta(); }
inline uint32_t off_of(uint32_t addr) { return addr & 0x0fffffffu; }

// Runs the grid as concurrent CLUSTERS of `csize` CTAs (2 or 4): blocks csize k .. csize k + csize - 1 together.
template <typename F>
void launch_clusters(dim3 grid, dim3 block, unsigned csize, F&& body) {
  const unsigned nthreads = block.x;
  for (unsigned p = 0; p < grid.x / csize; ++p) {
    {
      std::lock_guard<std::mutex> l(g_mu);
      for (unsigned r = 0; r < csize; ++r) reset_cta((int)r);
    }
    g_cluster_bar = std::make_unique<std::barrier<>>(csize * nthreads);
    std::vector<::emu::BlockCtx> ctx(csize);
    for (unsigned r = 0; r < csize; ++r) {
      ctx[r].block_bar = std::make_unique<std::barrier<>>(nthreads);
      for (unsigned w = 0; w < (nthreads + 31) / 32; ++w)
        ctx[r].warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32u, nthreads - w * 32)));
      ctx[r].warp_buf.assign(((nthreads + 31) / 32) * 32, 0.f);
    }
    std::vector<std::thread> ts;
    for (unsigned r = 0; r < csize; ++r)
      for (unsigned t = 0; t < nthreads; ++t)
        ts.emplace_back([&, r, t] {
          threadIdx = {t, 0, 0};
          blockIdx = {csize * p + r, 0, 0};
          blockDim = block;
          gridDim = grid;
          t_cta = &g_cta[r];
          ::emu::t_ctx = &ctx[r];
          body();
          ctx[r].block_bar->arrive_and_drop();
        });
    for (auto& th : ts) th.join();
  }
}
template <typename F>
void launch_pairs(dim3 grid, dim3 block, F&& body) { launch_clusters(grid, block, 2, body); }
}  // namespace emu_tc

inline uint32_t smem_u32(const void* p) { return (uint32_t)((const uint8_t*)p - emu_tc::cta().smem); }

// ---- mbarrier (addresses may be shared::cluster addresses naming the peer CTA)
inline emu_tc::Mbar& mbar_at(uint32_t addr) { return emu_tc::cta_of(addr).bars.at(emu_tc::off_of(addr)); }
inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  emu_tc::Mbar b;
  b.expected = b.pending = (int)count;
  emu_tc::cta().bars[smem_u32(bar)] = b;
}
inline void fence_barrier_init() {}
inline void fence_proxy_async_smem() {}
inline void fence_proxy_async_all() {}
#ifndef EMU_HOST_ATOMICS       // host_emu.h (never included together with this header) has its own
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
#endif
inline unsigned ld_acquire_sys(const unsigned* p) { return std::atomic_ref<unsigned>(*const_cast<unsigned*>(p)).load(); }
inline unsigned ld_acquire_gpu(const unsigned* p) { return ld_acquire_sys(p); }
inline void st_release_sys(unsigned* p, unsigned v) { std::atomic_ref<unsigned>(*p).store(v); }
inline void st_release_gpu(unsigned* p, unsigned v) { st_release_sys(p, v); }
inline void st_relaxed_gpu(unsigned* p, unsigned v) { st_release_sys(p, v); }
inline uint4 lds_u4(const void* p) { uint4 v; std::memcpy(&v, p, 16); return v; }
inline void sts_u4(void* p, uint4 v) { std::memcpy(p, &v, 16); }
inline uint4 ld_cg_u4(const float* p) { uint4 v; std::memcpy(&v, p, 16); return v; }
inline void st_cg_u4(float* p, uint4 v) { std::memcpy(p, &v, 16); }
inline void spin_until_sys(const unsigned* p, unsigned target) { while ((int)(ld_acquire_sys(p) - target) < 0) std::this_thread::yield(); }
inline void spin_until_gpu(const unsigned* p, unsigned target) { while (ld_acquire_gpu(p) < target) std::this_thread::yield(); }
inline void mbar_arrive_addr(uint32_t addr) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  auto& b = mbar_at(addr);
  b.pending -= 1;
  if (b.pending < 0) { std::fprintf(stderr, "emu: mbarrier over-arrival @%u\n", addr); std::abort(); }
  emu_tc::maybe_complete(b);
}
inline void mbar_arrive(uint64_t* bar) { mbar_arrive_addr(smem_u32(bar)); }
inline void mbar_arrive_cluster(uint32_t cluster_addr) { mbar_arrive_addr(cluster_addr); }
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  auto& b = mbar_at(smem_u32(bar));
  b.tx += bytes;
  b.pending -= 1;
  emu_tc::maybe_complete(b);
}
inline void mbar_complete_tx_addr(uint32_t addr, uint32_t bytes) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  auto& b = mbar_at(addr);
  b.tx -= bytes;
  emu_tc::maybe_complete(b);
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  return (uint32_t)mbar_at(smem_u32(bar)).phase != parity;
}
inline bool mbar_test_wait(uint64_t* bar, uint32_t parity) { return mbar_try_wait(bar, parity); }
inline void mbar_wait(uint64_t* bar, uint32_t parity);
inline void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned) { mbar_wait(bar, parity); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  const auto t0 = std::chrono::steady_clock::now();
  while (!mbar_try_wait(bar, parity)) {
    std::this_thread::yield();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(EMU_DEADLOCK_MS)) {
      std::fprintf(stderr, "emu: DEADLOCK block (%u,%u,%u) thread %u waits bar@%u parity %u\n", blockIdx.x, blockIdx.y,
                   blockIdx.z, threadIdx.x, smem_u32(bar), parity);
      std::abort();
    }
  }
}

// ---- named barriers (bar.sync id, n waits for the n-th arrival of the current generation; bar.arrive id, n only arrives)
inline void named_bar(int id, int n, bool wait) {
  long my_gen;
  {
    std::lock_guard<std::mutex> l(emu_tc::g_mu);
    auto& b = emu_tc::cta().nbar[id];
    my_gen = b.gen;
    if (++b.count == n) {
      b.count = 0;
      b.gen++;
      return;
    }
  }
  if (!wait) return;
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    {
      std::lock_guard<std::mutex> l(emu_tc::g_mu);
      if (emu_tc::cta().nbar[id].gen != my_gen) return;
    }
    std::this_thread::yield();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(EMU_DEADLOCK_MS)) {
      std::fprintf(stderr, "emu: DEADLOCK thread %u waits named barrier %d\n", threadIdx.x, id);
      std::abort();
    }
  }
}
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}

// ---- cluster
inline uint32_t cluster_ctarank() { return emu_tc::cta().rank; }
inline void cluster_sync_all() { emu_tc::g_cluster_bar->arrive_and_wait(); }
inline uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) { return emu_tc::off_of(local_addr) | ((rank + 1u) << 28); }

// ---- TMA
inline void tma_prefetch_desc(const CUtensorMap*) {}
inline void tma_prefetch_l2_2d(const CUtensorMap*, int, int) {}
inline void tma_load_tile_to(emu_tc::Cta& dst_cta, uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2) {
  const uint32_t row_bytes = m->box[0] * m->elem_bytes;
  uint8_t* sm = dst_cta.smem;
  for (uint32_t r = 0; r < m->box[1]; ++r)
    for (uint32_t e = 0; e < m->box[0]; ++e) {
      const long col = (long)c0 + e, row = (long)c1 + r;
      uint16_t v = 0;
      if (col >= 0 && row >= 0 && (uint64_t)col < m->dims[0] && (uint64_t)row < m->dims[1] && (uint64_t)c2 < m->dims[2])
        std::memcpy(&v, m->base + col * 2 + row * m->strides[0] + (uint64_t)c2 * m->strides[1], 2);
      std::memcpy(sm + emu_tc::swz(dst + r * row_bytes + e * 2, m->swizzle), &v, 2);
    }
}
inline void tma_load_tile(void* smem_dst, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1, int c2) {
  tma_load_tile_to(emu_tc::cta(), smem_u32(smem_dst), m, c0, c1, c2);
  mbar_complete_tx_addr(bar_addr, m->box[0] * m->box[1] * m->elem_bytes);
}
// cta_group::2 multicast: the box lands at the same offset in every CTA of the mask; each destination's bytes are counted on
// the barrier at the given offset in that destination's PAIR LEADER (even rank)
inline void tma_load_2d_pair_mcast(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, uint16_t mask) {
  const uint32_t dst = smem_u32(smem_dst), boff = emu_tc::off_of(bar_cluster_addr);
  for (uint32_t d = 0; d < 4; ++d)
    if (mask & (1u << d)) {
      tma_load_tile_to(emu_tc::g_cta[d], dst, m, c0, c1, 0);
      mbar_complete_tx_addr(boff | (((d & ~1u) + 1u) << 28), m->box[0] * m->box[1] * m->elem_bytes);
    }
}
inline void tma_load_3d(void* d, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) { tma_load_tile(d, m, smem_u32(bar), c0, c1, c2); }
inline void tma_load_2d(void* d, const CUtensorMap* m, uint64_t* bar, int c0, int c1) { tma_load_tile(d, m, smem_u32(bar), c0, c1, 0); }
inline void tma_load_2d_pair(void* d, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  tma_load_tile(d, m, bar_cluster_addr, c0, c1, 0);
}
inline void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  const uint32_t src = smem_u32(smem_src);
  const uint32_t row_bytes = m->box[0] * m->elem_bytes;
  const uint8_t* sm = emu_tc::cta().smem;
  for (uint32_t r = 0; r < m->box[1]; ++r)
    for (uint32_t e = 0; e < m->box[0]; ++e) {
      const long col = (long)c0 + e, row = (long)c1 + r;
      if (col < 0 || row < 0 || (uint64_t)col >= m->dims[0] || (uint64_t)row >= m->dims[1]) continue;   // clipped
      std::memcpy(const_cast<uint8_t*>(m->base) + col * m->elem_bytes + row * m->strides[0],
                  sm + emu_tc::swz(src + r * row_bytes + e * m->elem_bytes, m->swizzle), m->elem_bytes);
    }
}
inline void tma_store_commit() {}
template <int N> inline void tma_store_wait_read() {}
template <int N> inline void tma_store_wait() {}

// ---- TMEM
inline void tmem_alloc(uint32_t* smem_dst, uint32_t) { *smem_dst = 0; }
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tmem_alloc_pair(uint32_t* smem_dst, uint32_t) { *smem_dst = 0; }
inline void tmem_relinquish_pair() {}
inline void tmem_dealloc_pair(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void tmem_wait_ld() {}
inline void tmem_wait_st() {}
inline void tmem_ld(uint32_t* regs, uint32_t taddr, int n) {
  const uint32_t lane = (taddr >> 16) + (::emu::linear_tid() & 31), col = taddr & 0xffff;
  for (int i = 0; i < n; ++i) regs[i] = emu_tc::cta().tmem[lane][col + i];
}
inline void tmem_st(const uint32_t* regs, uint32_t taddr, int n) {
  const uint32_t lane = (taddr >> 16) + (::emu::linear_tid() & 31), col = taddr & 0xffff;
  for (int i = 0; i < n; ++i) emu_tc::cta().tmem[lane][col + i] = regs[i];
}
#define LP_TMEM_LD_32(regs, taddr) lp::tmem_ld(regs, taddr, 32)
#define LP_TMEM_LD_16(regs, taddr) lp::tmem_ld(regs, taddr, 16)
#define LP_TMEM_ST_16(regs, taddr) lp::tmem_st(regs, taddr, 16)
#define LP_TMEM_ST_32(regs, taddr) lp::tmem_st(regs, taddr, 32)

// ---- UMMA (descriptor bit layouts: see umma_desc_sw128 / umma_idesc_bf16, taken from lp_common.cuh by the tests)
// pair = false: one CTA, M = 128.  pair = true (cta_group::2): M = 256; accumulator / A row m lives in CTA m / 128, B row n
// in CTA n / (N / 2); the descriptors name the same shared-memory offsets in both CTAs.
inline void umma_exec(uint32_t tmem_d, bool a_in_tmem, uint32_t tmem_a, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                      uint32_t accumulate, bool pair) {
  const int M = (int)((idesc >> 24) & 0x1f) << 4, N = (int)((idesc >> 17) & 0x3f) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (M != (pair ? 256 : 128) || (a_mn && a_in_tmem) || (pair && a_in_tmem)) {
    std::fprintf(stderr, "emu: unsupported UMMA shape M=%d a_mn=%d pair=%d\n", M, (int)a_mn, (int)pair);
    std::abort();
  }
  auto start = [](uint64_t d) { return (uint32_t)(d & 0x3FFF) << 4; };
  auto lbo = [](uint64_t d) { return (uint32_t)((d >> 16) & 0x3FFF) << 4; };
  auto sbo = [](uint64_t d) { return (uint32_t)((d >> 32) & 0x3FFF) << 4; };
  auto ld = [](const uint8_t* sm, uint32_t addr) {
    uint16_t h;
    std::memcpy(&h, sm + emu_tc::swz128(addr), 2);
    return emu_tc::bf_bits(h);
  };
  emu_tc::Cta& self = emu_tc::cta();
  const uint32_t dcol = tmem_d & 0xffff, acol = tmem_a & 0xffff;
  const int nb = pair ? N / 2 : N;   // B rows held by one CTA
  for (int m = 0; m < M; ++m) {
    emu_tc::Cta& cm = pair ? emu_tc::g_cta[(self.rank & ~1u) + (m >> 7)] : self;
    const int ml = m & 127;
    float a[16];
    for (int k = 0; k < 16; ++k) {
      if (a_in_tmem) {
        const uint32_t w = cm.tmem[ml][acol + (k >> 1)];
        a[k] = emu_tc::bf_bits((uint16_t)((k & 1) ? (w >> 16) : (w & 0xffff)));
      } else {
        // MN-major A (transposed operand): 64-row (m) slabs LBO apart, k rows of 128 B in 8-row groups SBO apart — the same
        // reading of the descriptor as the MN-major B below, which the attention kernels validate on hardware
        a[k] = a_mn ? ld(cm.smem, start(desc_a) + (ml >> 6) * lbo(desc_a) + (k >> 3) * sbo(desc_a) + (k & 7) * 128 + (ml & 63) * 2)
                    : ld(cm.smem, start(desc_a) + (ml >> 3) * sbo(desc_a) + (ml & 7) * 128 + k * 2);
      }
    }
    for (int n = 0; n < N; ++n) {
      const uint8_t* bs = pair ? emu_tc::g_cta[(self.rank & ~1u) + n / nb].smem : self.smem;
      const int nl = n % nb;
      float acc = 0.f;
      if (accumulate) std::memcpy(&acc, &cm.tmem[ml][dcol + n], 4);
      for (int k = 0; k < 16; ++k) {
        const float b = b_mn ? ld(bs, start(desc_b) + (nl >> 6) * lbo(desc_b) + (k >> 3) * sbo(desc_b) + (k & 7) * 128 + (nl & 63) * 2)
                             : ld(bs, start(desc_b) + (nl >> 3) * sbo(desc_b) + (nl & 7) * 128 + k * 2);
        acc += a[k] * b;
      }
      std::memcpy(&cm.tmem[ml][dcol + n], &acc, 4);
    }
  }
}
inline void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { umma_exec(d, false, 0, da, db, idesc, acc, false); }
inline void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) { umma_exec(d, true, a, 0, db, idesc, acc, false); }
inline void umma_ss_pair(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { umma_exec(d, false, 0, da, db, idesc, acc, true); }
inline void umma_commit(uint64_t* bar) { mbar_arrive(bar); }   // MMAs run synchronously: everything issued has retired
inline void umma_commit_mask(uint64_t* bar, uint16_t mask) {   // arrives at the same offset in every CTA of the mask
  const uint32_t off = smem_u32(bar);
  for (uint32_t d = 0; d < 4; ++d)
    if (mask & (1u << d)) mbar_arrive_addr(mapa_u32(off, d));
}
inline void umma_commit_pair(uint64_t* bar) { umma_commit_mask(bar, (uint16_t)(3u << (emu_tc::cta().rank & ~1u))); }
}  // namespace lp

// ---- the few CUDA built-ins these kernels use beyond cuda_emu.h
inline int __shfl_sync(unsigned, int v, int) { return v; }   // only used to broadcast a warp-uniform value
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
