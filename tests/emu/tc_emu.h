// Test infrastructure: functional stand-ins for the tcgen05 / TMEM / TMA / mbarrier wrappers of leopard_b200/csrc/lp_common.cuh,
// on top of cuda_emu.h, so that the SOURCE of lp_attn_bwd.cu can run on the CPU (tests/test_attention_kernels_emulated.py).
// What is modelled — and only as far as the kernels rely on it:
//   mbarrier   arrival count + transaction bytes + phase parity; try_wait(parity) is true once the phase of that parity has
//              completed (a fresh barrier passes parity 1); a wait that makes no progress for EMU_DEADLOCK_MS aborts the run
//   TMA        3-D tiled bf16 loads with 128-byte swizzle (16-byte chunk index XOR row-within-8), zero fill out of bounds
//   UMMA       kind::f16 bf16 x bf16 -> fp32, M = 128, K = 16 per instruction, executed synchronously in issue order;
//              A from shared memory (K-major) or from TMEM (packed bf16 pairs), B K-major or MN-major, both through the
//              64-bit shared-memory descriptors (start, LBO, SBO) with the swizzle applied to the absolute address bits
//   TMEM       128 lanes x 512 columns of 32-bit cells; 32x32b loads / stores address lane = base lane + lane id
// The layouts are the author's reading of the hardware, the same reading the GPU-verified forward kernels are built on; this
// emulator checks the kernels' LOGIC (protocol, indexing, masks, operand roles), not the hardware.
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
};

namespace lp {
typedef __nv_bfloat16 bf16;
// the kernels declare `extern __shared__ uint8_t smem_raw[];` at block scope inside namespace lp: that names lp::smem_raw
alignas(1024) uint8_t smem_raw[232448];

namespace emu_tc {
struct Mbar { int expected = 0, pending = 0; long tx = 0; int phase = 0; };
inline std::mutex g_mu;
inline std::unordered_map<uint32_t, Mbar> g_bars;
inline uint32_t g_tmem[128][512];
inline std::unique_ptr<std::barrier<>> g_named;   // named barrier 1 of the kernels: the 128 row threads
struct NamedBar { int count = 0; long gen = 0; };
inline NamedBar g_nbar[16];                        // bar.sync / bar.arrive id, n: completes when n threads have arrived
inline void reset() {
  std::lock_guard<std::mutex> l(g_mu);
  g_bars.clear();
  for (auto& b : g_nbar) b = NamedBar();
  g_named = std::make_unique<std::barrier<>>(128);
  std::memset(g_tmem, 0xff, sizeof(g_tmem));     // NaN pattern: reading an accumulator that was never written shows up
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
inline uint32_t swz(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }
}  // namespace emu_tc

inline uint32_t smem_u32(const void* p) { return (uint32_t)((const uint8_t*)p - smem_raw); }

// ---- mbarrier
inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  emu_tc::Mbar b;
  b.expected = b.pending = (int)count;
  emu_tc::g_bars[smem_u32(bar)] = b;
}
inline void fence_barrier_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_arrive(uint64_t* bar) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  auto& b = emu_tc::g_bars.at(smem_u32(bar));
  b.pending -= 1;
  if (b.pending < 0) { std::fprintf(stderr, "emu: mbarrier over-arrival @%u\n", smem_u32(bar)); std::abort(); }
  emu_tc::maybe_complete(b);
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  auto& b = emu_tc::g_bars.at(smem_u32(bar));
  b.tx += bytes;
  b.pending -= 1;
  emu_tc::maybe_complete(b);
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  auto& b = emu_tc::g_bars.at(smem_u32(bar));
  b.tx -= bytes;
  emu_tc::maybe_complete(b);
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  std::lock_guard<std::mutex> l(emu_tc::g_mu);
  return (uint32_t)emu_tc::g_bars.at(smem_u32(bar)).phase != parity;
}
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
    auto& b = emu_tc::g_nbar[id];
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
      if (emu_tc::g_nbar[id].gen != my_gen) return;
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

// ---- TMA
inline void tma_prefetch_desc(const CUtensorMap*) {}
inline void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  const uint32_t dst = smem_u32(smem_dst);
  for (uint32_t r = 0; r < m->box[1]; ++r)
    for (uint32_t e = 0; e < m->box[0]; ++e) {
      const long col = (long)c0 + e, row = (long)c1 + r;
      uint16_t v = 0;
      if (col >= 0 && row >= 0 && (uint64_t)col < m->dims[0] && (uint64_t)row < m->dims[1] && (uint64_t)c2 < m->dims[2])
        std::memcpy(&v, m->base + col * 2 + row * m->strides[0] + (uint64_t)c2 * m->strides[1], 2);
      std::memcpy(smem_raw + emu_tc::swz(dst + r * 128 + e * 2), &v, 2);
    }
  mbar_complete_tx(bar, m->box[0] * m->box[1] * 2);
}

// ---- TMEM
inline void tmem_alloc(uint32_t* smem_dst, uint32_t) { *smem_dst = 0; }
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void tmem_wait_ld() {}
inline void tmem_wait_st() {}
inline void tmem_ld(uint32_t* regs, uint32_t taddr, int n) {
  const uint32_t lane = (taddr >> 16) + (::emu::linear_tid() & 31), col = taddr & 0xffff;
  for (int i = 0; i < n; ++i) regs[i] = emu_tc::g_tmem[lane][col + i];
}
inline void tmem_st(const uint32_t* regs, uint32_t taddr, int n) {
  const uint32_t lane = (taddr >> 16) + (::emu::linear_tid() & 31), col = taddr & 0xffff;
  for (int i = 0; i < n; ++i) emu_tc::g_tmem[lane][col + i] = regs[i];
}
#define LP_TMEM_LD_32(regs, taddr) lp::tmem_ld(regs, taddr, 32)
#define LP_TMEM_LD_16(regs, taddr) lp::tmem_ld(regs, taddr, 16)
#define LP_TMEM_ST_16(regs, taddr) lp::tmem_st(regs, taddr, 16)

// ---- UMMA (descriptor bit layouts: see umma_desc_sw128 / umma_idesc_bf16, taken from lp_common.cuh by the test)
inline void umma_exec(uint32_t tmem_d, bool a_in_tmem, uint32_t tmem_a, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                      uint32_t accumulate) {
  const int M = (int)((idesc >> 24) & 0x1f) << 4, N = (int)((idesc >> 17) & 0x3f) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (M != 128 || a_mn) { std::fprintf(stderr, "emu: unsupported UMMA shape M=%d a_mn=%d\n", M, (int)a_mn); std::abort(); }
  auto start = [](uint64_t d) { return (uint32_t)(d & 0x3FFF) << 4; };
  auto lbo = [](uint64_t d) { return (uint32_t)((d >> 16) & 0x3FFF) << 4; };
  auto sbo = [](uint64_t d) { return (uint32_t)((d >> 32) & 0x3FFF) << 4; };
  auto ld = [](uint32_t addr) {
    uint16_t h;
    std::memcpy(&h, smem_raw + emu_tc::swz(addr), 2);
    return emu_tc::bf_bits(h);
  };
  const uint32_t dcol = tmem_d & 0xffff, acol = tmem_a & 0xffff;
  for (int m = 0; m < M; ++m) {
    float a[16];
    for (int k = 0; k < 16; ++k) {
      if (a_in_tmem) {
        const uint32_t w = emu_tc::g_tmem[m][acol + (k >> 1)];
        a[k] = emu_tc::bf_bits((uint16_t)((k & 1) ? (w >> 16) : (w & 0xffff)));
      } else {
        a[k] = ld(start(desc_a) + (m >> 3) * sbo(desc_a) + (m & 7) * 128 + k * 2);
      }
    }
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      if (accumulate) std::memcpy(&acc, &emu_tc::g_tmem[m][dcol + n], 4);
      for (int k = 0; k < 16; ++k) {
        const float b = b_mn ? ld(start(desc_b) + (n >> 6) * lbo(desc_b) + (k >> 3) * sbo(desc_b) + (k & 7) * 128 + (n & 63) * 2)
                             : ld(start(desc_b) + (n >> 3) * sbo(desc_b) + (n & 7) * 128 + k * 2);
        acc += a[k] * b;
      }
      std::memcpy(&emu_tc::g_tmem[m][dcol + n], &acc, 4);
    }
  }
}
inline void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) { umma_exec(d, false, 0, da, db, idesc, acc); }
inline void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) { umma_exec(d, true, a, 0, db, idesc, acc); }
inline void umma_commit(uint64_t* bar) { mbar_arrive(bar); }   // MMAs run synchronously: everything issued has retired
}  // namespace lp

// ---- the few CUDA built-ins these kernels use beyond cuda_emu.h
inline int __shfl_sync(unsigned, int v, int) { return v; }   // only used to broadcast a warp-uniform value
inline void __syncwarp() {}
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
