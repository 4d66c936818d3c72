// Test infrastructure: a minimal thread-per-CUDA-thread emulator so that the SOURCE of the element-wise / reduction
// kernels in leopard_b200/csrc/lp_train.cu can be compiled with g++ and executed on the CPU (tests/test_train_kernels_emulated.py).
// It emulates exactly what those kernels use: threadIdx / blockIdx / blockDim / gridDim, __shared__ (function-local statics;
// blocks run one after another), __syncthreads, __shfl_xor_sync, atomicAdd(float), bf16 types with round-to-nearest-even,
// the vector types and a few math intrinsics.  Nothing here is part of the product; nothing tcgen05 / TMA can run here.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
inline float2 make_float2(float a, float b) { return {a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

// ---- bf16 ------------------------------------------------------------------------------------------------------------
struct __nv_bfloat16;
inline __nv_bfloat16 __float2bfloat16(float f);
struct __nv_bfloat16 {
  uint16_t bits;
  __nv_bfloat16() = default;
  explicit __nv_bfloat16(float f);
  operator float() const {
    uint32_t u = (uint32_t)bits << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }
};
inline __nv_bfloat16 __float2bfloat16(float f) {   // round to nearest even, NaN kept quiet
  uint32_t u;
  std::memcpy(&u, &f, 4);
  __nv_bfloat16 r;
  if ((u & 0x7fffffffu) > 0x7f800000u) { r.bits = 0x7fff; return r; }
  u += 0x7fffu + ((u >> 16) & 1u);
  r.bits = (uint16_t)(u >> 16);
  return r;
}
inline __nv_bfloat16::__nv_bfloat16(float f) { bits = __float2bfloat16(f).bits; }
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline __nv_bfloat162 __floats2bfloat162_rn(float lo, float hi) { return {__float2bfloat16(lo), __float2bfloat16(hi)}; }
inline float2 __bfloat1622float2(__nv_bfloat162 v) { return {(float)v.x, (float)v.y}; }

// ---- math ------------------------------------------------------------------------------------------------------------
inline float __expf(float x) { return std::exp(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
using std::min;
using std::max;

// ---- block / warp synchronisation ----------------------------------------------------------------------------------------
namespace emu {
struct BlockCtx {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<float> warp_buf;   // [warps][32]
  // shuffles with a partial member mask synchronise only the lanes named in the mask (one barrier per (warp, mask))
  std::mutex sub_mu;
  std::map<std::pair<unsigned, unsigned>, std::unique_ptr<std::barrier<>>> sub_bar;
  std::barrier<>& group(unsigned warp, unsigned mask) {
    if (mask == 0xffffffffu) return *warp_bar[warp];
    std::lock_guard<std::mutex> l(sub_mu);
    auto& b = sub_bar[{warp, mask}];
    if (!b) b = std::make_unique<std::barrier<>>(__builtin_popcount(mask));
    return *b;
  }
};
inline thread_local BlockCtx* t_ctx = nullptr;   // the block this OS thread belongs to
inline std::function<void()> g_block_hook;   // run before each block's threads start (per-CTA state reset)

// One pool of `nthreads` OS threads walks the grid block by block (blocks run one after another); per block, thread 0 rebuilds
// the block context (fresh __syncthreads / warp barriers) between two pool-wide barriers.
template <typename F>
void launch(dim3 grid, dim3 block, F&& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  if (std::getenv("EMU_TRACE")) std::fprintf(stderr, "emu: launch grid (%u,%u,%u) block %u\n", grid.x, grid.y, grid.z, nthreads);
  std::barrier<> pool_bar(nthreads);
  std::unique_ptr<BlockCtx> ctx;
  std::vector<std::thread> ts;
  ts.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t)
    ts.emplace_back([&, t] {
      threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      blockDim = block;
      gridDim = grid;
      for (unsigned long long b = 0; b < nblocks; ++b) {
        pool_bar.arrive_and_wait();              // every thread has left the previous block
        if (t == 0) {
          ctx = std::make_unique<BlockCtx>();
          ctx->block_bar = std::make_unique<std::barrier<>>(nthreads);
          for (unsigned w = 0; w < nwarps; ++w)
            ctx->warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32u, nthreads - w * 32)));
          ctx->warp_buf.assign(nwarps * 32, 0.f);
          if (g_block_hook) g_block_hook();
        }
        pool_bar.arrive_and_wait();              // the block's context is ready
        t_ctx = ctx.get();
        blockIdx = {(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y))};
        body();
        ctx->block_bar->arrive_and_drop();       // a thread that has returned no longer takes part in __syncthreads
      }
    });
  for (auto& th : ts) th.join();
}
inline unsigned linear_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
}  // namespace emu

inline void __syncthreads() { emu::t_ctx->block_bar->arrive_and_wait(); }
inline float __shfl_xor_sync(unsigned mask, float v, int lane_mask) {
  const unsigned t = emu::linear_tid(), w = t >> 5, l = t & 31;
  float* buf = emu::t_ctx->warp_buf.data() + w * 32;
  std::barrier<>& bar = emu::t_ctx->group(w, mask);
  buf[l] = v;
  bar.arrive_and_wait();
  const float r = buf[l ^ (unsigned)lane_mask];
  bar.arrive_and_wait();
  return r;
}
inline bool __any_sync(unsigned, bool pred) {
  const unsigned t = emu::linear_tid(), w = t >> 5, l = t & 31;
  float* buf = emu::t_ctx->warp_buf.data() + w * 32;
  buf[l] = pred ? 1.f : 0.f;
  emu::t_ctx->warp_bar[w]->arrive_and_wait();
  bool any = false;
  for (int i = 0; i < 32; ++i) any |= buf[i] != 0.f;
  emu::t_ctx->warp_bar[w]->arrive_and_wait();
  return any;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_ctx->warp_bar[emu::linear_tid() >> 5]->arrive_and_wait(); }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
