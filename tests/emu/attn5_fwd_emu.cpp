// Test infrastructure: the persistent short-sequence attention kernel (leopard_b200/csrc/lp_attn5.cu, attn5_fwd_kernel) on the CPU
// under tests/emu/tc_emu.h; `grid` CTAs walk the (batch, head, query-tile pair) items, so a small grid exercises the item loop,
// the barrier parities that run across items, Q replacement (q_empty) and the accumulator hand-back (o_free).
#include "tc_emu.h"
#include "emu_helpers.inc"

namespace lp {
#include "attn_args.inc"
constexpr int A5_BOX = 128 * 128;
constexpr int A5_THREADS = 384;
inline float a5_ex2(float x) { return exp2f(x); }
inline float a5_ex2_fma(float x) {   // same arithmetic as the kernel file's FMA-pipe exp2
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05459282547f, 0.24221783876f);
  p = fmaf(p, f, 0.69336861372f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
inline void a5_bar_sync(int id, int n) { named_bar(id, n, true); }
inline void a5_bar_arrive(int id, int n) { named_bar(id, n, false); }
inline uint64_t a5_pk2(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }
inline void a5_upk2(uint64_t r, float& a, float& b) { a = __uint_as_float((uint32_t)r); b = __uint_as_float((uint32_t)(r >> 32)); }
inline uint64_t a5_fma2(uint64_t a, uint64_t b, uint64_t c) {
  float a0, a1, b0, b1, c0, c1;
  a5_upk2(a, a0, a1); a5_upk2(b, b0, b1); a5_upk2(c, c0, c1);
  return a5_pk2(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
}
inline uint64_t a5_add2(uint64_t a, uint64_t b) {
  float a0, a1, b0, b1;
  a5_upk2(a, a0, a1); a5_upk2(b, b0, b1);
  return a5_pk2(a0 + b0, a1 + b1);
}
#define LP5_TMEM_ST_16(regs, taddr) lp::tmem_st(regs, taddr, 16)
#include "attn5_kernel.inc"
}  // namespace lp

using namespace lp;

static void run(const CUtensorMap& tm, const AttnArgs& a, int grid) {
  const int n_pairs = (a.T + 255) / 256, n_items = a.B * a.n_heads * n_pairs;
  emu::g_block_hook = [] { emu_tc::reset(); };
  emu::launch(dim3(grid < n_items ? grid : n_items), dim3(A5_THREADS), [&] { attn5_fwd_kernel<80>(tm, a, n_pairs, n_items); });
  emu::g_block_hook = nullptr;
}

extern "C" void emu_attention5_fwd(const void* qkv, long long ld, int B, int T, int n_heads, int n_kv_heads, int head_dim,
                                   int head_stride, int q_off, int k_off, int v_off, int causal, float scale, void* out,
                                   long long ldo, int pingpong, int grid) {
  AttnArgs a;
  a.T = T; a.Tk = T; a.B = B; a.n_heads = n_heads; a.group = n_heads / n_kv_heads;
  a.q_off = q_off; a.k_off = k_off; a.v_off = v_off; a.hs = head_stride;
  a.dh_out = head_dim;
  a.O = reinterpret_cast<bf16*>(out);
  a.ldo = ldo;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.pingpong = pingpong;
  CUtensorMap tm;
  tm.base = reinterpret_cast<const uint8_t*>(qkv);
  tm.dims[0] = ld; tm.dims[1] = T; tm.dims[2] = B;
  tm.strides[0] = ld * 2; tm.strides[1] = (uint64_t)T * ld * 2;
  tm.box[0] = 64; tm.box[1] = 128; tm.box[2] = 1;
  (void)causal;
  run(tm, a, grid);
}
