// Test infrastructure: runs the two kernels of leopard_b200/csrc/lp_attn_bwd.cu on the CPU under tests/emu/tc_emu.h.
#define LP_EMULATE 1
#include "../../leopard_b200/csrc/lp_attn_bwd.cu"

using namespace lp;

static CUtensorMap tmap3(const void* base, uint64_t ld, uint64_t T, uint64_t B) {
  CUtensorMap m;
  m.base = reinterpret_cast<const uint8_t*>(base);
  m.dims[0] = ld; m.dims[1] = T; m.dims[2] = B;
  m.strides[0] = ld * 2; m.strides[1] = T * ld * 2;
  m.box[0] = 64; m.box[1] = 128; m.box[2] = 1;
  return m;
}

template <int DH, bool CAUSAL>
static void run(const CUtensorMap& tq, const CUtensorMap& td, const AttnBwdArgs& a, int n_kv_heads, int which) {
  const int tiles = (a.T + AB_T - 1) / AB_T;
  emu::g_block_hook = [] { emu_tc::reset(); };
  if (which & 1) emu::launch(dim3(a.n_heads, tiles, a.B), dim3(AB_THREADS), [&] { attn_bwd_dq_kernel<DH, CAUSAL>(tq, td, a); });
  if (which & 2) emu::launch(dim3(n_kv_heads, tiles, a.B), dim3(AB_THREADS), [&] { attn_bwd_dkv_kernel<DH, CAUSAL>(tq, td, a); });
  emu::g_block_hook = nullptr;
}

extern "C" void emu_attention_bwd(const void* qkv, long long ld, const void* out, const void* dout, long long ldo, int B, int T,
                                  int n_heads, int n_kv_heads, int head_dim, int head_stride, int q_off, int k_off, int v_off,
                                  int causal, float scale, void* dqkv, long long ldd, float* lse2, float* delta, int which) {
  AttnBwdArgs a;
  a.T = T; a.B = B; a.n_heads = n_heads; a.group = n_heads / n_kv_heads;
  a.Tpad = (T + AB_T - 1) / AB_T * AB_T;
  a.q_off = q_off; a.k_off = k_off; a.v_off = v_off;
  a.hs = head_stride; a.dho = head_dim;
  a.O = reinterpret_cast<const bf16*>(out);
  a.dO = reinterpret_cast<const bf16*>(dout);
  a.ldo = ldo;
  a.dqkv = reinterpret_cast<bf16*>(dqkv);
  a.ldd = ldd;
  a.lse2 = lse2; a.delta = delta;
  a.scale = scale; a.scale_log2 = scale * 1.4426950408889634f;
  const CUtensorMap tq = tmap3(qkv, ld, T, B), td = tmap3(dout, ldo, T, B);
  if (head_dim == 72) run<80, false>(tq, td, a, n_kv_heads, which);
  else if (causal) run<128, true>(tq, td, a, n_kv_heads, which);
  else run<128, false>(tq, td, a, n_kv_heads, which);
}
