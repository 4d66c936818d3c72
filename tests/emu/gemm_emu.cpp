// Test infrastructure: the persistent tcgen05 GEMM (leopard_b200/csrc/lp_gemm.cu, gemm_bf16_kernel) run on the CPU under
// tests/emu/tc_emu.h, one CTA per tile and as CTA PAIRS (cluster of 2, cta_group::2).  `gemm_kernel.inc` is the kernel part of
// lp_gemm.cu, cut by the test; three lines are rewritten there: the dynamic shared-memory declaration (per-CTA buffer), the
// epilogue's named-barrier PTX statement and the hardware tanh of the GELU helper.
#include "tc_emu.h"
enum { LP_ACT_NONE = 0, LP_ACT_GELU_TANH = 1, LP_ACT_GELU_ERF = 2, LP_ACT_SWIGLU = 3 };
enum { LP_NORM_NONE = 0, LP_NORM_RMS = 1, LP_NORM_LAYER = 2 };
#include "emu_helpers.inc"
#include "gemm_kernel.inc"
#include "gemm4_kernel.inc"
#include "gemm_chain_kernel.inc"

using namespace lp;

static CUtensorMap tmap2(const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t b0, uint32_t b1, int elem, int swz) {
  CUtensorMap m;
  m.base = reinterpret_cast<const uint8_t*>(base);
  m.dims[0] = inner; m.dims[1] = outer; m.dims[2] = 1;
  m.strides[0] = ld * elem; m.strides[1] = 0;
  m.box[0] = b0; m.box[1] = b1; m.box[2] = 1;
  m.elem_bytes = elem; m.swizzle = swz;
  return m;
}

template <int BN, int STAGES, bool CTA2>
static void run(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmArgs& args, int ctas) {
  const bool lres = args.residual != nullptr;     // launch_gemm's choice: the register-residual kernel when there is a residual
  if constexpr (CTA2) {
    emu_tc::launch_pairs(dim3(ctas), dim3(GEMM_THREADS), [&] {
      if (lres) gemm_bf16_lres_kernel<BN, STAGES, true>(a, b, c, args); else gemm_bf16_kernel<BN, STAGES, true>(a, b, c, args);
    });
  } else {
    emu::g_block_hook = [] { emu_tc::reset(); };
    emu::launch(dim3(ctas), dim3(GEMM_THREADS), [&] {
      if (lres) gemm_bf16_lres_kernel<BN, STAGES, false>(a, b, c, args); else gemm_bf16_kernel<BN, STAGES, false>(a, b, c, args);
    });
    emu::g_block_hook = nullptr;
  }
}

// Same argument preparation as lp_gemm_bf16 (host part of lp_gemm.cu) for an explicit block_n / pair / grid choice.
struct EmuFusions {   // the epilogue fusions of lp_gemm_bf16_ex
  int norm_mode; const float* stats_in; int stats_in_parts; const float* colsum; float norm_eps;
  float* stats_out; const float* rope_cos; const float* rope_sin; int rope_heads;
};

template <int BN, int STAGES>
static void run_res(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& r, const GemmArgs& args, int ctas) {
  emu_tc::launch_pairs(dim3(ctas), dim3(GEMM_THREADS), [&] { gemm_bf16_res_kernel<BN, STAGES, true>(a, b, c, r, args); });
}
static int g_tma_res = 0;   // pair-mode residual launches through gemm_bf16_res_kernel (lp_gemm_set_tma_residual)
extern "C" void emu_gemm_set_tma_residual(int on) { g_tma_res = on; }

// stream-K tail (lp_gemm_set_streamk_workspace): set by emu_gemm_set_streamk before a launch, same split rule as lp_gemm_bf16_ex
static void* g_sk_ws = nullptr;
static int g_sk_pure = 0;     // pure stream-K over a column-block-fastest raster (lp_gemm_set_streamk mode 2, norm-consuming launches)
extern "C" void emu_gemm_set_streamk(void* ws) { g_sk_ws = ws; }
extern "C" void emu_gemm_set_streamk_pure(int on) { g_sk_pure = on; }

extern "C" void emu_gemm_ex(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K,
                            const float* bias, const void* residual, long long ldr, int res_row_mod, int act, int out_fp32,
                            int block_n, int pair, int group_m, int ctas, const EmuFusions* fx) {
  GemmArgs args = {};
  if (fx) {
    args.norm_mode = fx->norm_mode; args.stats_in = reinterpret_cast<const float2*>(fx->stats_in);
    args.stats_in_parts = fx->stats_in_parts; args.colsum = fx->colsum; args.norm_eps = fx->norm_eps;
    args.norm_inv_k = 1.0f / (float)K;
    args.stats_out = reinterpret_cast<float2*>(fx->stats_out);
    args.rope_cos = fx->rope_cos; args.rope_sin = fx->rope_sin; args.rope_heads = fx->rope_heads;
  }
  args.M = M; args.N = N; args.K = K;
  args.C = C; args.ldc = ldc; args.out_fp32 = out_fp32;
  args.bias = bias;
  args.residual = reinterpret_cast<const bf16*>(residual);
  args.ldr = ldr; args.res_row_mod = res_row_mod;
  args.act = act;
  const int bm = pair ? 2 * GEMM_BM : GEMM_BM;
  args.num_m_blk = (M + bm - 1) / bm;
  args.num_n_blk = (N + block_n - 1) / block_n;
  args.num_k_blk = (K + GEMM_BK - 1) / GEMM_BK;
  args.group_m = group_m > args.num_m_blk ? args.num_m_blk : group_m;
  args.prefetch_a = 0;
  if (g_sk_ws != nullptr) {
    const int T = args.num_m_blk * args.num_n_blk, n_units = pair ? ctas / 2 : ctas;
    if (g_sk_pure && T >= n_units) {
      args.group_m = 1;
      args.dp_tiles = 0;
      args.sk_tiles = T;
      args.sk_flags = reinterpret_cast<unsigned*>(g_sk_ws);
      args.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_sk_ws) + 4096);
    } else if (T > n_units && T % n_units != 0) {
      args.dp_tiles = (T / n_units - 1) * n_units;
      args.sk_tiles = T - args.dp_tiles;
      args.sk_flags = reinterpret_cast<unsigned*>(g_sk_ws);
      args.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_sk_ws) + 4096);
    }
  }
  CUtensorMap tmC = tmap2(C, act == LP_ACT_SWIGLU ? N / 2 : N, M, ldc, 32, 32, 2, 64);
  CUtensorMap tmA = tmap2(A, K, M, lda, GEMM_BK, GEMM_BM, 2, 128);
  CUtensorMap tmB = tmap2(W, K, N, ldw, GEMM_BK, pair ? block_n / 2 : block_n, 2, 128);
  if (pair && g_tma_res && residual != nullptr && res_row_mod == 0 && !out_fp32 && act != LP_ACT_SWIGLU && !(fx && fx->rope_heads)) {
    CUtensorMap tmR = tmap2(residual, N, M, ldr, 32, 32, 2, 64);
    switch (block_n) {
      case 256: run_res<256, 5>(tmA, tmB, tmC, tmR, args, ctas); break;
      case 192: run_res<192, 6>(tmA, tmB, tmC, tmR, args, ctas); break;
      default: run_res<128, 7>(tmA, tmB, tmC, tmR, args, ctas); break;
    }
    return;
  }
  if (pair) {
    switch (block_n) {
      case 256: run<256, 6, true>(tmA, tmB, tmC, args, ctas); break;
      case 192: run<192, 7, true>(tmA, tmB, tmC, args, ctas); break;
      default: run<128, 8, true>(tmA, tmB, tmC, args, ctas); break;
    }
  } else {
    switch (block_n) {
      case 256: run<256, 4, false>(tmA, tmB, tmC, args, ctas); break;
      case 160: run<160, 5, false>(tmA, tmB, tmC, args, ctas); break;
      default: run<128, 6, false>(tmA, tmB, tmC, args, ctas); break;
    }
  }
}

extern "C" void emu_gemm(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K,
                         const float* bias, const void* residual, long long ldr, int res_row_mod, int act, int out_fp32,
                         int block_n, int pair, int group_m, int ctas) {
  emu_gemm_ex(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, res_row_mod, act, out_fp32, block_n, pair, group_m, ctas,
              nullptr);
}


// 4-CTA clusters (lp_gemm4.cu): argument preparation as in lp_gemm_bf16's mode-2 branch
template <int BN, int STAGES>
static void run4(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmArgs& args, int ctas) {
  emu_tc::launch_clusters(dim3(ctas), dim3(GEMM_THREADS), 4, [&] { gemm4_bf16_kernel<BN, STAGES>(a, b, c, args); });
}

extern "C" void emu_gemm4(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K,
                          const float* bias, const void* residual, long long ldr, int res_row_mod, int act, int out_fp32,
                          int block_n, int group_m, int ctas) {
  GemmArgs args = {};
  args.M = M; args.N = N; args.K = K;
  args.C = C; args.ldc = ldc; args.out_fp32 = out_fp32;
  args.bias = bias;
  args.residual = reinterpret_cast<const bf16*>(residual);
  args.ldr = ldr; args.res_row_mod = res_row_mod;
  args.act = act;
  args.num_m_blk = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  args.num_n_blk = ((N + block_n - 1) / block_n + 1) / 2;      // column-block PAIRS
  args.num_k_blk = (K + GEMM_BK - 1) / GEMM_BK;
  args.group_m = group_m > args.num_m_blk ? args.num_m_blk : group_m;
  args.prefetch_a = 0;
  CUtensorMap tmC = tmap2(C, act == LP_ACT_SWIGLU ? N / 2 : N, M, ldc, 32, 32, 2, 64);
  CUtensorMap tmAh = tmap2(A, K, M, lda, GEMM_BK, GEMM_BM / 2, 2, 128);
  CUtensorMap tmB = tmap2(W, K, N, ldw, GEMM_BK, block_n / 2, 2, 128);
  switch (block_n) {
    case 256: run4<256, 6>(tmAh, tmB, tmC, args, ctas); break;
    case 192: run4<192, 7>(tmAh, tmB, tmC, args, ctas); break;
    default: run4<128, 8>(tmAh, tmB, tmC, args, ctas); break;
  }
}


// Tensor-parallel launches (lp_gemm_bf16_tp): ONE rank's kernel with the peers' buffers living in this process.  mode 1 = all-gather
// (peer_A[r] = rank r's copy of the A buffer, signal words pre-set by the test = the peers are ready), mode 2 = reduce-scatter
// (stage[r] = rank r's staging [tp, seg_rows, ldc]).  Argument preparation as in lp_gemm_bf16_tp.
extern "C" void emu_gemm_tp(void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K, int act,
                            int tp, int rank, int mode, void* const* peer_A, void* const* stage, unsigned* counter,
                            unsigned* signal_local, unsigned* const* signal_peer, unsigned epoch, int ctas) {
  constexpr int BN = 256;
  GemmArgs args = {};
  args.M = M; args.N = N; args.K = K; args.C = C; args.ldc = ldc; args.act = act;
  args.tp_size = tp; args.tp_rank = rank; args.seg_rows = M / tp;
  args.seg_m_blk = (args.seg_rows + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  args.num_m_blk = tp * args.seg_m_blk;
  args.num_n_blk = (N + BN - 1) / BN;
  args.num_k_blk = (K + GEMM_BK - 1) / GEMM_BK;
  args.group_m = args.seg_m_blk;
  args.epoch = epoch;
  const int n_out = act == LP_ACT_SWIGLU ? N / 2 : N;
  GemmSegMaps segs;
  for (int r = 0; r < 8; ++r) {
    const int rr = r < tp ? r : 0;
    if (mode == 1) {
      args.ag_src[r] = peer_A[rr];
      args.ag_wait[r] = signal_local + rr;
      args.ag_post[r] = signal_peer[rr] + rank;
      segs.m[r] = tmap2(reinterpret_cast<bf16*>(C) + (size_t)rr * args.seg_rows * ldc, n_out, args.seg_rows, ldc, 32, 32, 2, 64);
    } else {
      args.rs_post[r] = signal_peer[rr] + 8 + rank;
      segs.m[r] = tmap2(reinterpret_cast<bf16*>(stage[rr]) + (size_t)rank * args.seg_rows * ldc, n_out, args.seg_rows, ldc, 32, 32, 2, 64);
    }
  }
  if (mode == 1) { args.ag_counter = counter; args.ag_dst = A; args.ag_ld = lda; args.ag_cols = K; }
  else args.rs_counter = counter;
  CUtensorMap tmA = tmap2(A, K, M, lda, GEMM_BK, GEMM_BM, 2, 128);
  CUtensorMap tmB = tmap2(W, K, N, ldw, GEMM_BK, BN / 2, 2, 128);
  emu_tc::launch_pairs(dim3(ctas), dim3(GEMM_THREADS), [&] { gemm_bf16_tp_kernel<BN, 6, true>(tmA, tmB, segs, args); });
}


// Chained projector launch (lp_gemm_chain.cu): argument preparation as in launch_gemm_chain.  The emulator runs CTA pairs one after
// another, so the cross-unit dataflow wait can only be exercised with ONE pair (ctas = 2); the work list, the two sets of tensor
// maps, the per-layer epilogues and the counter protocol (publish / wait / re-zero) are what this checks.
extern "C" void emu_gemm_chain(const void* X, long long ldx, const void* W1, const float* b1, int act1, void* mid, long long ld_mid,
                               const void* W2, const float* b2, void* out, long long ldo, int M, int N1, int K1, int N2,
                               unsigned* counters, int ctas) {
  constexpr int BN = 256;
  ChainArgs a = {};
  a.M = M;
  a.N[0] = N1; a.K[0] = K1; a.N[1] = N2; a.K[1] = N1;
  a.num_m_blk = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  for (int l = 0; l < 2; ++l) {
    a.num_n_blk[l] = (a.N[l] + BN - 1) / BN;
    a.num_k_blk[l] = (a.K[l] + GEMM_BK - 1) / GEMM_BK;
  }
  a.bias[0] = b1; a.bias[1] = b2;
  a.act[0] = act1; a.act[1] = LP_ACT_NONE;
  a.row_done = counters + 1;
  a.exit_count = counters;
  CUtensorMap tmA1 = tmap2(X, K1, M, ldx, GEMM_BK, GEMM_BM, 2, 128);
  CUtensorMap tmB1 = tmap2(W1, K1, N1, K1, GEMM_BK, BN / 2, 2, 128);
  CUtensorMap tmC1 = tmap2(mid, N1, M, ld_mid, 32, 32, 2, 64);
  CUtensorMap tmA2 = tmap2(mid, N1, M, ld_mid, GEMM_BK, GEMM_BM, 2, 128);
  CUtensorMap tmB2 = tmap2(W2, N1, N2, N1, GEMM_BK, BN / 2, 2, 128);
  CUtensorMap tmC2 = tmap2(out, N2, M, ldo, 32, 32, 2, 64);
  emu_tc::launch_pairs(dim3(ctas), dim3(GEMM_THREADS), [&] { gemm_chain_kernel<BN, 6>(tmA1, tmB1, tmC1, tmA2, tmB2, tmC2, a); });
}


// Transposed operands (lp_gemm_desc a_t / b_t): A given as At [K, M], W as Wt [K, N]; pair mode, 256-wide tiles, plain epilogue.
// Argument preparation as in lp_gemm_bf16_ex's a_t / b_t branch.
extern "C" void emu_gemm_mn(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K,
                            const float* bias, int out_fp32, int a_t, int b_t, int ctas) {
  constexpr int BN = 256;
  GemmArgs args = {};
  args.M = M; args.N = N; args.K = K; args.C = C; args.ldc = ldc; args.out_fp32 = out_fp32; args.bias = bias; args.act = LP_ACT_NONE;
  args.num_m_blk = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  args.num_n_blk = (N + BN - 1) / BN;
  args.num_k_blk = (K + GEMM_BK - 1) / GEMM_BK;
  args.group_m = args.num_m_blk;
  if (g_sk_ws != nullptr) {
    const int T = args.num_m_blk * args.num_n_blk, n_units = ctas / 2;
    if (T > n_units && T % n_units != 0) {
      args.dp_tiles = (T / n_units - 1) * n_units;
      args.sk_tiles = T - args.dp_tiles;
      args.sk_flags = reinterpret_cast<unsigned*>(g_sk_ws);
      args.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_sk_ws) + 4096);
    }
  }
  CUtensorMap tmC = tmap2(C, N, M, ldc, 32, 32, 2, 64);
  CUtensorMap tmA = a_t ? tmap2(A, M, K, lda, 64, GEMM_BK, 2, 128) : tmap2(A, K, M, lda, GEMM_BK, GEMM_BM, 2, 128);
  CUtensorMap tmB = b_t ? tmap2(W, N, K, ldw, 64, GEMM_BK, 2, 128) : tmap2(W, K, N, ldw, GEMM_BK, BN / 2, 2, 128);
  auto go = [&](auto kern) { emu_tc::launch_pairs(dim3(ctas), dim3(GEMM_THREADS), [&] { kern(tmA, tmB, tmC, args); }); };
  if (a_t && b_t) go(gemm_bf16_mn_kernel<BN, 6, true, true>);
  else if (a_t) go(gemm_bf16_mn_kernel<BN, 6, true, false>);
  else go(gemm_bf16_mn_kernel<BN, 6, false, true>);
}
