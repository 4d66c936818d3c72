// leopard_b200 — flash attention v4: the v3 kernel (lp_attn3.cu: two 128-row query tiles per CTA, S delivered in 64-key halves with
// their own TMEM columns, one MMA-issuing thread per query tile, lazy O rescale) with ONE softmax warpgroup for both tiles: thread
// r owns row r of tile A and row r of tile B and runs their exponentials as two independent streams in the same instruction
// window.  Same contract as attn3_fwd_kernel (SigLIP non-causal 16x72 with DH = 80 padded, decoder causal GQA 32/8x128; replaces
// flash_attn_varlen_func at MP/model/llava/transformer.py:476-515,869-877 and TEDotProductAttention at clip_encoder.py:58-63).
//
// Why (tools/micro/mufu_bench.cu, profiles/r02t_mufu_bench.jsonl): the softmax stream of one query tile (fma2 -> 2 x ex2 -> add2 ->
// pack per pair of S values) issues a MUFU.EX2 only every 14.6 cycles when ONE warp per SM sub-partition runs it — the warp's
// in-order issue stalls on the consumers of earlier exponentials — while the unit accepts one every 8.  v3's turn protocol keeps the
// two tiles' exp sections apart, so each runs at that one-warp rate (930 cycles per 64 values; the ncu profile's "XU pipe 55 % busy
// inside exp sections").  Two independent streams in one warp reach 9.7 cycles per MUFU (1244 cycles per 2 x 64 values instead of
// 2 x 930): the exp work of a key half-step of both tiles shrinks by a third, without named-barrier hand-overs.
//   warp 0      TMA producer : Q_A, Q_B once; K_j, V_j (128 rows) into 2-deep rings shared by both tiles
//   warp 1 / 2  MMA issuer of tile A / B (unchanged from v3)
//   warp 3      TMEM allocation
//   warps 4-7   softmax of BOTH tiles: one thread per row pair
// Causal pairs: tile A has one key tile fewer than tile B; the steps where only one tile is active run the single-stream body.
// RESULT (profiles/r02v_attn4_time.log): bit-identical to v3 but 23-30 % SLOWER (decoder S = 4796: 693 vs 905 TFLOP/s) — with one warp
// per sub-partition nothing hides the barrier / TMEM round trips between the exp sections (the phase trace of v3 later showed
// that those, not the exponentials, are most of a half step).  Kept as variant bit 7 (128) for A/B and as the cross-check that
// pinned v3's arithmetic while v3 itself was being restructured.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

constexpr int A4_BOX = 128 * 128;  // one TMA box: 128 rows x 64 bf16
constexpr int A4_THREADS = 256;

__device__ __forceinline__ float a4_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], max rel. error 1.1e-4 — below the
// bf16 rounding of P): takes a share of the exponentials off the 16-lane/clk MUFU unit, which bounds the exp
// sections of this kernel (ncu: XU pipe ~75 % busy inside them).
__device__ __forceinline__ float a4_ex2_fma(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.0f);           // f in [-0.5, 0.5]
  float p = fmaf(f, 0.05459282547f, 0.24221783876f);
  p = fmaf(p, f, 0.69336861372f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));  // * 2^round(x)
}
__device__ __forceinline__ void a4_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void a4_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint64_t a4_pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void a4_upk2(uint64_t r, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ uint64_t a4_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t a4_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

#define LP4_TMEM_ST_16(regs, taddr)                                                                                 \
  asm volatile(                                                                                                     \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, " \
      "%15, %16};" ::"r"(taddr),                                                                                    \
      "r"(regs[0]), "r"(regs[1]), "r"(regs[2]), "r"(regs[3]), "r"(regs[4]), "r"(regs[5]), "r"(regs[6]),             \
      "r"(regs[7]), "r"(regs[8]), "r"(regs[9]), "r"(regs[10]), "r"(regs[11]), "r"(regs[12]), "r"(regs[13]),         \
      "r"(regs[14]), "r"(regs[15])                                                                                  \
      : "memory")

// ---- kernel code (no inline PTX below this line: tests/emu cuts it out verbatim)
// One 64-key half step of this thread's row in NQ query tiles (stream s works on tile q0 + s): S -> registers, row max, lazy O
// rescale (threshold 2^8), exponentials of all NQ rows as independent instruction streams in one window, P -> bf16 -> TMEM over
// the first 32 columns of its own S half, then the hand-over to the tile's MMA issuer.  (m0, l0) / (m1, l1): running maximum (log2
// domain) and fp32 row sum of streams 0 / 1.  t = half steps this tile has published so far.
template <int DH, bool CAUSAL, int EMU, int NQ>
__device__ __forceinline__ void a4_half_step(const AttnArgs& args, uint64_t* s_full, uint64_t* p_full, uint64_t* pv_done,
                                             uint32_t lane_addr, int q0, int qp, int j, int h, int T, int r, int t, float& m0,
                                             float& l0, float& m1, float& l1) {
  constexpr uint32_t TM_S = 0, TM_O = 256;
  const int kv0 = j * 128 + h * 64;
  uint32_t v[NQ][64];
#pragma unroll
  for (int s = 0; s < NQ; ++s) {
    const int q = q0 + s;
    mbar_wait(&s_full[q * 2 + h], j & 1);
    tc_fence_after();
    LP_TMEM_LD_32((v[s] + 0), lane_addr + TM_S + (uint32_t)q * 128 + h * 64 + 0);
    LP_TMEM_LD_32((v[s] + 32), lane_addr + TM_S + (uint32_t)q * 128 + h * 64 + 32);
  }
  tmem_wait_ld();
  bool pv_waited[NQ];
  uint64_t nm2[NQ];
#pragma unroll
  for (int s = 0; s < NQ; ++s) {
    const int q = q0 + s;
    float& m_used = (s == 0) ? m0 : m1;
    float& l_sum = (s == 0) ? l0 : l1;
    const int diag = 2 * qp + q;  // kv tile that holds this q tile's diagonal
    const bool need_mask = (kv0 + 64 > T) || (CAUSAL && j == diag);
    int lim = T - kv0;
    if (CAUSAL && j == diag) lim = min(lim, r + 1 - h * 64);
    if (need_mask) {
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i >= lim) v[s][i] = 0xff800000u;  // -inf
    }
    float mx;
    {
      float x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY, x3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        x0 = fmaxf(x0, fmaxf(__uint_as_float(v[s][i]), __uint_as_float(v[s][i + 1])));
        x1 = fmaxf(x1, fmaxf(__uint_as_float(v[s][i + 2]), __uint_as_float(v[s][i + 3])));
        x2 = fmaxf(x2, fmaxf(__uint_as_float(v[s][i + 4]), __uint_as_float(v[s][i + 5])));
        x3 = fmaxf(x3, fmaxf(__uint_as_float(v[s][i + 6]), __uint_as_float(v[s][i + 7])));
      }
      mx = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
    }
    const float m_new = fmaxf(m_used, mx * args.scale_log2);
    const bool grow = (m_new - m_used) > 8.0f;
    // pv_done_q completes one phase per P·V half; phase t-1 is consumed exactly once during half step t — early when O has to be
    // rescaled, otherwise just before P(t) is published (by then it has normally retired)
    pv_waited[s] = (t == 0);
    if (t == 0) {
      m_used = m_new;
    } else if (__any_sync(0xffffffffu, grow)) {
      // every P·V issued so far (halves 0 .. t-1) must have retired before O_q is rescaled in place
      mbar_wait(&pv_done[q], (t - 1) & 1);
      pv_waited[s] = true;
      tc_fence_after();
      float alpha = 1.0f;
      if (grow) {
        alpha = a4_ex2(m_used - m_new);
        m_used = m_new;
        l_sum *= alpha;
      }
      const uint32_t o_addr = lane_addr + TM_O + (uint32_t)q * 128;
#pragma unroll 1
      for (int c = 0; c < DH / 16; ++c) {
        uint32_t o[16];
        LP_TMEM_LD_16(o, o_addr + c * 16);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        LP4_TMEM_ST_16(o, o_addr + c * 16);
      }
      tmem_wait_st();
    }
    const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
    nm2[s] = a4_pk2(neg_m, neg_m);
  }
  const uint64_t sc2 = a4_pk2(args.scale_log2, args.scale_log2);
  uint64_t acc2[NQ];
#pragma unroll
  for (int s = 0; s < NQ; ++s) acc2[s] = a4_pk2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t pk[NQ][16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
#pragma unroll
      for (int s = 0; s < NQ; ++s) {
        float x0, x1;
        a4_upk2(a4_fma2(a4_pk2(__uint_as_float(v[s][c * 32 + i]), __uint_as_float(v[s][c * 32 + i + 1])), sc2, nm2[s]), x0, x1);
        // compile-time pattern: element (i & 3) < EMU of each group of four goes to the FMA pipe
        const float p0 = ((i & 3) < EMU) ? a4_ex2_fma(x0) : a4_ex2(x0);
        const float p1 = (((i + 1) & 3) < EMU) ? a4_ex2_fma(x1) : a4_ex2(x1);
        // row sum in fp32 before the bf16 rounding of P (as flash-attention does)
        acc2[s] = a4_add2(acc2[s], a4_pk2(p0, p1));
        pk[s][i >> 1] = pack_bf16x2(p0, p1);
      }
    }
#pragma unroll
    for (int s = 0; s < NQ; ++s)
      LP4_TMEM_ST_16(pk[s], lane_addr + TM_S + (uint32_t)(q0 + s) * 128 + h * 64 + c * 16);
  }
  tmem_wait_st();
  tc_fence_before();
#pragma unroll
  for (int s = 0; s < NQ; ++s) {
    const int q = q0 + s;
    if (!pv_waited[s]) mbar_wait(&pv_done[q], (t - 1) & 1);
    mbar_arrive(&p_full[q * 2 + h]);
    float a0, a1;
    a4_upk2(acc2[s], a0, a1);
    ((s == 0) ? l0 : l1) += a0 + a1;
  }
}

// EMU: how many of every 4 exponentials are evaluated with a4_ex2_fma instead of MUFU.EX2 (0, 1 or 2)
template <int DH, bool CAUSAL, int EMU>
__global__ void __launch_bounds__(A4_THREADS, 1)
attn4_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  constexpr int NB = (DH + 63) / 64;
  constexpr int KS_QK = DH / 16;
  constexpr int HEAD_BYTES = NB * A4_BOX;
  constexpr int Q_OFF = 0;                       // 2 tiles
  constexpr int K_OFF = Q_OFF + 2 * HEAD_BYTES;  // 2 stages
  constexpr int V_OFF = K_OFF + 2 * HEAD_BYTES;  // 2 stages
  constexpr int BAR_OFF = V_OFF + 2 * HEAD_BYTES;
  constexpr uint32_t TM_S = 0, TM_O = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [2]   per q tile
  uint64_t* k_full = q_full + 2;                                    // [2]   per stage
  uint64_t* v_full = k_full + 2;                                    // [2]
  uint64_t* k_empty = v_full + 2;                                   // [4]   q tile x stage
  uint64_t* v_empty = k_empty + 4;                                  // [4]
  uint64_t* s_full = v_empty + 4;                                   // [4]   q tile x half
  uint64_t* p_full = s_full + 4;                                    // [4]
  uint64_t* pv_done = p_full + 4;                                   // [2]   one phase per P·V half
  uint64_t* o_done = pv_done + 2;                                   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int qp = (int)gridDim.x - 1 - (int)blockIdx.x;  // heavy (late) causal pairs first
  const int head = blockIdx.y, batch = blockIdx.z;
  const int kv_head = head / args.group;
  const int T = args.T;  // self-attention only: Tk == T
  const int q0A = qp * 256, q0B = q0A + 128;
  const int n_total = (T + 127) / 128;
  const bool actB = q0B < T;
  const int nA = CAUSAL ? min(n_total, 2 * qp + 1) : n_total;
  const int nB = actB ? (CAUSAL ? min(n_total, 2 * qp + 2) : n_total) : 0;
  const int n_kv = max(nA, nB);

  // K/V tile loader (producer thread only): tile j -> ring slot j & 1; slot free once BOTH tiles' issuers released it
  auto load_kv = [&](int j) {
    const int st = j & 1;
    const uint32_t ph = (j >> 1) & 1;
    if (j >= 2) {
      if (j - 2 < nA) mbar_wait(&k_empty[0 * 2 + st], ph ^ 1);
      if (j - 2 < nB) mbar_wait(&k_empty[1 * 2 + st], ph ^ 1);
    }
    mbar_arrive_expect_tx(&k_full[st], HEAD_BYTES);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      tma_load_3d(smem + K_OFF + st * HEAD_BYTES + b * A4_BOX, &tmQKV, &k_full[st],
                  args.k_off + kv_head * args.hs + b * 64, j * 128, batch);
    if (j >= 2) {
      if (j - 2 < nA) mbar_wait(&v_empty[0 * 2 + st], ph ^ 1);
      if (j - 2 < nB) mbar_wait(&v_empty[1 * 2 + st], ph ^ 1);
    }
    mbar_arrive_expect_tx(&v_full[st], HEAD_BYTES);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      tma_load_3d(smem + V_OFF + st * HEAD_BYTES + b * A4_BOX, &tmQKV, &v_full[st],
                  args.v_off + kv_head * args.hs + b * 64, j * 128, batch);
  };

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    fence_barrier_init();
    pdl_wait();  // the only thread that reads global memory before the CTA-wide barrier below
    // start the first loads right away: they overlap the TMEM allocation and the CTA-wide barrier below
    for (int q = 0; q < (actB ? 2 : 1); ++q) {
      mbar_arrive_expect_tx(&q_full[q], HEAD_BYTES);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        tma_load_3d(smem + Q_OFF + q * HEAD_BYTES + b * A4_BOX, &tmQKV, &q_full[q],
                    args.q_off + head * args.hs + b * 64, q0A + q * 128, batch);
    }
    for (int j = 0; j < min(2, n_kv); ++j) load_kv(j);
  }
  if (warp == 3) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everyone else: before any global store (O rows) or further loads

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (tiles 0, 1 already in flight)
    if (lane == 0) {
      for (int j = 2; j < n_kv; ++j) load_kv(j);
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuer of tile q
    const int q = warp - 1;
    const int n_q = q ? nB : nA;
    if (lane == 0 && n_q > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
      const uint32_t q_addr = smem_u32(smem + Q_OFF + q * HEAD_BYTES);
      const uint32_t s_tm = tmem_base + TM_S + (uint32_t)q * 128;
      const uint32_t d_o = tmem_base + TM_O + (uint32_t)q * 128;
      auto issue_S = [&](int j, int h) {  // S_q(j).h = Q_q K_j[64h .. 64h+63]^T
        const uint32_t k_addr = smem_u32(smem + K_OFF + (j & 1) * HEAD_BYTES) + (uint32_t)h * 64 * 128;
#pragma unroll
        for (int ks = 0; ks < KS_QK; ++ks) {
          const uint32_t off = (ks >> 2) * A4_BOX + (ks & 3) * 32;
          umma_ss(s_tm + (uint32_t)h * 64, umma_desc_sw128(q_addr + off, 0, 1024),
                  umma_desc_sw128(k_addr + off, 0, 1024), idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[q * 2 + h]);
      };
      auto issue_PV = [&](int j, int h) {  // O_q += P_q(j).h V_j[64h .. 64h+63]
        const uint32_t v_addr = smem_u32(smem + V_OFF + (j & 1) * HEAD_BYTES);
        mbar_wait(&p_full[q * 2 + h], j & 1);
        tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int ks = h * 4 + k4;
          const uint64_t vdesc = umma_desc_sw128(v_addr + ks * 2048, A4_BOX, 1024);
          umma_ts(d_o, s_tm + (uint32_t)h * 64 + k4 * 8, vdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[q]);
      };
      mbar_wait(&q_full[q], 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_S(0, 0);
      const bool hi0 = 64 < T;
      if (hi0) issue_S(0, 1);
      umma_commit(&k_empty[q * 2 + 0]);
      for (int j = 0; j < n_q; ++j) {
        const bool hi = j * 128 + 64 < T;            // second half of this tile holds real keys
        const bool more = j + 1 < n_q;
        const bool hi_next = more && ((j + 1) * 128 + 64 < T);
        mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        issue_PV(j, 0);
        if (more) {
          mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_S(j + 1, 0);
        }
        if (hi) issue_PV(j, 1);
        umma_commit(&v_empty[q * 2 + (j & 1)]);
        if (more) {
          if (hi_next) issue_S(j + 1, 1);
          umma_commit(&k_empty[q * 2 + ((j + 1) & 1)]);
        }
        if (!more) umma_commit(&o_done[q]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax warpgroup: row r of tile A and of tile B
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    float mA = -INFINITY, lA = 0.f, mB = -INFINITY, lB = 0.f;
    int t = 0;  // half steps done so far (the same count for both tiles while both are active)
    const int n_dual = actB ? nA : 0;  // nA <= nB whenever tile B exists
    for (int j = 0; j < n_dual; ++j) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && j * 128 + 64 >= T) break;  // trailing half tile beyond the sequence: skipped by the issuers too
        a4_half_step<DH, CAUSAL, EMU, 2>(args, s_full, p_full, pv_done, lane_addr, 0, qp, j, h, T, r, t, mA, lA, mB, lB);
        ++t;
      }
    }
    if (!actB) {
      for (int j = 0; j < nA; ++j) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && j * 128 + 64 >= T) break;
          a4_half_step<DH, CAUSAL, EMU, 1>(args, s_full, p_full, pv_done, lane_addr, 0, qp, j, h, T, r, t, mA, lA, mA, lA);
          ++t;
        }
      }
    } else {
      for (int j = n_dual; j < nB; ++j) {  // causal: tile B's diagonal key tile
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && j * 128 + 64 >= T) break;
          a4_half_step<DH, CAUSAL, EMU, 1>(args, s_full, p_full, pv_done, lane_addr, 1, qp, j, h, T, r, t, mB, lB, mB, lB);
          ++t;
        }
      }
    }
    // ---- epilogue: O_q / l_q -> bf16 rows, tile A then tile B
#pragma unroll 1
    for (int q = 0; q < (actB ? 2 : 1); ++q) {
      mbar_wait(&o_done[q], 0);
      tc_fence_after();
      const float inv_l = 1.0f / (q ? lB : lA);
      const int q_idx = q0A + q * 128 + r;
      const bool row_ok = q_idx < T;
      const uint32_t o_addr = lane_addr + TM_O + (uint32_t)q * 128;
      bf16* op = args.O + ((long long)batch * T + q_idx) * args.ldo + (long long)head * args.dh_out;
#pragma unroll 1
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t o[32];
        LP_TMEM_LD_32(o, o_addr + c * 32);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c * 32 + g * 8 < args.dh_out) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
              u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
              u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
              u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(op + c * 32 + g * 8) = u;
            }
          }
        }
      }
      if (DH % 32 != 0) {  // DH = 80: the last 16 columns
        constexpr int c0 = (DH / 32) * 32;
        uint32_t o[16];
        LP_TMEM_LD_16(o, o_addr + c0);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (c0 + g * 8 < args.dh_out) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
              u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
              u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
              u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(op + c0 + g * 8) = u;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH, bool CAUSAL, int EMU>
static int launch4(const CUtensorMap& tm, const AttnArgs& a, cudaStream_t stream) {
  constexpr int NB = (DH + 63) / 64;
  constexpr int SMEM = 6 * NB * A4_BOX + 512 + 1024;
  auto kern = attn4_fwd_kernel<DH, CAUSAL, EMU>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid((a.T + 255) / 256, a.n_heads, a.B);
  LP_CHECK_CUDA(launch_ex(kern, grid, dim3(A4_THREADS), SMEM, stream, 1, prefill_pdl(), tm, a));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_attn4(const CUtensorMap& tm, const AttnArgs& a, int head_dim, int causal, int emu, cudaStream_t stream) {
  if (head_dim == 72) {
    if (emu == 2) return launch4<80, false, 2>(tm, a, stream);
    return emu == 1 ? launch4<80, false, 1>(tm, a, stream) : launch4<80, false, 0>(tm, a, stream);
  }
  if (causal) {
    if (emu == 2) return launch4<128, true, 2>(tm, a, stream);
    return emu == 1 ? launch4<128, true, 1>(tm, a, stream) : launch4<128, true, 0>(tm, a, stream);
  }
  if (emu == 2) return launch4<128, false, 2>(tm, a, stream);
  return emu == 1 ? launch4<128, false, 1>(tm, a, stream) : launch4<128, false, 0>(tm, a, stream);
}

}  // namespace lp
