// leopard_b200 — flash attention v3: two 128-row query tiles per CTA, S delivered in 64-key halves, one MMA issuer
// per tile.  Same contract as attn2_fwd_kernel (lp_attn2.cu): SigLIP non-causal 16x72 (DH = 80 padded) and decoder
// causal GQA 32/8x128, replacing flash_attn_varlen_func (MP/model/llava/transformer.py:476-515,869-877) and the ViT's
// TEDotProductAttention (MP/model/llava/clip_encoder.py:58-63).
//
// Why: in the v2 kernel P_q(j) aliases S_q(j) in TMEM, so S_q(j+1) = Q K_{j+1}^T can only be issued after P_q(j)·V —
// every tile runs the serial chain softmax -> P·V -> QK^T -> softmax, and ncu showed the softmax warps spending 36 %
// of all samples waiting for S while the tensor pipe was 47 % and the MUFU pipe 47 % busy.  Here each S_q is split
// into two 64-key halves with their own TMEM columns; P of a half overwrites only that half.  While the softmax
// warpgroup works on the hi half of tile j, the issuer has already consumed P.lo(j) and written S.lo(j+1) into the lo
// columns, so in steady state the softmax never waits for the tensor pipe.  Each query tile has its OWN MMA-issuing
// thread (a blocking in-order loop), so the two tiles are decoupled from each other as well.
//   warp 0      TMA producer : Q_A, Q_B once; K_j, V_j (128 rows) into 2-deep rings shared by both tiles
//   warp 1 / 2  MMA issuer of tile A / B:  S_q(j).h = Q_q K_j[h]^T (N = 64), O_q += P_q(j).h V_j[h] (A from TMEM)
//   warp 3      TMEM allocation
//   warps 4-7   softmax A    : one thread per row, 64 S values per half-step in registers, exp2 online softmax,
//   warps 8-11  softmax B      lazy O rescale (threshold 2^8), P -> bf16 -> TMEM over its own S half
// TMEM map (512 columns): S_A.lo [0,64) S_A.hi [64,128) S_B.lo [128,192) S_B.hi [192,256) O_A [256,384) O_B [384,512).
// Ordering facts relied upon: tcgen05.mma ops issued by ONE thread execute in issue order (P·V(j).h is issued before
// S(j+1).h, so the new S cannot overwrite P early), and tcgen05.commit covers all MMAs previously issued by that
// thread.  O_q is rescaled in place only after pv_done_q shows that every P·V issued so far has retired (a pair of alternating
// barriers per tile, looked at only on that path).  The S half of step t + 1 is loaded into the registers of step t's S before
// P(t) is handed over (its barrier is probed before and after the exponentials).  Grid = (heads, query-tile pairs, batch), pairs
// descending: heaviest causal pairs first over the WHOLE grid (see the kernel body; found with the phase trace, A3_TRACE below).
// A trailing half tile that lies entirely beyond T (T = 676: keys 704..767) is skipped by issuer and softmax alike.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

constexpr int A3_BOX = 128 * 128;  // one TMA box: 128 rows x 64 bf16
constexpr int A3_THREADS = 384;

__device__ __forceinline__ float a3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], max rel. error 1.1e-4 — below the
// bf16 rounding of P): takes a share of the exponentials off the 16-lane/clk MUFU unit, which bounds the exp
// sections of this kernel (ncu: XU pipe ~75 % busy inside them).
__device__ __forceinline__ float a3_ex2_fma(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.0f);           // f in [-0.5, 0.5]
  float p = fmaf(f, 0.05459282547f, 0.24221783876f);
  p = fmaf(p, f, 0.69336861372f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));  // * 2^round(x)
}
__device__ __forceinline__ void a3_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void a3_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint64_t a3_pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void a3_upk2(uint64_t r, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ uint64_t a3_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t a3_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

#define LP3_TMEM_ST_16(regs, taddr)                                                                                 \
  asm volatile(                                                                                                     \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, " \
      "%15, %16};" ::"r"(taddr),                                                                                    \
      "r"(regs[0]), "r"(regs[1]), "r"(regs[2]), "r"(regs[3]), "r"(regs[4]), "r"(regs[5]), "r"(regs[6]),             \
      "r"(regs[7]), "r"(regs[8]), "r"(regs[9]), "r"(regs[10]), "r"(regs[11]), "r"(regs[12]), "r"(regs[13]),         \
      "r"(regs[14]), "r"(regs[15])                                                                                  \
      : "memory")

// Phase trace (tools/micro/attn_trace.cu compiles this file with -DLP_ATTN_TRACE): lane 0 of the two issuer warps and of the first
// warp of each softmax warpgroup of ONE chosen CTA stamp clock64() at the protocol points below; no code in normal builds.
#ifdef LP_ATTN_TRACE
__device__ unsigned long long* g_a3_trace = nullptr;   // [4 slots][4096]: slot 0 / 1 = issuer A / B, 2 / 3 = softmax A / B
__device__ int g_a3_trace_blk = 0;
#define A3_TRACE_INIT()                                                                                                   \
  unsigned long long* tr_buf = nullptr;                                                                                   \
  int tr_n = 0;                                                                                                           \
  if ((int)blockIdx.y == g_a3_trace_blk && blockIdx.x == 0 && blockIdx.z == 0 && lane == 0 && g_a3_trace != nullptr &&    \
      (warp == 1 || warp == 2 || warp == 4 || warp == 8))                                                                 \
    tr_buf = g_a3_trace + (warp == 1 ? 0 : warp == 2 ? 1 : warp == 4 ? 2 : 3) * 4096
#define A3_TRACE(id)                                                                                                      \
  do {                                                                                                                    \
    if (tr_buf != nullptr && tr_n < 4094) {                                                                               \
      tr_buf[1 + tr_n] = ((unsigned long long)(id) << 48) | ((unsigned long long)clock64() & 0xffffffffffffULL);          \
      ++tr_n;                                                                                                             \
      tr_buf[0] = (unsigned long long)tr_n;                                                                               \
    }                                                                                                                     \
  } while (0)
#else
#define A3_TRACE_INIT() do {} while (0)
#define A3_TRACE(id) do {} while (0)
#endif

// EMU: how many of every 4 exponentials are evaluated with a3_ex2_fma instead of MUFU.EX2 (0, 1 or 2)
template <int DH, bool CAUSAL, int EMU>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn3_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  constexpr int NB = (DH + 63) / 64;
  constexpr int KS_QK = DH / 16;
  constexpr int HEAD_BYTES = NB * A3_BOX;
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
  uint64_t* pv_done = p_full + 4;                                   // [4]   q tile x (half-step index & 1): one phase per P·V half
  uint64_t* o_done = pv_done + 4;                                   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  A3_TRACE_INIT();

  // Grid = (heads, query-tile pairs, batch) with the heads fastest and the pairs in descending order: the block scheduler hands out
  // CTAs in linear order, so every head's heaviest causal pair (2 qp + 2 key tiles) starts first and the lightest pairs fill the
  // tail — longest-processing-time-first over the WHOLE grid.  (With the pairs fastest only each head's own pairs were ordered; the
  // last heads' 38-tile CTAs started at ~90 % of the kernel and ran alone: 440 k cycles instead of 353 k at S = 4796 by list-scheduling
  // simulation, 0.208 -> 0.174 ms measured.)
  const int qp = (int)gridDim.y - 1 - (int)blockIdx.y;
  const int head = blockIdx.x, batch = blockIdx.z;
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
      tma_load_3d(smem + K_OFF + st * HEAD_BYTES + b * A3_BOX, &tmQKV, &k_full[st],
                  args.k_off + kv_head * args.hs + b * 64, j * 128, batch);
    if (j >= 2) {
      if (j - 2 < nA) mbar_wait(&v_empty[0 * 2 + st], ph ^ 1);
      if (j - 2 < nB) mbar_wait(&v_empty[1 * 2 + st], ph ^ 1);
    }
    mbar_arrive_expect_tx(&v_full[st], HEAD_BYTES);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      tma_load_3d(smem + V_OFF + st * HEAD_BYTES + b * A3_BOX, &tmQKV, &v_full[st],
                  args.v_off + kv_head * args.hs + b * 64, j * 128, batch);
  };

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&pv_done[i], 1);
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
        tma_load_3d(smem + Q_OFF + q * HEAD_BYTES + b * A3_BOX, &tmQKV, &q_full[q],
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
      int t_pv = 0;  // P·V halves issued so far
      auto issue_S = [&](int j, int h) {  // S_q(j).h = Q_q K_j[64h .. 64h+63]^T
        const uint32_t k_addr = smem_u32(smem + K_OFF + (j & 1) * HEAD_BYTES) + (uint32_t)h * 64 * 128;
#pragma unroll
        for (int ks = 0; ks < KS_QK; ++ks) {
          const uint32_t off = (ks >> 2) * A3_BOX + (ks & 3) * 32;
          umma_ss(s_tm + (uint32_t)h * 64, umma_desc_sw128(q_addr + off, 0, 1024),
                  umma_desc_sw128(k_addr + off, 0, 1024), idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[q * 2 + h]);
        A3_TRACE(14 + h);
      };
      auto issue_PV = [&](int j, int h) {  // O_q += P_q(j).h V_j[64h .. 64h+63]
        const uint32_t v_addr = smem_u32(smem + V_OFF + (j & 1) * HEAD_BYTES);
        A3_TRACE(10 + h);
        mbar_wait(&p_full[q * 2 + h], j & 1);
        A3_TRACE(12 + h);
        tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int ks = h * 4 + k4;
          const uint64_t vdesc = umma_desc_sw128(v_addr + ks * 2048, A3_BOX, 1024);
          umma_ts(d_o, s_tm + (uint32_t)h * 64 + k4 * 8, vdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[q * 2 + (t_pv & 1)]);
        ++t_pv;
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
    // ------------------------------------------------------------------ softmax warpgroups
    const int q = (warp - 4) >> 2;
    const int n_q = q ? nB : nA;
    if (n_q > 0) {
      const int lg = warp & 3;
      const int r = lg * 32 + lane;
      const int q0 = q ? q0B : q0A;
      const int q_idx = q0 + r;
      const int diag = q0 >> 7;  // kv tile that holds this q tile's diagonal
      const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
      const uint32_t s_addr = lane_addr + TM_S + (uint32_t)q * 128;
      const uint32_t o_addr = lane_addr + TM_O + (uint32_t)q * 128;
      float m_used = -INFINITY, l_sum = 0.f;
      int t = 0;  // P·V halves published so far == index of the current half step
      // Optional turn protocol around the EXP sections only (args.pingpong): tile A's and tile B's exponentials
      // alternate A(0) B(0) A(1) B(1) ... so each runs at the full MUFU rate while the other tile's TMEM loads, row
      // max and P stores overlap it.  Named barrier 3 = "B may exp" (A arrives, B syncs), 2 = "A may exp".  Only the
      // first hA half steps take part (tile A never has more half steps than tile B).
      auto halves = [&](int n) { return 2 * n - ((n > 0 && (n - 1) * 128 + 64 >= T) ? 1 : 0); };
      const int hA = halves(nA);
      const bool turns = args.pingpong != 0 && actB;
      // The S half of step t + 1 goes into the registers of step t's S as soon as its exponentials are done — before P(t) is
      // handed to the issuer — so the barrier / tcgen05.ld round trips overlap the tcgen05.st drain and the hand-over instead of
      // heading the next step (phase trace, tools/micro/attn_trace.cu: ~300 of ~2080 cycles per half step).  S(t + 1) never
      // depends on P(t) (it follows P·V(t - 1) in the issuer's order), so waiting for it first cannot deadlock; it is only tried,
      // not waited for, before the hand-over.
      uint32_t v[64];
      A3_TRACE(1);
      mbar_wait(&s_full[q * 2 + 0], 0);
      A3_TRACE(2);
      tc_fence_after();
      LP_TMEM_LD_32((v + 0), s_addr + 0);
      LP_TMEM_LD_32((v + 32), s_addr + 32);
      int j = 0, h = 0;
#pragma unroll 1
      for (;;) {
        {
          const int kv0 = j * 128 + h * 64;
          tmem_wait_ld();
          A3_TRACE(3);
          const bool need_mask = (kv0 + 64 > T) || (CAUSAL && j == diag);
          int lim = T - kv0;
          if (CAUSAL && j == diag) lim = min(lim, r + 1 - h * 64);
          if (need_mask) {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i >= lim) v[i] = 0xff800000u;  // -inf
          }
          A3_TRACE(31);
          float mx;
          {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; i += 8) {
              m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
              m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
              m2 = fmaxf(m2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
              m3 = fmaxf(m3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
            }
            mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          }
          const float m_new = fmaxf(m_used, mx * args.scale_log2);
          A3_TRACE(32);
          const bool grow = (m_new - m_used) > 8.0f;
          // pv_done[q][b] completes one phase per P·V half with (index & 1) == b; it is looked at only when O has to be rescaled.
          // Half t - 1 is phase (t - 1) >> 1 of barrier (t - 1) & 1: the phase before it has completed by now (S(t - 1) followed
          // P·V(t - 3) on the tensor pipe) and the one after it cannot start before P(t + 1), so the parity is unambiguous.
          if (t == 0) {
            m_used = m_new;
          } else if (__any_sync(0xffffffffu, grow)) {
            // every P·V issued so far (halves 0 .. t-1) must have retired before O_q is rescaled in place
            mbar_wait(&pv_done[q * 2 + ((t - 1) & 1)], ((t - 1) >> 1) & 1);
            tc_fence_after();
            float alpha = 1.0f;
            if (grow) {
              alpha = a3_ex2(m_used - m_new);
              m_used = m_new;
              l_sum *= alpha;
            }
#pragma unroll 1
            for (int c = 0; c < DH / 16; ++c) {
              uint32_t o[16];
              LP_TMEM_LD_16(o, o_addr + c * 16);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              LP3_TMEM_ST_16(o, o_addr + c * 16);
            }
            tmem_wait_st();
          }
          const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
          const uint64_t sc2 = a3_pk2(args.scale_log2, args.scale_log2), nm2 = a3_pk2(neg_m, neg_m);
          uint64_t acc2 = a3_pk2(0.f, 0.f);
          A3_TRACE(4);
          if (turns && t < hA) {
            if (q == 0) { if (t > 0) a3_bar_sync(2, 256); }
            else a3_bar_sync(3, 256);
          }
          A3_TRACE(5);
          // next half step (the trailing half tile beyond the sequence is skipped by the issuer too); its S barrier is probed here so
          // that the probe's shared-memory round trip hides under the exponentials
          int jn = j, hn = 1;
          if (h == 1 || j * 128 + 64 >= T) { jn = j + 1; hn = 0; }
          const bool has_next = jn < n_q;
          const bool s_ready = has_next && mbar_test_wait(&s_full[q * 2 + hn], jn & 1);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float x0, x1;
              a3_upk2(a3_fma2(a3_pk2(__uint_as_float(v[c * 32 + i]), __uint_as_float(v[c * 32 + i + 1])), sc2, nm2), x0,
                      x1);
              // compile-time pattern: element (i & 3) < EMU of each group of four goes to the FMA pipe
              const float p0 = ((i & 3) < EMU) ? a3_ex2_fma(x0) : a3_ex2(x0);
              const float p1 = (((i + 1) & 3) < EMU) ? a3_ex2_fma(x1) : a3_ex2(x1);
              // row sum in fp32 before the bf16 rounding of P (as flash-attention does): the rounding errors of P are
              // unbiased and average out in O = sum(P V) / l; this saves the two unpack instructions per pair
              acc2 = a3_add2(acc2, a3_pk2(p0, p1));
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
            LP3_TMEM_ST_16(pk, s_addr + h * 64 + c * 16);  // P(j).h over the first 32 columns of its own S half
          }
          A3_TRACE(6);
          if (turns && t < hA) {
            if (q == 0) a3_bar_arrive(3, 256);
            else if (t + 1 < hA) a3_bar_arrive(2, 256);
          }
          bool loaded = false;
          if (s_ready || (has_next && mbar_test_wait(&s_full[q * 2 + hn], jn & 1))) {   // second probe: S often lands during the exponentials
            tc_fence_after();
            LP_TMEM_LD_32((v + 0), s_addr + hn * 64 + 0);
            LP_TMEM_LD_32((v + 32), s_addr + hn * 64 + 32);
            loaded = true;
          }
          tmem_wait_st();
          tc_fence_before();
          A3_TRACE(7);
          mbar_arrive(&p_full[q * 2 + h]);
          A3_TRACE(8);
          {
            float a0, a1;
            a3_upk2(acc2, a0, a1);
            l_sum += a0 + a1;
          }
          ++t;
          if (!has_next) break;
          if (!loaded) {
            A3_TRACE(1);
            mbar_wait(&s_full[q * 2 + hn], jn & 1);
            A3_TRACE(2);
            tc_fence_after();
            LP_TMEM_LD_32((v + 0), s_addr + hn * 64 + 0);
            LP_TMEM_LD_32((v + 32), s_addr + hn * 64 + 32);
          }
          j = jn;
          h = hn;
        }
      }
      // ---- epilogue
      mbar_wait(&o_done[q], 0);
      A3_TRACE(9);
      tc_fence_after();
      const float inv_l = 1.0f / l_sum;
      const bool row_ok = q_idx < T;
      bf16* op = args.O + ((long long)batch * T + q_idx) * args.ldo + (long long)head * args.dh_out;
#pragma unroll 1
      for (int c = 0; c < DH / 16; ++c) {
        uint32_t o[16];
        LP_TMEM_LD_16(o, o_addr + c * 16);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (c * 16 + g * 8 < args.dh_out) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
              u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
              u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
              u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(op + c * 16 + g * 8) = u;
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
static int launch3(const CUtensorMap& tm, const AttnArgs& a, cudaStream_t stream) {
  constexpr int NB = (DH + 63) / 64;
  constexpr int SMEM = 6 * NB * A3_BOX + 512 + 1024;
  auto kern = attn3_fwd_kernel<DH, CAUSAL, EMU>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid(a.n_heads, (a.T + 255) / 256, a.B);
  LP_CHECK_CUDA(launch_ex(kern, grid, dim3(A3_THREADS), SMEM, stream, 1, prefill_pdl(), tm, a));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_attn3(const CUtensorMap& tm, const AttnArgs& a, int head_dim, int causal, int emu, cudaStream_t stream) {
  if (head_dim == 72) {
    if (emu == 2) return launch3<80, false, 2>(tm, a, stream);
    return emu == 1 ? launch3<80, false, 1>(tm, a, stream) : launch3<80, false, 0>(tm, a, stream);
  }
  if (causal) {
    if (emu == 2) return launch3<128, true, 2>(tm, a, stream);
    return emu == 1 ? launch3<128, true, 1>(tm, a, stream) : launch3<128, true, 0>(tm, a, stream);
  }
  if (emu == 2) return launch3<128, false, 2>(tm, a, stream);
  return emu == 1 ? launch3<128, false, 1>(tm, a, stream) : launch3<128, false, 0>(tm, a, stream);
}

}  // namespace lp
