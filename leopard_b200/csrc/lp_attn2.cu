// leopard_b200 — flash attention v2: TWO 128-row query tiles per CTA, ping-ponged on one tensor pipe.
//
// Same contract as attn_fwd_kernel (lp_attn.cu) — SigLIP non-causal 16x72 (DH=80 padded) and decoder causal GQA
// 32/8x128 — but the CTA owns 256 query rows of one head as tiles A and B with their own softmax warpgroups.  While
// warpgroup A runs softmax on S_A(j), the single MMA thread keeps the tensor pipe busy with S_B(j) / P_B·V, and vice
// versa, which roughly doubles tensor-pipe occupancy over the one-tile kernel (ncu: 31 % -> see profiles/).
//   warp 0      TMA producer : Q_A, Q_B once; K_j, V_j (128 rows) into 2-deep rings shared by both tiles
//   warp 1      MMA issuer   : S_q(j) = Q_q K_j^T  (SS, into TMEM S_q);  O_q += P_q(j) V_j  (A = P from TMEM, V MN-major)
//   warps 2-5   softmax A    : one thread per row, whole S row in registers (single TMEM pass), exp2 online softmax,
//   warps 6-9   softmax B      lazy O rescale (threshold 2^8), P -> bf16 -> TMEM (aliases S_q) published in two halves so
//                              P·V starts before the row is fully exponentiated, final O/l -> global
// TMEM map (512 columns): S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512).
// Ordering facts relied upon: tcgen05.mma ops of one thread execute in issue order, and tcgen05.commit covers ALL
// previously issued MMAs — so s_full_q(j) also implies P_q(j-1)·V has retired (O_q may be rescaled), and S_q(j+1),
// issued after P_q(j)·V, cannot overwrite P_q(j) early.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

constexpr int A2_BOX = 128 * 128;  // one TMA box: 128 rows x 64 bf16
constexpr int A2_THREADS = 320;

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], max rel. error 1.1e-4 — below the
// bf16 rounding of P): offloads part of the softmax exponentials from the 16-lane/clk MUFU unit, which is the
// co-bottleneck of attention on sm_100 (as many MUFU cycles as MMA cycles per tile at head_dim 128).
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.0f);           // f in [-0.5, 0.5]
  float p = fmaf(f, 0.05459282547f, 0.24221783876f);
  p = fmaf(p, f, 0.69336861372f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));  // * 2^round(x)
}

// packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): halves the issue slots of the per-element softmax math, which
// is issue-bound before it is MUFU-bound (measured: moving exponentials to the FMA pipe made the kernel slower)
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t r, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// named barriers for the softmax turn protocol (ids 2 and 3; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

#define LP_TMEM_ST_16(regs, taddr)                                                                                  \
  asm volatile(                                                                                                     \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, " \
      "%15, %16};" ::"r"(taddr),                                                                                    \
      "r"(regs[0]), "r"(regs[1]), "r"(regs[2]), "r"(regs[3]), "r"(regs[4]), "r"(regs[5]), "r"(regs[6]),             \
      "r"(regs[7]), "r"(regs[8]), "r"(regs[9]), "r"(regs[10]), "r"(regs[11]), "r"(regs[12]), "r"(regs[13]),         \
      "r"(regs[14]), "r"(regs[15])                                                                                  \
      : "memory")

// EMU: how many of every 4 exponentials are evaluated with ex2_fma instead of MUFU.EX2 (0, 1 or 2)
template <int DH, bool CAUSAL, int EMU>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  constexpr int NB = (DH + 63) / 64;
  constexpr int KS_QK = DH / 16;
  constexpr int KS_PV = 8;
  constexpr int HEAD_BYTES = NB * A2_BOX;
  constexpr int Q_OFF = 0;                       // 2 tiles
  constexpr int K_OFF = Q_OFF + 2 * HEAD_BYTES;  // 2 stages
  constexpr int V_OFF = K_OFF + 2 * HEAD_BYTES;  // 2 stages
  constexpr int BAR_OFF = V_OFF + 2 * HEAD_BYTES;
  constexpr uint32_t TM_S = 0, TM_O = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);  // [2]
  uint64_t* k_full = q_full + 2;                                    // [2]
  uint64_t* k_empty = k_full + 2;                                   // [2]
  uint64_t* v_full = k_empty + 2;                                   // [2]
  uint64_t* v_empty = v_full + 2;                                   // [2]
  uint64_t* s_full = v_empty + 2;                                   // [2] per q tile
  uint64_t* p_full = s_full + 2;                                    // [4]: q tile x {first, second} half of P
  uint64_t* o_done = p_full + 4;                                    // [2] per q tile
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

  // K/V tile loader (producer thread only): tile j -> ring slot j & 1
  auto load_kv = [&](int j) {
    const int st = j & 1;
    const uint32_t ph = (j >> 1) & 1;
    mbar_wait(&k_empty[st], ph ^ 1);
    mbar_arrive_expect_tx(&k_full[st], HEAD_BYTES);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      tma_load_3d(smem + K_OFF + st * HEAD_BYTES + b * A2_BOX, &tmQKV, &k_full[st],
                  args.k_off + kv_head * args.hs + b * 64, j * 128, batch);
    mbar_wait(&v_empty[st], ph ^ 1);
    mbar_arrive_expect_tx(&v_full[st], HEAD_BYTES);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      tma_load_3d(smem + V_OFF + st * HEAD_BYTES + b * A2_BOX, &tmQKV, &v_full[st],
                  args.v_off + kv_head * args.hs + b * 64, j * 128, batch);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 128);
      mbar_init(&p_full[2 * i + 1], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
    // start the first loads right away: they overlap the TMEM allocation and the CTA-wide barrier below
    for (int q = 0; q < (actB ? 2 : 1); ++q) {
      mbar_arrive_expect_tx(&q_full[q], HEAD_BYTES);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        tma_load_3d(smem + Q_OFF + q * HEAD_BYTES + b * A2_BOX, &tmQKV, &q_full[q],
                    args.q_off + head * args.hs + b * 64, q0A + q * 128, batch);
    }
    for (int j = 0; j < min(2, n_kv); ++j) load_kv(j);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (tiles 0, 1 already in flight)
    if (lane == 0) {
      for (int j = 2; j < n_kv; ++j) load_kv(j);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
      auto issue_S = [&](int q, int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + Q_OFF + q * HEAD_BYTES);
        const uint32_t k_addr = smem_u32(smem + K_OFF + st * HEAD_BYTES);
        const uint32_t d_s = tmem_base + TM_S + (uint32_t)q * 128;
#pragma unroll
        for (int ks = 0; ks < KS_QK; ++ks) {
          const uint32_t off = (ks >> 2) * A2_BOX + (ks & 3) * 32;
          umma_ss(d_s, umma_desc_sw128(q_addr + off, 0, 1024), umma_desc_sw128(k_addr + off, 0, 1024), idesc_s,
                  ks > 0 ? 1u : 0u);
        }
        const bool last_reader = (q == 1) || !(j < nB);
        if (last_reader) umma_commit(&k_empty[st]);
        umma_commit(&s_full[q]);
      };
      auto issue_PV = [&](int q, int j, int n_q) {
        const int st = j & 1;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        const uint32_t v_addr = smem_u32(smem + V_OFF + st * HEAD_BYTES);
        const uint32_t d_o = tmem_base + TM_O + (uint32_t)q * 128;
        const uint32_t a_t = tmem_base + TM_S + (uint32_t)q * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          // P arrives in two halves (kv columns 0-63, 64-127): the first four K steps start while the softmax
          // warpgroup is still exponentiating the second half
          mbar_wait(&p_full[2 * q + half], j & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = half * (KS_PV / 2); ks < (half + 1) * (KS_PV / 2); ++ks) {
            const uint64_t vdesc = umma_desc_sw128(v_addr + ks * 2048, A2_BOX, 1024);
            umma_ts(d_o, a_t + ks * 8, vdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          }
        }
        const bool last_reader = (q == 1) || !(j < nB);
        if (last_reader) umma_commit(&v_empty[st]);
        if (j == n_q - 1) umma_commit(&o_done[q]);
      };
      mbar_wait(&q_full[0], 0);
      issue_S(0, 0);
      if (actB) {
        mbar_wait(&q_full[1], 0);
        issue_S(1, 0);
      }
      for (int j = 0; j < n_kv; ++j) {
        if (j < nA) {
          issue_PV(0, j, nA);
          if (j + 1 < nA) issue_S(0, j + 1);
        }
        if (j < nB) {
          issue_PV(1, j, nB);
          if (j + 1 < nB) issue_S(1, j + 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int q = (warp - 2) >> 2;
    const int n_q = q ? nB : nA;
    // Turn protocol (args.pingpong): the two softmax warpgroups alternate strictly — A(j), B(j), A(j+1), ... — so that
    // one tile's exponentials (MUFU-bound, ~1000 cycles per 128 x 128 tile) run while the tensor pipe works on the
    // OTHER tile's P·V and next QK^T.  Without it both warpgroups start in phase, share the MUFU unit, finish together
    // and then both wait for their MMAs: softmax and MMA time add up instead of overlapping (ncu: tensor pipe 47 %).
    // Barrier 3 = "B's turn" (A arrives, B syncs), barrier 2 = "A's turn" (B arrives, A syncs from its second step);
    // both warpgroups run n_kv = nB >= nA turns, a finished tile just passes the turn on.
    const bool turns = args.pingpong != 0 && actB;
    const int n_turn = turns ? n_kv : n_q;
    if (n_turn > 0) {
      const int lg = warp & 3;
      const int r = lg * 32 + lane;
      const int q0 = q ? q0B : q0A;
      const int q_idx = q0 + r;
      const int diag = q0 >> 7;  // kv tile that holds this q tile's diagonal
      const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
      const uint32_t s_addr = lane_addr + TM_S + (uint32_t)q * 128;
      const uint32_t o_addr = lane_addr + TM_O + (uint32_t)q * 128;
      float m_used = -INFINITY, l_sum = 0.f;
      for (int j = 0; j < n_turn; ++j) {
        if (turns) {
          if (q == 0) { if (j > 0) named_bar_sync(2, 256); }
          else named_bar_sync(3, 256);
        }
        if (j < n_q) {
        const int kv0 = j * 128;
        mbar_wait(&s_full[q], j & 1);
        tc_fence_after();
        uint32_t v[128];
        LP_TMEM_LD_32((v + 0), s_addr + 0);
        LP_TMEM_LD_32((v + 32), s_addr + 32);
        LP_TMEM_LD_32((v + 64), s_addr + 64);
        LP_TMEM_LD_32((v + 96), s_addr + 96);
        tmem_wait_ld();
        const bool need_mask = (kv0 + 128 > T) || (CAUSAL && j == diag);
        int lim = T - kv0;
        if (CAUSAL && j == diag) lim = min(lim, r + 1);
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= lim) v[i] = 0xff800000u;  // -inf
        }
        float mx;
        {  // four independent max chains (a single chain is 64 dependent FMNMX3)
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        }
        const float m_new = fmaxf(m_used, mx * args.scale_log2);
        const bool grow = (m_new - m_used) > 8.0f;
        if (j == 0) {
          m_used = m_new;
        } else if (__any_sync(0xffffffffu, grow)) {
          // P_q(j-1)·V has retired (s_full_q(j) was committed after it): O_q may be rescaled in place
          float alpha = 1.0f;
          if (grow) {
            alpha = ex2a(m_used - m_new);
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
            LP_TMEM_ST_16(o, o_addr + c * 16);
          }
          tmem_wait_st();
        }
        const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
        const uint64_t sc2 = pk2(args.scale_log2, args.scale_log2), nm2 = pk2(neg_m, neg_m);
        uint64_t acc2 = pk2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1;
            upk2(fma2(pk2(__uint_as_float(v[c * 32 + i]), __uint_as_float(v[c * 32 + i + 1])), sc2, nm2), x0, x1);
            // compile-time pattern: element (i & 3) < EMU of each group of four goes to the FMA pipe
            const float p0 = ((i & 3) < EMU) ? ex2_fma(x0) : ex2a(x0);
            const float p1 = (((i + 1) & 3) < EMU) ? ex2_fma(x1) : ex2a(x1);
            const uint32_t u = pack_bf16x2(p0, p1);
            const float2 f = unpack_bf16x2(u);   // row sum from the bf16-rounded values the MMA will actually use
            acc2 = add2(acc2, pk2(f.x, f.y));
            pk[i >> 1] = u;
          }
          LP_TMEM_ST_16(pk, s_addr + c * 16);  // P_q(j) overwrites the first 64 columns of S_q (already in registers)
          if (c & 1) {  // half of P complete (chunks 0-1 / 2-3): publish it
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&p_full[2 * q + (c >> 1)]);
          }
        }
        {
          float a0, a1;
          upk2(acc2, a0, a1);
          l_sum += a0 + a1;
        }
        }  // j < n_q
        if (turns) {
          if (q == 0) named_bar_arrive(3, 256);
          else if (j + 1 < n_turn) named_bar_arrive(2, 256);
        }
      }
      // ---- epilogue
      mbar_wait(&o_done[q], 0);
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
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH, bool CAUSAL, int EMU>
static int launch2(const CUtensorMap& tm, const AttnArgs& a, cudaStream_t stream) {
  constexpr int NB = (DH + 63) / 64;
  constexpr int SMEM = 6 * NB * A2_BOX + 256 + 1024;
  auto kern = attn2_fwd_kernel<DH, CAUSAL, EMU>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid((a.T + 255) / 256, a.n_heads, a.B);
  kern<<<grid, A2_THREADS, SMEM, stream>>>(tm, a);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_attn2(const CUtensorMap& tm, const AttnArgs& a, int head_dim, int causal, int emu, cudaStream_t stream) {
  if (head_dim == 72) {
    if (emu == 2) return launch2<80, false, 2>(tm, a, stream);
    return emu == 1 ? launch2<80, false, 1>(tm, a, stream) : launch2<80, false, 0>(tm, a, stream);
  }
  if (causal) {
    if (emu == 2) return launch2<128, true, 2>(tm, a, stream);
    return emu == 1 ? launch2<128, true, 1>(tm, a, stream) : launch2<128, true, 0>(tm, a, stream);
  }
  if (emu == 2) return launch2<128, false, 2>(tm, a, stream);
  return emu == 1 ? launch2<128, false, 1>(tm, a, stream) : launch2<128, false, 0>(tm, a, stream);
}

}  // namespace lp
