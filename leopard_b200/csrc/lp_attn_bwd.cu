// leopard_b200 — tcgen05 flash-attention BACKWARD, sm_100a: the decoder's causal GQA attention (head_dim 128) and the vision
// tower's full attention (head_dim 72 in the zero-padded 80-wide head layout of the fused qkv buffer).
//
// STATUS: written and compiled after the round's GPU budget was spent — NOT yet run on hardware.  train.py keeps using
// the materialised per-head backward (lp_softmax_bwd) until this one has passed tools/gpu_check_attn_bwd.py.
//
// Reference: autograd through flash_attn_varlen_func (Pai-Megatron-Patch/megatron_patch/model/llava/transformer.py:476-515,
// 869-877).  Specification: oracle/train.py attention_bwd.  With P = exp(scale q k^T - lse), delta = rowsum(dO * O):
//     dV = P^T dO        dS = P * (dO V^T - delta)        dQ = scale dS K        dK = scale dS^T Q
// Two kernels, both deterministic (no atomics), operands straight from the fused qkv buffer / the dO buffer by TMA:
//   attn_bwd_dq_kernel   one CTA per (128 query rows, q head).  Pass 1 recomputes the row log-sum-exp from S = Q K^T
//                        (no forward change needed) and delta; pass 2 per kv tile: S and dP = dO V^T by two SS MMAs into
//                        TMEM, 128 threads (one per row) form dS in registers and store it as bf16 over S's columns, then
//                        dQ += dS K with dS as the TMEM A operand and the K tile as an MN-major B operand.
//                        Writes dQ, and lse / delta for the second kernel.
//   attn_bwd_dkv_kernel  one CTA per (128 kv rows, kv head); loops over the group's q heads and the q tiles at or below
//                        the diagonal.  S^T = K Q^T and dP^T = V dO^T (so that P^T and dS^T have kv rows on the TMEM
//                        lanes), then dV += P^T dO and dK += dS^T Q with TMEM A operands and the dO / Q tiles as MN-major
//                        B operands; the sums over the group's heads never leave TMEM (4 x 128 columns = all 512).
// Warp roles as in lp_attn.cu: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warps 2-5 one thread per row.
#ifdef LP_EMULATE                  // CPU emulation build (tests/test_attention_kernels_emulated.py): same kernel source, the
#include "tc_emu.h"                // tcgen05 / TMA / mbarrier wrappers replaced by functional stand-ins (tests/emu/)
#include "emu_helpers.inc"         // descriptor builders + numerics helpers, cut from lp_common.cuh by the test
#else
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"
#endif

namespace lp {

constexpr int AB_T = 128;                 // tile edge (query rows, kv rows)
constexpr int AB_THREADS = 192;
constexpr int AB_BOX = 128 * 128;         // one TMA box: 128 rows x 64 bf16
constexpr int AB_HEAD = 2 * AB_BOX;       // one 128 x 128 bf16 tile = two boxes
constexpr int AB_KS = 8;                  // K = 16 steps per 128-deep contraction

struct AttnBwdArgs {
  int T, B, n_heads, group, Tpad;
  int q_off, k_off, v_off;                // column offsets of the q / k / v blocks inside a qkv (and dqkv) row
  int hs;                                 // columns per head in qkv / dqkv (128, or 80 = 72 + zero padding)
  int dho;                                // columns per head in O / dO (128 or 72)
  const bf16* O;                          // forward output [B*T, n_heads*dho]
  const bf16* dO;                         // its gradient, same layout
  long long ldo;
  bf16* dqkv;                             // [B*T, ldd], same column layout as qkv
  long long ldd;
  float* lse2;                            // [B, n_heads, Tpad]  log2-domain log-sum-exp of the scaled scores
  float* delta;                           // [B, n_heads, Tpad]
  float scale, scale_log2;
};

#ifdef LP_EMULATE
__device__ __forceinline__ float bw_ex2(float x) { return exp2f(x); }
__device__ __forceinline__ void row_threads_barrier() { named_bar(1, 128, true); }
#else
__device__ __forceinline__ float bw_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// named barrier 1: the four row warps (128 threads) only
__device__ __forceinline__ void row_threads_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

#define LP_TMEM_ST_16(regs, taddr)                                                                                  \
  asm volatile(                                                                                                     \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, " \
      "%15, %16};" ::"r"(taddr),                                                                                    \
      "r"(regs[0]), "r"(regs[1]), "r"(regs[2]), "r"(regs[3]), "r"(regs[4]), "r"(regs[5]), "r"(regs[6]),             \
      "r"(regs[7]), "r"(regs[8]), "r"(regs[9]), "r"(regs[10]), "r"(regs[11]), "r"(regs[12]), "r"(regs[13]),         \
      "r"(regs[14]), "r"(regs[15])                                                                                  \
      : "memory")
#endif

// D[128 x 128] (+)= A[128 x DH] B[128 x DH]^T, both K-major 128B-swizzled tiles of two boxes (the forward's S = Q K^T form)
template <int DH>
__device__ __forceinline__ void mma_ss_kmajor(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr) {
  constexpr uint32_t idesc = umma_idesc_bf16(AB_T, AB_T, 0, 0);
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) {   // contraction over the head dimension: DH / 16 steps (8, or 5 for the 80-wide layout)
    const uint32_t off = (ks >> 2) * AB_BOX + (ks & 3) * 32;
    umma_ss(d_tmem, umma_desc_sw128(a_addr + off, 0, 1024), umma_desc_sw128(b_addr + off, 0, 1024), idesc, ks > 0 ? 1u : 0u);
  }
}
// D[128 x DH] (+)= A[128 x 128 bf16 in TMEM, 64 columns] B, B = a [128 rows = contraction, DH contiguous] tile read as an
// MN-major operand (the forward's O += P V form: 16 contraction rows = 2 KiB per K step, LBO = next 64-wide slab)
template <int DH>
__device__ __forceinline__ void mma_ts_mnmajor(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, bool accumulate) {
  constexpr uint32_t idesc = umma_idesc_bf16(AB_T, DH, 0, 1);
#pragma unroll
  for (int ks = 0; ks < AB_KS; ++ks)
    umma_ts(d_tmem, a_tmem + ks * 8, umma_desc_sw128(b_addr + ks * 2048, AB_BOX, 1024), idesc,
            (accumulate || ks > 0) ? 1u : 0u);
}

// one thread's DH fp32 accumulators (TMEM lane) -> bf16 row in global memory; columns >= valid are written as zeros
// (dV of the 80-wide layout: its last 8 columns would otherwise carry the next head's dO through the wide TMA box)
template <int DH>
__device__ __forceinline__ void store_acc_row(uint32_t taddr, bf16* dst, bool ok, int valid) {
#pragma unroll 1
  for (int c = 0; c < DH / 16; ++c) {
    uint32_t o[16];
    LP_TMEM_LD_16(o, taddr + c * 16);
    tmem_wait_ld();
    if (ok) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (c * 16 + g * 8 < valid) {    // valid is a multiple of 8
          u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
        }
        *reinterpret_cast<uint4*>(dst + c * 16 + g * 8) = u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- dQ
template <int DH, bool CAUSAL>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const AttnBwdArgs args) {
  constexpr int Q_OFF = 0, DO_OFF = AB_HEAD, K_OFF = 2 * AB_HEAD, V_OFF = 4 * AB_HEAD, BAR_OFF = 6 * AB_HEAD;
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DQ = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* k_full = q_full + 1;     // [2]
  uint64_t* k_empty = k_full + 2;    // [2]
  uint64_t* v_full = k_empty + 2;    // [2]
  uint64_t* v_empty = v_full + 2;    // [2]
  uint64_t* s_full = v_empty + 2;    // [2]  pass 1: S_j ready in buffer j & 1
  uint64_t* s_done = s_full + 2;     // [2]  pass 1: the 128 row threads have read buffer j & 1
  uint64_t* sdp_full = s_done + 2;   // pass 2: S_j and dP_j ready
  uint64_t* ds_full = sdp_full + 1;  // pass 2: dS_j stored (128 arrivals)
  uint64_t* dq_done = ds_full + 1;   // pass 2: dQ += dS_j K_j retired
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(dq_done + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // grid = (heads, tiles, batch), heads fastest: the heaviest causal tiles of EVERY head are dispatched first (see lp_attn3.cu)
  const int q_tile = (int)gridDim.y - 1 - (int)blockIdx.y;
  const int head = blockIdx.x, batch = blockIdx.z;
  const int kv_head = head / args.group;
  const int q0 = q_tile * AB_T;
  const int T = args.T;
  int n_kv = (T + AB_T - 1) / AB_T;
  if (CAUSAL) n_kv = min(n_kv, q_tile + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_done[i], 128);
    }
    mbar_init(sdp_full, 1);
    mbar_init(ds_full, 128);
    mbar_init(dq_done, 1);
    fence_barrier_init();
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * AB_HEAD);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        tma_load_3d(smem + Q_OFF + b * AB_BOX, &tmQKV, q_full, args.q_off + head * args.hs + b * 64, q0, batch);
        tma_load_3d(smem + DO_OFF + b * AB_BOX, &tmDO, q_full, head * args.dho + b * 64, q0, batch);
      }
      int kt = 0;   // K tiles loaded so far (both passes share the ring)
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < n_kv; ++j, ++kt) {
          const int st = kt & 1;
          mbar_wait(&k_empty[st], ((kt >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], AB_HEAD);
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_3d(smem + K_OFF + st * AB_HEAD + b * AB_BOX, &tmQKV, &k_full[st],
                        args.k_off + kv_head * args.hs + b * 64, j * AB_T, batch);
          if (pass == 1) {
            const int vs = j & 1;
            mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&v_full[vs], AB_HEAD);
#pragma unroll
            for (int b = 0; b < 2; ++b)
              tma_load_3d(smem + V_OFF + vs * AB_HEAD + b * AB_BOX, &tmQKV, &v_full[vs],
                          args.v_off + kv_head * args.hs + b * 64, j * AB_T, batch);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t q_addr = smem_u32(smem + Q_OFF), do_addr = smem_u32(smem + DO_OFF);
      mbar_wait(q_full, 0);
      int kt = 0;
      // pass 1: S_j = Q K_j^T into the two S buffers (columns 0 and 128), for the row statistics
      for (int j = 0; j < n_kv; ++j, ++kt) {
        const int st = kt & 1;
        mbar_wait(&k_full[st], (kt >> 1) & 1);
        if (j >= 2) mbar_wait(&s_done[j & 1], ((j - 2) >> 1) & 1);
        tc_fence_after();
        mma_ss_kmajor<DH>(tmem_base + (uint32_t)(j & 1) * 128, q_addr, smem_u32(smem + K_OFF + st * AB_HEAD));
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[j & 1]);
      }
      // both buffers must have been read before pass 2 reuses their columns
      for (int b = 0; b < 2; ++b) {
        const int cnt = (n_kv + 1 - b) >> 1;   // tiles that used buffer b
        if (cnt >= 1) mbar_wait(&s_done[b], (cnt - 1) & 1);
      }
      tc_fence_after();
      // pass 2
      for (int j = 0; j < n_kv; ++j, ++kt) {
        const int st = kt & 1, vs = j & 1;
        mbar_wait(&k_full[st], (kt >> 1) & 1);
        mbar_wait(&v_full[vs], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + K_OFF + st * AB_HEAD);
        mma_ss_kmajor<DH>(tmem_base + TM_S, q_addr, k_addr);                                   // S_j  = Q  K_j^T
        mma_ss_kmajor<DH>(tmem_base + TM_DP, do_addr, smem_u32(smem + V_OFF + vs * AB_HEAD));  // dP_j = dO V_j^T
        umma_commit(sdp_full);
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
        mma_ts_mnmajor<DH>(tmem_base + TM_DQ, tmem_base + TM_S, k_addr, j > 0);                // dQ += dS_j K_j
        umma_commit(&k_empty[st]);
        umma_commit(&v_empty[vs]);
        umma_commit(dq_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ one thread per query row
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int q_idx = q0 + r;
    const bool row_ok = q_idx < T;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    // ---- pass 1: log-sum-exp of the scaled scores (log2 domain), online over the kv tiles
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t s_addr = lane_addr + (uint32_t)(j & 1) * 128;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      int lim = T - j * AB_T;                            // valid columns: c < lim
      if (CAUSAL && j == q_tile) lim = min(lim, r + 1);  // and c <= r on the diagonal tile
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < AB_T / 32; ++c) {
        uint32_t v[32];
        LP_TMEM_LD_32(v, s_addr + c * 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < lim) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m, mx * args.scale_log2);
      const float neg_m = (m_new == -INFINITY) ? 0.f : -m_new;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < AB_T / 32; ++c) {
        uint32_t v[32];
        LP_TMEM_LD_32(v, s_addr + c * 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < lim) sum += bw_ex2(fmaf(__uint_as_float(v[i]), args.scale_log2, neg_m));
      }
      l = l * ((m == -INFINITY) ? 0.f : bw_ex2(m - m_new)) + sum;
      m = m_new;
      tc_fence_before();
      mbar_arrive(&s_done[j & 1]);
    }
    const float lse2 = (l > 0.f) ? m + log2f(l) : 0.f;   // l == 0 only for rows without a valid key (padding rows)
    // ---- delta = rowsum(dO * O) from global memory (two head rows of 256 or 144 bytes per thread)
    float delta = 0.f;
    if (row_ok) {
      const uint4* o4 = reinterpret_cast<const uint4*>(args.O + ((long long)batch * T + q_idx) * args.ldo + head * args.dho);
      const uint4* d4 = reinterpret_cast<const uint4*>(args.dO + ((long long)batch * T + q_idx) * args.ldo + head * args.dho);
      const int n8 = args.dho >> 3;   // 16 or 9 sixteen-byte chunks per head row
#pragma unroll 4
      for (int c = 0; c < n8; ++c) {
        const uint4 a = o4[c], b = d4[c];
        const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
        const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
        delta += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
      }
    }
    {
      const long long si = ((long long)batch * args.n_heads + head) * args.Tpad + q_idx;   // q_idx < Tpad always
      args.lse2[si] = lse2;
      args.delta[si] = delta;
    }
    // ---- pass 2: dS_j = scale * P_j * (dP_j - delta) -> bf16 over S's columns
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(sdp_full, j & 1);
      if (j > 0) mbar_wait(dq_done, (j - 1) & 1);   // keeps this barrier's phases consumed one by one
      tc_fence_after();
      int lim = T - j * AB_T;
      if (CAUSAL && j == q_tile) lim = min(lim, r + 1);
#pragma unroll 1
      for (int c = 0; c < AB_T / 32; ++c) {
        uint32_t s[32], d[32], pk[16];
        LP_TMEM_LD_32(s, lane_addr + TM_S + c * 32);
        LP_TMEM_LD_32(d, lane_addr + TM_DP + c * 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float g0 = 0.f, g1 = 0.f;
          if (c * 32 + i < lim)
            g0 = bw_ex2(fmaf(__uint_as_float(s[i]), args.scale_log2, -lse2)) * (__uint_as_float(d[i]) - delta) * args.scale;
          if (c * 32 + i + 1 < lim)
            g1 = bw_ex2(fmaf(__uint_as_float(s[i + 1]), args.scale_log2, -lse2)) * (__uint_as_float(d[i + 1]) - delta) *
                 args.scale;
          pk[i >> 1] = pack_bf16x2(g0, g1);
        }
        // chunk c (32 values = 16 packed columns) lands on S columns 16c .. 16c+15, below every unread S chunk (>= 32(c+1))
        LP_TMEM_ST_16(pk, lane_addr + TM_S + c * 16);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    // ---- epilogue: dQ row -> global
    mbar_wait(dq_done, (n_kv - 1) & 1);
    tc_fence_after();
    store_acc_row<DH>(lane_addr + TM_DQ,
                      args.dqkv + ((long long)batch * T + q_idx) * args.ldd + args.q_off + (long long)head * args.hs, row_ok, DH);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------- dK, dV
template <int DH, bool CAUSAL>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const AttnBwdArgs args) {
  constexpr int K_OFF = 0, V_OFF = AB_HEAD, Q_OFF = 2 * AB_HEAD, DO_OFF = 4 * AB_HEAD, STAT_OFF = 6 * AB_HEAD;
  constexpr int BAR_OFF = STAT_OFF + 2 * 2 * AB_T * 4;
  constexpr uint32_t TM_ST = 0, TM_DPT = 128, TM_DV = 256, TM_DK = 384;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stat = reinterpret_cast<float*>(smem + STAT_OFF);   // [stage][lse2 | delta][128]
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* qd_full = kv_full + 1;    // [2]
  uint64_t* qd_empty = qd_full + 2;   // [2]
  uint64_t* st_full = qd_empty + 2;   // S^T and dP^T ready
  uint64_t* pd_full = st_full + 1;    // P^T and dS^T stored (128 arrivals)
  uint64_t* acc_done = pd_full + 1;   // dV / dK accumulation of the step retired
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int kv_tile = blockIdx.y;     // tile 0 sees every q tile: heaviest first, over all kv heads (grid = (kv heads, tiles, batch))
  const int kv_head = blockIdx.x, batch = blockIdx.z;
  const int kv0 = kv_tile * AB_T;
  const int T = args.T;
  const int n_q = (T + AB_T - 1) / AB_T;
  const int i0 = CAUSAL ? kv_tile : 0;
  const int n_i = n_q - i0;
  const int n_steps = args.group * n_i;   // step t: q head kv_head * group + t / n_i, q tile i0 + t % n_i

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
    }
    mbar_init(st_full, 1);
    mbar_init(pd_full, 128);
    mbar_init(acc_done, 1);
    fence_barrier_init();
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * AB_HEAD);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        tma_load_3d(smem + K_OFF + b * AB_BOX, &tmQKV, kv_full, args.k_off + kv_head * args.hs + b * 64, kv0, batch);
        tma_load_3d(smem + V_OFF + b * AB_BOX, &tmQKV, kv_full, args.v_off + kv_head * args.hs + b * 64, kv0, batch);
      }
      for (int t = 0; t < n_steps; ++t) {
        const int st = t & 1;
        const int head = kv_head * args.group + t / n_i, qi = i0 + t % n_i;
        mbar_wait(&qd_empty[st], ((t >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qd_full[st], 2 * AB_HEAD);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          tma_load_3d(smem + Q_OFF + st * AB_HEAD + b * AB_BOX, &tmQKV, &qd_full[st], args.q_off + head * args.hs + b * 64,
                      qi * AB_T, batch);
          tma_load_3d(smem + DO_OFF + st * AB_HEAD + b * AB_BOX, &tmDO, &qd_full[st], head * args.dho + b * 64, qi * AB_T, batch);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t k_addr = smem_u32(smem + K_OFF), v_addr = smem_u32(smem + V_OFF);
      mbar_wait(kv_full, 0);
      for (int t = 0; t < n_steps; ++t) {
        const int st = t & 1;
        mbar_wait(&qd_full[st], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + Q_OFF + st * AB_HEAD), do_addr = smem_u32(smem + DO_OFF + st * AB_HEAD);
        mma_ss_kmajor<DH>(tmem_base + TM_ST, k_addr, q_addr);      // S^T  = K  Q^T   (kv rows on the lanes)
        mma_ss_kmajor<DH>(tmem_base + TM_DPT, v_addr, do_addr);    // dP^T = V dO^T
        umma_commit(st_full);
        mbar_wait(pd_full, t & 1);
        tc_fence_after();
        mma_ts_mnmajor<DH>(tmem_base + TM_DV, tmem_base + TM_ST, do_addr, t > 0);   // dV += P^T  dO
        mma_ts_mnmajor<DH>(tmem_base + TM_DK, tmem_base + TM_DPT, q_addr, t > 0);   // dK += dS^T Q
        umma_commit(&qd_empty[st]);
        umma_commit(acc_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ one thread per kv row
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int kv_idx = kv0 + r;
    const bool kv_ok = kv_idx < T;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    for (int t = 0; t < n_steps; ++t) {
      const int head = kv_head * args.group + t / n_i, qi = i0 + t % n_i;
      float* st_lse = stat + (t & 1) * 2 * AB_T;
      float* st_del = st_lse + AB_T;
      {
        const long long si = ((long long)batch * args.n_heads + head) * args.Tpad + qi * AB_T + r;
        st_lse[r] = args.lse2[si];
        st_del[r] = args.delta[si];
      }
      row_threads_barrier();   // the four row warps only; the other stage is still being read
      mbar_wait(st_full, t & 1);
      if (t > 0) mbar_wait(acc_done, (t - 1) & 1);
      tc_fence_after();
      const int q_base = qi * AB_T;
#pragma unroll 1
      for (int c = 0; c < AB_T / 32; ++c) {
        uint32_t s[32], d[32], pp[16], pd[16];
        LP_TMEM_LD_32(s, lane_addr + TM_ST + c * 32);
        LP_TMEM_LD_32(d, lane_addr + TM_DPT + c * 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = 0.f, p1 = 0.f, g0 = 0.f, g1 = 0.f;
          const int c0 = c * 32 + i, qa = q_base + c0;
          if (kv_ok && qa < T && (!CAUSAL || qa >= kv_idx)) {
            p0 = bw_ex2(fmaf(__uint_as_float(s[i]), args.scale_log2, -st_lse[c0]));
            g0 = p0 * (__uint_as_float(d[i]) - st_del[c0]) * args.scale;
          }
          if (kv_ok && qa + 1 < T && (!CAUSAL || qa + 1 >= kv_idx)) {
            p1 = bw_ex2(fmaf(__uint_as_float(s[i + 1]), args.scale_log2, -st_lse[c0 + 1]));
            g1 = p1 * (__uint_as_float(d[i + 1]) - st_del[c0 + 1]) * args.scale;
          }
          pp[i >> 1] = pack_bf16x2(p0, p1);
          pd[i >> 1] = pack_bf16x2(g0, g1);
        }
        LP_TMEM_ST_16(pp, lane_addr + TM_ST + c * 16);
        LP_TMEM_ST_16(pd, lane_addr + TM_DPT + c * 16);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(pd_full);
    }
    mbar_wait(acc_done, (n_steps - 1) & 1);
    tc_fence_after();
    bf16* row = args.dqkv + ((long long)batch * T + kv_idx) * args.ldd + (long long)kv_head * args.hs;
    store_acc_row<DH>(lane_addr + TM_DV, row + args.v_off, kv_ok, args.dho);   // columns dho .. DH-1 (padding): zeros
    store_acc_row<DH>(lane_addr + TM_DK, row + args.k_off, kv_ok, DH);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lp

#ifndef LP_EMULATE
using namespace lp;

extern "C" size_t lp_attention_bwd_workspace_bytes(int B, int T, int n_heads) {
  const size_t tpad = (size_t)((T + AB_T - 1) / AB_T) * AB_T;
  return 2 * (size_t)B * (size_t)n_heads * tpad * sizeof(float) + 256;
}

template <int DH, bool CAUSAL>
static int launch_attn_bwd(const CUtensorMap& tq, const CUtensorMap& td, const AttnBwdArgs& a, int n_kv_heads,
                           cudaStream_t stream) {
  constexpr int SMEM_DQ = 6 * AB_HEAD + 256 + 1024;
  constexpr int SMEM_DKV = 6 * AB_HEAD + 2 * 2 * AB_T * 4 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<DH, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ));
    LP_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<DH, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV));
    configured = true;
  }
  const int tiles = (a.T + AB_T - 1) / AB_T;
  attn_bwd_dq_kernel<DH, CAUSAL><<<dim3(a.n_heads, tiles, a.B), AB_THREADS, SMEM_DQ, stream>>>(tq, td, a);
  LP_CHECK_CUDA(cudaGetLastError());
  attn_bwd_dkv_kernel<DH, CAUSAL><<<dim3(n_kv_heads, tiles, a.B), AB_THREADS, SMEM_DKV, stream>>>(tq, td, a);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_attention_bwd(const void* qkv, long long ld, const void* out, const void* dout, long long ldo, int B, int T,
                                int n_heads, int n_kv_heads, int head_dim, int head_stride, int q_off, int k_off, int v_off,
                                int causal, float scale, void* dqkv, long long ldd, void* workspace, size_t ws_bytes,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(qkv && out && dout && dqkv && workspace, "lp_attention_bwd: null argument");
  LP_REQUIRE(B > 0 && T > 0 && n_heads > 0 && n_kv_heads > 0 && n_heads % n_kv_heads == 0,
             "lp_attention_bwd: bad sizes B=%d T=%d heads=%d/%d", B, T, n_heads, n_kv_heads);
  LP_REQUIRE((head_dim == 72 && head_stride == 80) || (head_dim == 128 && head_stride == 128),
             "lp_attention_bwd: supported (head_dim, head_stride) are (72, 80) [zero-padded] and (128, 128); got (%d, %d)",
             head_dim, head_stride);
  LP_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ldd % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
             "lp_attention_bwd: leading dimensions and offsets must be multiples of 8 elements");
  LP_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(dout) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
             "lp_attention_bwd: buffers must be 16-byte aligned");
  LP_REQUIRE(ws_bytes >= lp_attention_bwd_workspace_bytes(B, T, n_heads), "lp_attention_bwd: workspace too small");
  AttnBwdArgs a;
  a.T = T; a.B = B; a.n_heads = n_heads; a.group = n_heads / n_kv_heads;
  a.Tpad = (T + AB_T - 1) / AB_T * AB_T;
  a.q_off = q_off; a.k_off = k_off; a.v_off = v_off;
  a.hs = head_stride;
  a.dho = head_dim;
  a.O = reinterpret_cast<const bf16*>(out);
  a.dO = reinterpret_cast<const bf16*>(dout);
  a.ldo = ldo;
  a.dqkv = reinterpret_cast<bf16*>(dqkv);
  a.ldd = ldd;
  a.lse2 = reinterpret_cast<float*>(workspace);
  a.delta = a.lse2 + (size_t)B * n_heads * a.Tpad;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  CUtensorMap tq, td;
  int rc = make_tmap_bf16_3d(&tq, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld, (uint64_t)T * ld, 64, 128, 1);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&td, dout, (uint64_t)ldo, (uint64_t)T, (uint64_t)B, (uint64_t)ldo, (uint64_t)T * ldo, 64, 128, 1);
  if (rc) return rc;
  note_launch(2);
  if (head_dim == 72) {
    LP_REQUIRE(!causal, "lp_attention_bwd: head_dim 72 is the (non-causal) vision-tower configuration");
    return launch_attn_bwd<80, false>(tq, td, a, n_kv_heads, stream);
  }
  return causal ? launch_attn_bwd<128, true>(tq, td, a, n_kv_heads, stream)
                : launch_attn_bwd<128, false>(tq, td, a, n_kv_heads, stream);
}
#endif  // !LP_EMULATE
