// leopard_b200 — warp-specialised tcgen05 flash attention (forward / prefill) for sm_100a.
//
//   O[b, t, h, :] = softmax( scale · Q[b,t,h,:] · K[b,:,h/g,:]^T  (+ causal mask) ) · V[b,:,h/g,:]
//
// Replaces, on Leopard's prefill path:
//   K3  SigLIP non-causal MHA, 16 heads x 72 (HF SiglipAttention; Megatron TEDotProductAttention,
//       Pai-Megatron-Patch/megatron_patch/model/llava/clip_encoder.py:58-63)            -> DH = 80 (72 zero-padded)
//   K12 decoder causal GQA attention, 32 q / 8 kv heads x 128 (flash_attn_varlen_func call at
//       megatron_patch/model/llava/transformer.py:476-515,869-877; KV replication :829-836 is NOT materialised
//       here — q head h reads kv head h / group directly)                                    -> DH = 128
//
// Q, K, V are column slices of one row-major bf16 buffer (the fused QKV GEMM output): for head h the q columns
// start at q_off + h*hs, k at k_off + (h/g)*hs, v at v_off + (h/g)*hs.  One CTA owns 128 query rows of one head.
//   warp 0     TMA producer : Q once; K_j, V_j tiles (128 rows) into 2-deep rings (128B-swizzled smem)
//   warp 1     MMA issuer   : S_j = Q·K_j^T (SS, K-major x K-major) into a double-buffered TMEM S;
//                             O += P_j·V_j with V as an MN-major B operand straight from its TMA tile
//   warps 2-5  softmax      : one thread per query row (tcgen05.ld 32x32b), online softmax in fp32 with exp2 and
//                             lazy rescaling of the TMEM-resident O (only when the row max grows by > 2^8),
//                             P_j -> bf16 -> swizzled smem (or TMEM when P_TMEM), final O / l -> global.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

constexpr int ATT_BM = 128;  // query rows per CTA
constexpr int ATT_BN = 128;  // kv rows per tile
constexpr int ATT_THREADS = 192;
constexpr int ATT_BOX_BYTES = 128 * 128;  // one TMA box: 128 rows x 64 bf16

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH, bool CAUSAL, bool P_TMEM>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQKV,
                const AttnArgs args) {
  constexpr int NB = (DH + 63) / 64;   // 64-column TMA boxes per head
  constexpr int KS_QK = DH / 16;       // K=16 steps of S = Q K^T
  constexpr int KS_PV = ATT_BN / 16;   // K=16 steps of O += P V
  constexpr int HEAD_BYTES = NB * ATT_BOX_BYTES;
  constexpr int Q_OFF = 0;
  constexpr int K_OFF = Q_OFF + HEAD_BYTES;
  constexpr int V_OFF = K_OFF + 2 * HEAD_BYTES;
  constexpr int P_OFF = V_OFF + 2 * HEAD_BYTES;
  constexpr int BAR_OFF = P_OFF + (P_TMEM ? 0 : 2 * ATT_BOX_BYTES);
  constexpr uint32_t TM_S0 = 0, TM_O = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* k_full = q_full + 1;    // [2]
  uint64_t* k_empty = k_full + 2;   // [2]
  uint64_t* v_full = k_empty + 2;   // [2]
  uint64_t* v_empty = v_full + 2;   // [2]
  uint64_t* s_full = v_empty + 2;   // [2]
  uint64_t* p_full = s_full + 2;    // [2]
  uint64_t* pv_done = p_full + 2;   // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const int q_tile = (int)gridDim.x - 1 - (int)blockIdx.x;  // heavy (late) causal tiles first
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int kv_head = head / args.group;
  const int q0 = q_tile * ATT_BM;
  const int T = args.Tk;   // key/value length (masks, tile count)
  const int Tq = args.T;   // query length (output rows)
  int n_kv = (T + ATT_BN - 1) / ATT_BN;
  if (CAUSAL) n_kv = min(n_kv, q_tile + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(pv_done, 1);
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
      mbar_arrive_expect_tx(q_full, HEAD_BYTES);
#pragma unroll
      for (int b = 0; b < NB; ++b)
        tma_load_3d(smem + Q_OFF + b * ATT_BOX_BYTES, &tmQ, q_full, args.q_off + head * args.hs + b * 64, q0, batch);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], HEAD_BYTES);
#pragma unroll
        for (int b = 0; b < NB; ++b)
          tma_load_3d(smem + K_OFF + st * HEAD_BYTES + b * ATT_BOX_BYTES, &tmQKV, &k_full[st],
                      args.k_off + kv_head * args.hs + b * 64, j * ATT_BN, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], HEAD_BYTES);
#pragma unroll
        for (int b = 0; b < NB; ++b)
          tma_load_3d(smem + V_OFF + st * HEAD_BYTES + b * ATT_BOX_BYTES, &tmQKV, &v_full[st],
                      args.v_off + kv_head * args.hs + b * 64, j * ATT_BN, batch);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, DH, 0, 1);       // P (K-major) x V (MN-major)
      const uint32_t q_addr = smem_u32(smem + Q_OFF);
      const uint32_t p_addr = smem_u32(smem + P_OFF);
      mbar_wait(q_full, 0);
      for (int j = 0; j <= n_kv; ++j) {
        if (j < n_kv) {  // S_j = Q K_j^T
          const int st = j & 1;
          mbar_wait(&k_full[st], (j >> 1) & 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(smem + K_OFF + st * HEAD_BYTES);
          const uint32_t d_s = tmem_base + TM_S0 + (uint32_t)(j & 1) * ATT_BN;
#pragma unroll
          for (int ks = 0; ks < KS_QK; ++ks) {
            const uint32_t off = (ks >> 2) * ATT_BOX_BYTES + (ks & 3) * 32;
            umma_ss(d_s, umma_desc_sw128(q_addr + off, 0, 1024), umma_desc_sw128(k_addr + off, 0, 1024), idesc_s,
                    ks > 0 ? 1u : 0u);
          }
          umma_commit(&k_empty[st]);
          umma_commit(&s_full[j & 1]);
        }
        if (j > 0) {  // O += P_{j-1} V_{j-1}
          const int jj = j - 1;
          const int st = jj & 1;
          const uint32_t ph = (jj >> 1) & 1;
          mbar_wait(&p_full[jj & 1], ph);
          mbar_wait(&v_full[st], ph);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(smem + V_OFF + st * HEAD_BYTES);
          const uint32_t d_o = tmem_base + TM_O;
#pragma unroll
          for (int ks = 0; ks < KS_PV; ++ks) {
            // V tile: rows = kv (K dim), 128 B rows along dh (MN dim); 16 kv rows = 2 KiB per K step,
            // next 64-wide dh slab (LBO) one TMA box away, 8-row groups (SBO) 1 KiB apart.
            const uint64_t vdesc = umma_desc_sw128(v_addr + ks * 2048, ATT_BOX_BYTES, 1024);
            const uint32_t acc = (jj > 0 || ks > 0) ? 1u : 0u;
            if (P_TMEM) {
              const uint32_t a_t = tmem_base + TM_S0 + (uint32_t)(jj & 1) * ATT_BN + ks * 8;  // 16 bf16 = 8 cols
              umma_ts(d_o, a_t, vdesc, idesc_o, acc);
            } else {
              const uint32_t off = (ks >> 2) * ATT_BOX_BYTES + (ks & 3) * 32;
              umma_ss(d_o, umma_desc_sw128(p_addr + off, 0, 1024), vdesc, idesc_o, acc);
            }
          }
          umma_commit(&v_empty[st]);
          umma_commit(pv_done);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int lg = warp & 3;
    const int r = lg * 32 + lane;           // row within the q tile == TMEM lane
    const int q_idx = q0 + r;               // position inside the sequence
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    float m_used = -INFINITY;               // running (possibly stale) max, log2 domain
    float l_sum = 0.f;
    uint8_t* p_row = smem + P_OFF + r * 128;
    const int rsw = r & 7;

    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * ATT_BN;
      const uint32_t s_addr = lane_addr + TM_S0 + (uint32_t)(j & 1) * ATT_BN;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const bool need_mask = (kv0 + ATT_BN > T) || (CAUSAL && j == q_tile);
      int lim = T - kv0;                                 // valid columns: c < lim
      if (CAUSAL && j == q_tile) lim = min(lim, r + 1);  // and c <= r on the diagonal tile
      // ---- pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t v[32];
        LP_TMEM_LD_32(v, s_addr + c * 32);
        tmem_wait_ld();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < lim) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_used, mx * args.scale_log2);
      // ---- previous PV must have retired before O is rescaled / the P buffer is overwritten
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      const bool grow = (m_new - m_used) > 8.0f;  // also true on the first tile (m_used = -inf)
      if (j == 0) {
        m_used = m_new;
      } else if (__any_sync(0xffffffffu, grow)) {
        float alpha = 1.0f;
        if (grow) {
          alpha = ex2f(m_used - m_new);
          m_used = m_new;
          l_sum *= alpha;
        }
#pragma unroll 1
        for (int c = 0; c < DH / 16; ++c) {
          uint32_t o[16];
          LP_TMEM_LD_16(o, lane_addr + TM_O + c * 16);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
              "%14, %15, %16};" ::"r"(lane_addr + TM_O + c * 16),
              "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]),
              "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15])
              : "memory");
        }
        tmem_wait_st();
      }
      // ---- pass 2: p = 2^(s*scale - m), row sum, P -> bf16
      const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
#pragma unroll 1
      for (int c = 0; c < ATT_BN / 32; ++c) {
        uint32_t v[32];
        LP_TMEM_LD_32(v, s_addr + c * 32);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2f(fmaf(__uint_as_float(v[i]), args.scale_log2, neg_m));
          float p1 = ex2f(fmaf(__uint_as_float(v[i + 1]), args.scale_log2, neg_m));
          if (need_mask) {
            if (c * 32 + i >= lim) p0 = 0.f;
            if (c * 32 + i + 1 >= lim) p1 = 0.f;
          }
          // accumulate the row sum from the bf16-rounded values the MMA will actually use
          const uint32_t u = pack_bf16x2(p0, p1);
          const float2 f = unpack_bf16x2(u);
          l_sum += f.x + f.y;
          pk[i >> 1] = u;
        }
        if (P_TMEM) {
          // P_j overwrites the first 64 columns of S_j (2 bf16 per 32-bit column); chunk c -> columns c*16..+15
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
              "%14, %15, %16};" ::"r"(s_addr + c * 16),
              "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]),
              "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
              : "memory");
        } else {
          // K-major 128B-swizzled tile: 16-byte chunk q of row r lives at chunk (q ^ (r & 7))
          uint8_t* box = p_row + (c >> 1) * ATT_BOX_BYTES;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(box + ((chunk ^ rsw) << 4)) =
                make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
      }
      if (P_TMEM) {
        tmem_wait_st();
        tc_fence_before();
      } else {
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        tc_fence_before();
      }
      mbar_arrive(&p_full[j & 1]);
    }

    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
    const bool row_ok = q_idx < Tq;
    bf16* op = args.O + ((long long)batch * Tq + q_idx) * args.ldo + (long long)head * args.dh_out;
#pragma unroll 1
    for (int c = 0; c < DH / 16; ++c) {
      uint32_t o[16];
      LP_TMEM_LD_16(o, lane_addr + TM_O + c * 16);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH, bool CAUSAL, bool P_TMEM>
static int launch_attn(const CUtensorMap& tmq, const CUtensorMap& tm, const AttnArgs& a, cudaStream_t stream) {
  constexpr int NB = (DH + 63) / 64;
  constexpr int SMEM = 5 * NB * ATT_BOX_BYTES + (P_TMEM ? 0 : 2 * ATT_BOX_BYTES) + 256 + 1024;
  auto kern = attn_fwd_kernel<DH, CAUSAL, P_TMEM>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid((a.T + ATT_BM - 1) / ATT_BM, a.n_heads, a.B);
  kern<<<grid, ATT_THREADS, SMEM, stream>>>(tmq, tm, a);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace lp

using namespace lp;

static int g_attn_pingpong = 0;  // default of the softmax turn protocol when the variant bits 4/5 do not force it

extern "C" int lp_attention_set_pingpong(int enable) {
  const int old = g_attn_pingpong;
  g_attn_pingpong = enable ? 1 : 0;
  return old;
}

extern "C" int lp_attention_fwd(const void* qkv, long long ld, int B, int T, int n_heads, int n_kv_heads, int head_dim,
                                int head_stride, int q_off, int k_off, int v_off, int causal, float scale, void* out,
                                long long ldo, int variant, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(B > 0 && T > 0 && n_heads > 0 && n_kv_heads > 0 && n_heads % n_kv_heads == 0,
             "lp_attention_fwd: bad sizes B=%d T=%d heads=%d/%d", B, T, n_heads, n_kv_heads);
  LP_REQUIRE((head_dim == 72 && head_stride == 80) || (head_dim == 128 && head_stride == 128),
             "lp_attention_fwd: supported (head_dim, head_stride) are (72, 80) [zero-padded] and (128, 128); got (%d, %d)",
             head_dim, head_stride);
  LP_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
             "lp_attention_fwd: ld/ldo/offsets must be multiples of 8 elements");
  LP_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "lp_attention_fwd: qkv/out must be 16-byte aligned");
  AttnArgs a;
  a.T = T; a.Tk = T; a.B = B; a.n_heads = n_heads; a.group = n_heads / n_kv_heads;
  a.q_off = q_off; a.k_off = k_off; a.v_off = v_off; a.hs = head_stride;
  a.dh_out = head_dim;
  a.O = reinterpret_cast<bf16*>(out);
  a.ldo = ldo;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.pingpong = (variant & 16) ? 1 : ((variant & 32) ? 0 : g_attn_pingpong);
  CUtensorMap tm;
  int rc = make_tmap_bf16_3d(&tm, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld, (uint64_t)T * ld, 64,
                             128, 1);
  if (rc) return rc;
  if ((variant & 256) && attn5_eligible(a, head_dim, causal)) {  // persistent v3 over (batch, head, query-tile pair) items (lp_attn5.cu)
    ProfScope prof(2, B, T, n_heads * 1000 + head_dim, stream);
    note_launch();
    return launch_attn5(tm, a, stream);
  }
  if (variant & 128) {  // v3's dataflow with one softmax warpgroup running both tiles' rows as two streams (lp_attn4.cu)
    ProfScope prof(2, B, T, n_heads * 1000 + head_dim, stream);
    note_launch();
    return launch_attn4(tm, a, head_dim, causal, (variant >> 2) & 3, stream);
  }
  if (variant & 64) {  // two query tiles, S in 64-key halves, one issuer per tile (lp_attn3.cu)
    ProfScope prof(2, B, T, n_heads * 1000 + head_dim, stream);
    note_launch();
    return launch_attn3(tm, a, head_dim, causal, (variant >> 2) & 3, stream);
  }
  if (variant & 2) {  // two query tiles per CTA (lp_attn2.cu)
    ProfScope prof(2, B, T, n_heads * 1000 + head_dim, stream);
    note_launch();
    return launch_attn2(tm, a, head_dim, causal, (variant >> 2) & 3, stream);
  }
  const bool ptm = (variant & 1) != 0;
  ProfScope prof(2, B, T, n_heads * 1000 + head_dim, stream);
  note_launch();
  if (head_dim == 72) {
    LP_REQUIRE(!causal, "lp_attention_fwd: head_dim 72 is the (non-causal) ViT configuration");
    return ptm ? launch_attn<80, false, true>(tm, tm, a, stream) : launch_attn<80, false, false>(tm, tm, a, stream);
  }
  if (causal)
    return ptm ? launch_attn<128, true, true>(tm, tm, a, stream) : launch_attn<128, true, false>(tm, tm, a, stream);
  return ptm ? launch_attn<128, false, true>(tm, tm, a, stream) : launch_attn<128, false, false>(tm, tm, a, stream);
}

// Cross-attention with separate query / key-value buffers (Idefics2 perceiver: 64 latent queries over
// concat(context, latents); megatron_patch/model/idefics2/perceiver_transformer.py:582-706): non-causal, one q tile.
extern "C" int lp_attention_cross(const void* q, long long ldq, const void* kv, long long ldkv, int B, int Tq, int Tk,
                                  int n_heads, int n_kv_heads, int head_dim, int k_off, int v_off, float scale,
                                  void* out, long long ldo, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(B > 0 && Tq > 0 && Tk > 0 && n_heads > 0 && n_kv_heads > 0 && n_heads % n_kv_heads == 0,
             "lp_attention_cross: bad sizes B=%d Tq=%d Tk=%d heads=%d/%d", B, Tq, Tk, n_heads, n_kv_heads);
  LP_REQUIRE(head_dim == 96 || head_dim == 128, "lp_attention_cross: head_dim must be 96 or 128 (got %d)", head_dim);
  LP_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
             "lp_attention_cross: strides/offsets must be multiples of 8 elements");
  LP_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "lp_attention_cross: q/kv/out must be 16-byte aligned");
  AttnArgs a;
  a.T = Tq; a.Tk = Tk; a.B = B; a.n_heads = n_heads; a.group = n_heads / n_kv_heads;
  a.q_off = 0; a.k_off = k_off; a.v_off = v_off; a.hs = head_dim;
  a.dh_out = head_dim;
  a.O = reinterpret_cast<bf16*>(out);
  a.ldo = ldo;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.pingpong = 0;
  CUtensorMap tq, tkv;
  int rc = make_tmap_bf16_3d(&tq, q, (uint64_t)ldq, (uint64_t)Tq, (uint64_t)B, (uint64_t)ldq, (uint64_t)Tq * ldq, 64, 128, 1);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tkv, kv, (uint64_t)ldkv, (uint64_t)Tk, (uint64_t)B, (uint64_t)ldkv, (uint64_t)Tk * ldkv, 64, 128,
                         1);
  if (rc) return rc;
  ProfScope prof(2, B, Tk, n_heads * 1000 + head_dim, stream);
  note_launch();
  if (head_dim == 96) return launch_attn<96, false, true>(tq, tkv, a, stream);
  return launch_attn<128, false, true>(tq, tkv, a, stream);
}
