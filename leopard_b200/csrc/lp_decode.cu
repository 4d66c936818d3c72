// leopard_b200 — single-token decode step over the prefill's KV cache (greedy generation after the prefill).
//
// The cached-decode branch of myLlavaForConditionalGeneration.forward
// (evaluations/models/llava_multiimg_siglip_anyres.py:291-320, driven by `generate(..., max_new_tokens=128,
// use_cache=True)` at :448-452): with B = 1 and no padding the rebuilt mask is all ones and the position is the
// number of cached rows.  HBM-bound: every weight is streamed once per token (M = 1 rows through
// the tcgen05 GEMM's TMA pipeline), attention is a split-K SIMT GQA kernel (heads x 16 key ranges, 4 keys in flight per
// warp) reading the bf16 KV cache once, plus a tiny combine pass.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

int rope_table_base(int pos_base, const float* inv_freq, float* cs, float* sn, int S, int half, cudaStream_t stream);

constexpr int DEC_WARPS = 8;
constexpr int DEC_SPLITS = 16;  // key ranges per head -> heads * DEC_SPLITS CTAs (>= 1 wave on 148 SMs)

// Split-K decode attention, pass 1: CTA (head h, split s) scans keys [s*chunk, (s+1)*chunk) of kv head h/group.
// A warp handles FOUR keys per step (8 lanes x 16 dims each, two 16-byte loads per lane, 3-step shuffle reduce) so that
// four independent load->dot->exp chains are in flight; partial (m, l, acc[128]) go to `part` for the combine pass.
__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_attn_partial_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kc, const bf16* __restrict__ vc,
                           int n_ctx, int kv_w, int group, float scale_log2, float* __restrict__ part) {
  const int h = blockIdx.x, split = blockIdx.y, g = h / group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, l8 = lane & 7;  // key slot within the warp step, 16-dim slice of the head
  const unsigned gmask = 0xFFu << (sub * 8);
  const int chunk = (n_ctx + DEC_SPLITS - 1) / DEC_SPLITS;
  const int t_begin = split * chunk, t_end = min(n_ctx, t_begin + chunk);
  __shared__ float s_m[DEC_WARPS * 4], s_l[DEC_WARPS * 4], s_acc[DEC_WARPS * 4][128];
  float qv[16];
  {
    const uint4 u0 = *reinterpret_cast<const uint4*>(q + h * 128 + l8 * 16);
    const uint4 u1 = *reinterpret_cast<const uint4*>(q + h * 128 + l8 * 16 + 8);
    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = unpack_bf16x2(w[i]);
      qv[2 * i] = f.x * scale_log2;
      qv[2 * i + 1] = f.y * scale_log2;
    }
  }
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int t = t_begin + warp * 4 + sub; t < t_end; t += DEC_WARPS * 4) {
    const bf16* kp = kc + (long long)t * kv_w + g * 128 + l8 * 16;
    const bf16* vp = vc + (long long)t * kv_w + g * 128 + l8 * 16;
    const uint4 k0 = *reinterpret_cast<const uint4*>(kp), k1 = *reinterpret_cast<const uint4*>(kp + 8);
    const uint4 v0 = *reinterpret_cast<const uint4*>(vp), v1 = *reinterpret_cast<const uint4*>(vp + 8);
    const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = unpack_bf16x2(kw[i]);
      s += qv[2 * i] * f.x + qv[2 * i + 1] * f.y;
    }
    // only the 8 lanes that share this key take part: key slots of one warp run different trip counts at the tail
    s += __shfl_xor_sync(gmask, s, 4);
    s += __shfl_xor_sync(gmask, s, 2);
    s += __shfl_xor_sync(gmask, s, 1);
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    l = l * corr + p;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = unpack_bf16x2(vw[i]);
      acc[2 * i] = acc[2 * i] * corr + p * f.x;
      acc[2 * i + 1] = acc[2 * i + 1] * corr + p * f.y;
    }
    m = mn;
  }
  // note: a lane whose key slots were all beyond t_end keeps m = -inf, l = 0 (handled by the combine weights)
  const int slot = warp * 4 + sub;
  if (l8 == 0) { s_m[slot] = m; s_l[slot] = l; }
#pragma unroll
  for (int i = 0; i < 16; ++i) s_acc[slot][l8 * 16 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int d = threadIdx.x;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < DEC_WARPS * 4; ++w) mm = fmaxf(mm, s_m[w]);
    float ll = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS * 4; ++w) {
      const float c = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - mm);
      ll += s_l[w] * c;
      o += s_acc[w][d] * c;
    }
    float* pp = part + ((long long)h * DEC_SPLITS + split) * 130;
    pp[2 + d] = o;
    if (d == 0) { pp[0] = mm; pp[1] = ll; }
  }
}

// pass 2: merge the DEC_SPLITS partials of each head
__global__ void __launch_bounds__(128)
decode_attn_combine_kernel(const float* __restrict__ part, bf16* __restrict__ out) {
  const int h = blockIdx.x, d = threadIdx.x;
  const float* pp = part + (long long)h * DEC_SPLITS * 130;
  float mm = -INFINITY;
#pragma unroll
  for (int s = 0; s < DEC_SPLITS; ++s) mm = fmaxf(mm, pp[s * 130]);
  float ll = 0.f, o = 0.f;
#pragma unroll
  for (int s = 0; s < DEC_SPLITS; ++s) {
    const float m = pp[s * 130];
    const float c = (m == -INFINITY) ? 0.f : exp2f(m - mm);
    ll += pp[s * 130 + 1] * c;
    o += pp[s * 130 + 2 + d] * c;
  }
  out[h * 128 + d] = __float2bfloat16(o / ll);
}

}  // namespace lp

using namespace lp;

static inline size_t bf(size_t n) { return n * 2; }

extern "C" size_t lp_decode_workspace_bytes(const lp_dec_dims* d) {
  const size_t qkv_w = (size_t)(d->heads + 2 * d->kv_heads) * d->head_dim;
  return 256 * 9 + bf(2 * (size_t)d->hidden) * 2 + bf(qkv_w) + bf((size_t)d->ffn) + (size_t)d->head_dim * 4 + 64 * 4 +
         (size_t)d->heads * DEC_SPLITS * 130 * 4;
}

extern "C" int lp_decode_step(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table, long long ld_embed,
                              const long long* token_id, int position, void* kv_cache, long long kv_capacity,
                              void* logits, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(d && w && embed_table && token_id && kv_cache && logits && workspace, "lp_decode_step: null argument");
  LP_REQUIRE(d->head_dim == 128, "lp_decode_step: head_dim must be 128");
  LP_REQUIRE(position >= 0 && position < kv_capacity, "lp_decode_step: position %d outside the cache capacity %lld",
             position, kv_capacity);
  LP_REQUIRE(ws_bytes >= lp_decode_workspace_bytes(d), "lp_decode_step: workspace too small");
  const int H = d->hidden, hd = d->head_dim;
  const int q_w = d->heads * hd, kv_w = d->kv_heads * hd, qkv_w = q_w + 2 * kv_w;
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t n) { void* r = p; p += (n + 255) & ~size_t(255); return r; };
  __nv_bfloat16* x = (__nv_bfloat16*)take(bf(H));
  __nv_bfloat16* h = (__nv_bfloat16*)take(bf(H));
  __nv_bfloat16* qkv = (__nv_bfloat16*)take(bf(qkv_w));
  __nv_bfloat16* att = (__nv_bfloat16*)take(bf(H));
  __nv_bfloat16* act = (__nv_bfloat16*)take(bf(d->ffn));
  float* cs = (float*)take(hd / 2 * 4);
  float* sn = (float*)take(hd / 2 * 4);
  int* plan = (int*)take(16);
  float* part = (float*)take((size_t)d->heads * DEC_SPLITS * 130 * 4);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  int rc;
#define LP_TRY(call) do { rc = (call); if (rc) return rc; } while (0)
  // new token's embedding row (ids stay on the device: no host round trip of the sampled token is needed here)
  LP_TRY(lp_embed_merge(token_id, 1, -(1LL << 62), 1, embed_table, ld_embed, d->vocab, nullptr, 0, 0, x, H, 1, H, plan,
                        stream));
  LP_TRY(rope_table_base(position, w->inv_freq, cs, sn, 1, hd / 2, stream));
  for (int l = 0; l < d->layers; ++l) {
    const lp_dec_layer_weights* L = &w->layers[l];
    __nv_bfloat16* kc = (__nv_bfloat16*)kv_cache + (size_t)l * 2 * kv_capacity * kv_w;
    __nv_bfloat16* vc = kc + (size_t)kv_capacity * kv_w;
    LP_TRY(lp_rmsnorm(x, H, L->ln1_g, h, H, 1, H, d->eps, stream));
    LP_TRY(lp_gemm_bf16(h, H, L->qkv_w, H, qkv, qkv_w, 1, qkv_w, H, nullptr, nullptr, 0, 0, LP_ACT_NONE, 0, 128,
                        stream));
    LP_TRY(lp_rope_inplace(qkv, qkv_w, 1, 0, d->heads + d->kv_heads, hd, cs, sn, stream));
    LP_CHECK_CUDA(cudaMemcpyAsync(kc + (size_t)position * kv_w, qkv + q_w, bf(kv_w), cudaMemcpyDeviceToDevice, stream));
    LP_CHECK_CUDA(cudaMemcpyAsync(vc + (size_t)position * kv_w, qkv + q_w + kv_w, bf(kv_w), cudaMemcpyDeviceToDevice,
                                  stream));
    {
      ProfScope prof(8, position + 1, d->heads, 0, stream);
      note_launch(2);
      decode_attn_partial_kernel<<<dim3(d->heads, DEC_SPLITS), DEC_WARPS * 32, 0, stream>>>(
          qkv, kc, vc, position + 1, kv_w, d->heads / d->kv_heads, scale_log2, part);
      decode_attn_combine_kernel<<<d->heads, 128, 0, stream>>>(part, att);
      LP_CHECK_CUDA(cudaGetLastError());
    }
    LP_TRY(lp_gemm_bf16(att, H, L->o_w, H, x, H, 1, H, H, nullptr, x, H, 0, LP_ACT_NONE, 0, 128, stream));
    LP_TRY(lp_rmsnorm(x, H, L->ln2_g, h, H, 1, H, d->eps, stream));
    LP_TRY(lp_gemm_bf16(h, H, L->gateup_w, H, act, d->ffn, 1, 2 * d->ffn, H, nullptr, nullptr, 0, 0, LP_ACT_SWIGLU, 0,
                        128, stream));
    LP_TRY(lp_gemm_bf16(act, d->ffn, L->down_w, d->ffn, x, H, 1, H, d->ffn, nullptr, x, H, 0, LP_ACT_NONE, 0, 128,
                        stream));
  }
  LP_TRY(lp_rmsnorm(x, H, w->final_g, h, H, 1, H, d->eps, stream));
  LP_TRY(lp_gemm_bf16(h, H, w->lm_head, H, logits, d->vocab, 1, d->vocab, H, nullptr, nullptr, 0, 0, LP_ACT_NONE, 1, 128,
                      stream));
  return 0;
}
