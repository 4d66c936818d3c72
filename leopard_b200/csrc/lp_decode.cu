// leopard_b200 — single-token decode step over the prefill's KV cache (greedy generation after the prefill).
//
// The cached-decode branch of myLlavaForConditionalGeneration.forward
// (evaluations/models/llava_multiimg_siglip_anyres.py:291-320, driven by `generate(..., max_new_tokens=128,
// use_cache=True)` at :448-452): with B = 1 and no padding the rebuilt mask is all ones and the position is the
// number of cached rows.  HBM-bound: every weight is streamed once per token (M = 1 rows through
// the tcgen05 GEMM's TMA pipeline), attention is a SIMT GQA kernel reading the bf16 KV cache once.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

int rope_table_base(int pos_base, const float* inv_freq, float* cs, float* sn, int S, int half, cudaStream_t stream);

constexpr int DEC_WARPS = 8;

// one CTA per query head; warps stride over cached positions; head_dim = 128 (4 dims per lane)
__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kc, const bf16* __restrict__ vc, int n_ctx,
                   int kv_w, int group, float scale_log2, bf16* __restrict__ out) {
  const int h = blockIdx.x, g = h / group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS], s_acc[DEC_WARPS][128];
  float qv[4];
  {
    const uint2 u = *reinterpret_cast<const uint2*>(q + h * 128 + lane * 4);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    qv[0] = a.x * scale_log2; qv[1] = a.y * scale_log2; qv[2] = b.x * scale_log2; qv[3] = b.y * scale_log2;
  }
  float m = -INFINITY, l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t = warp; t < n_ctx; t += DEC_WARPS) {
    const uint2 ku = *reinterpret_cast<const uint2*>(kc + (long long)t * kv_w + g * 128 + lane * 4);
    const uint2 vu = *reinterpret_cast<const uint2*>(vc + (long long)t * kv_w + g * 128 + lane * 4);
    const float2 k0 = unpack_bf16x2(ku.x), k1 = unpack_bf16x2(ku.y);
    float s = qv[0] * k0.x + qv[1] * k0.y + qv[2] * k1.x + qv[3] * k1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    const float2 v0 = unpack_bf16x2(vu.x), v1 = unpack_bf16x2(vu.y);
    l = l * corr + p;
    acc[0] = acc[0] * corr + p * v0.x; acc[1] = acc[1] * corr + p * v0.y;
    acc[2] = acc[2] * corr + p * v1.x; acc[3] = acc[3] * corr + p * v1.y;
    m = mn;
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
  for (int i = 0; i < 4; ++i) s_acc[warp][lane * 4 + i] = acc[i];
  __syncthreads();
  if (warp == 0) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) mm = fmaxf(mm, s_m[w]);
    float ll = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) {
      const float c = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - mm);
      ll += s_l[w] * c;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] += s_acc[w][lane * 4 + i] * c;
    }
    const float inv = 1.0f / ll;
    uint2 u;
    u.x = pack_bf16x2(o[0] * inv, o[1] * inv);
    u.y = pack_bf16x2(o[2] * inv, o[3] * inv);
    *reinterpret_cast<uint2*>(out + h * 128 + lane * 4) = u;
  }
}

}  // namespace lp

using namespace lp;

static inline size_t bf(size_t n) { return n * 2; }

extern "C" size_t lp_decode_workspace_bytes(const lp_dec_dims* d) {
  const size_t qkv_w = (size_t)(d->heads + 2 * d->kv_heads) * d->head_dim;
  return 256 * 8 + bf(2 * (size_t)d->hidden) * 2 + bf(qkv_w) + bf((size_t)d->ffn) + (size_t)d->head_dim * 4 + 64 * 4;
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
      note_launch();
      decode_attn_kernel<<<d->heads, DEC_WARPS * 32, 0, stream>>>(qkv, kc, vc, position + 1, kv_w,
                                                                 d->heads / d->kv_heads, scale_log2, att);
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
