// leopard_b200 — model-level entry points: the layer loops of the prefill path, enqueued on one stream.
//
//   lp_vit_forward      SigLIP-SO400M/14-364 tower over N crops, ln_post fused with the pixel-shuffle store
//                       (HF `self.vision_tower(pixel_values).last_hidden_state`, evaluations/models/
//                        llava_multiimg_siglip_anyres.py:268,273; Megatron Idefics2ViTModel.forward,
//                        Pai-Megatron-Patch/megatron_patch/model/idefics2/idefics_vision_tower.py:104-178)
//   lp_project          the 2-layer MLP projector on the already-shuffled rows
//                       (myLlavaMultiModalProjector.forward, EV/...anyres.py:187-192; vlm_model.py:468-475)
//   lp_decoder_prefill  Llama-3.1-8B / Mistral-7B stack + final norm + lm_head over packed sequences
//                       (HF `self.language_model(...)` EV/...anyres.py:322-333; Megatron ParallelTransformer.forward,
//                        megatron_patch/model/llava/transformer.py:1747-1860 and parallel_lm_logits vlm_model.py:32-60)
// No allocation, no host synchronisation: scratch comes from the caller's workspace.
#include "lp_host.h"
#include "../../include/leopard_b200.h"
#include <cuda_bf16.h>

using namespace lp;

// attention kernel used by the layer loops: lp_attention_fwd `variant` (default 336 = v3 kernel with 64-key S halves,
// one MMA issuer per query tile and the exp-section turn protocol, persistent form (lp_attn5.cu) for the tower's T = 676;
// 80 = v3 everywhere, 2 = v2 two-tile kernel)
static int g_attn_variant = 336;
extern "C" int lp_model_set_attention_variant(int variant) {
  const int old = g_attn_variant;
  g_attn_variant = variant;
  return old;
}

// norms (and RoPE) fused into the GEMM epilogues when the weight structs carry the gain-folded matrices (lp_gemm_bf16_ex)
static int g_fused_norm = 1;
extern "C" int lp_model_set_fused_norm(int enable) {
  const int old = g_fused_norm;
  g_fused_norm = enable ? 1 : 0;
  return old;
}

namespace {
// row-statistics buffer handed from a producing GEMM epilogue to the consuming one: float2 [rows, parts]
struct RowStats {
  float* buf;
  int parts;
};
constexpr int kMaxStatParts = 32;   // producers use 256-wide tiles (produce_stats): N <= 4096 -> 2 x 16 partials per row

inline lp_gemm_desc gemm_desc(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N,
                              int K, const float* bias, const void* residual, long long ldr, int res_row_mod, int act) {
  lp_gemm_desc d = {};
  d.A = A; d.lda = lda; d.W = W; d.ldw = ldw; d.C = C; d.ldc = ldc;
  d.M = M; d.N = N; d.K = K;
  d.bias = bias; d.residual = residual; d.ldr = ldr; d.res_row_mod = res_row_mod; d.act = act;
  return d;
}
inline void consume_stats(lp_gemm_desc& d, int mode, const RowStats& st, const float* colsum, float eps) {
  d.norm_mode = mode; d.stats_in = st.buf; d.stats_in_parts = st.parts; d.colsum = colsum; d.norm_eps = eps;
}
inline void produce_stats(lp_gemm_desc& d, RowStats& st) {
  // fixed tile width: the column partition of a row's partial sums (and with it the consumer's summation order) must not depend
  // on M, or a request packed with others would no longer be bit-identical to the same request alone
  d.block_n = 256;
  st.parts = 2 * ((d.N + d.block_n - 1) / d.block_n);
  d.stats_out = st.buf; d.stats_out_parts = st.parts;
}

struct Carver {
  uint8_t* p;
  size_t used = 0, cap;
  Carver(void* base, size_t bytes) : p(reinterpret_cast<uint8_t*>(base)), cap(bytes) {}
  void* take(size_t bytes) {
    size_t off = (used + 255) & ~size_t(255);
    used = off + bytes;
    return p ? p + off : nullptr;
  }
};
inline size_t bf(size_t n) { return n * 2; }
}  // namespace

// ------------------------------------------------------------------------------------------------ ViT
static size_t vit_ws(const lp_vit_dims* d, int N, Carver* c, void** A, void** x, void** h, void** qkv, void** att,
                     void** mlp, float** st0 = nullptr, float** st1 = nullptr) {
  const size_t rows = (size_t)N * d->tokens;
  Carver tmp(nullptr, 0);
  Carver& cv = c ? *c : tmp;
  void* a = cv.take(bf(rows * d->patch_k_pad));
  void* xx = cv.take(bf(rows * d->hidden));
  void* hh = cv.take(bf(rows * d->hidden));
  void* q = cv.take(bf(rows * 3 * d->heads * d->head_stride));
  void* at = cv.take(bf(rows * d->hidden));
  void* m = cv.take(bf(rows * d->ffn));
  float* s0 = (float*)cv.take(rows * kMaxStatParts * 8);
  float* s1 = (float*)cv.take(rows * kMaxStatParts * 8);
  if (A) { *A = a; *x = xx; *h = hh; *qkv = q; *att = at; *mlp = m; }
  if (st0) { *st0 = s0; *st1 = s1; }
  return cv.used + 256;
}

static int vit_attention(const lp_vit_dims* d, void* qkv, void* att, const int* cu, int n_seqs, int uniform_T, void* stream) {
  const int qkv_w = 3 * d->heads * d->head_stride;
  const float scale = 1.0f / sqrtf((float)d->head_dim);
  int rc;
#define LP_TRY(call) do { rc = (call); if (rc) return rc; } while (0)
  if (cu == nullptr) {
    LP_TRY(lp_attention_fwd(qkv, qkv_w, n_seqs, uniform_T, d->heads, d->heads, d->head_dim, d->head_stride, 0,
                            d->heads * d->head_stride, 2 * d->heads * d->head_stride, 0, scale, att, d->hidden,
                            g_attn_variant, stream));
  } else {
    for (int s = 0; s < n_seqs; ++s) {  // one launch per image (lengths differ)
      const int r0 = cu[s], T = cu[s + 1] - r0;
      LP_TRY(lp_attention_fwd((const __nv_bfloat16*)qkv + (size_t)r0 * qkv_w, qkv_w, 1, T, d->heads, d->heads,
                              d->head_dim, d->head_stride, 0, d->heads * d->head_stride,
                              2 * d->heads * d->head_stride, 0, scale, (__nv_bfloat16*)att + (size_t)r0 * d->hidden,
                              d->hidden, g_attn_variant, stream));
    }
  }
  return 0;
}

static bool vit_can_fuse(const lp_vit_dims* d, const lp_vit_weights* w) {
  if (!g_fused_norm) return false;
  for (int l = 0; l < d->layers; ++l)
    if (!w->layers[l].qkv_wf || !w->layers[l].fc1_wf) return false;
  return d->layers > 0;
}

// st_in: statistics of x's rows when the caller's producer (patch-embedding GEMM) already wrote them, else parts == 0
static int vit_layers_impl(const lp_vit_dims* d, const lp_vit_weights* w, void* x, void* h, void* qkv, void* att,
                           void* mlp, int rows, const int* cu, int n_seqs, int uniform_T, void* out, long long ldo,
                           int shuffle_side, RowStats stA, RowStats stB, void* stream) {
  const int qkv_w = 3 * d->heads * d->head_stride;
  int rc;
  if (vit_can_fuse(d, w) && stA.buf && stB.buf) {
    // LayerNorm folded into the consuming GEMM: 4 GEMMs + attention per layer, no norm launches
    if (stA.parts == 0) {
      LP_TRY(lp_row_stats(x, d->hidden, rows, d->hidden, stA.buf, stream));
      stA.parts = 1;
    }
    for (int l = 0; l < d->layers; ++l) {
      const lp_vit_layer_weights* L = &w->layers[l];
      lp_gemm_desc g = gemm_desc(x, d->hidden, L->qkv_wf, d->hidden, qkv, qkv_w, rows, qkv_w, d->hidden, L->qkv_c2, nullptr, 0, 0,
                                 LP_ACT_NONE);
      consume_stats(g, LP_NORM_LAYER, stA, L->qkv_c1, d->eps);
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
      LP_TRY(vit_attention(d, qkv, att, cu, n_seqs, uniform_T, stream));
      g = gemm_desc(att, d->hidden, L->out_w, d->hidden, x, d->hidden, rows, d->hidden, d->hidden, L->out_b, x, d->hidden, 0,
                    LP_ACT_NONE);
      produce_stats(g, stB);
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
      g = gemm_desc(x, d->hidden, L->fc1_wf, d->hidden, mlp, d->ffn, rows, d->ffn, d->hidden, L->fc1_c2, nullptr, 0, 0,
                    LP_ACT_GELU_TANH);
      consume_stats(g, LP_NORM_LAYER, stB, L->fc1_c1, d->eps);
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
      g = gemm_desc(mlp, d->ffn, L->fc2_w, d->ffn, x, d->hidden, rows, d->hidden, d->ffn, L->fc2_b, x, d->hidden, 0, LP_ACT_NONE);
      if (l + 1 < d->layers) produce_stats(g, stA);
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
    }
    LP_TRY(lp_layernorm(x, d->hidden, w->post_g, w->post_b, out, ldo, rows, d->hidden, d->eps, shuffle_side, stream));
    return 0;
  }
  for (int l = 0; l < d->layers; ++l) {
    const lp_vit_layer_weights* L = &w->layers[l];
    LP_TRY(lp_layernorm(x, d->hidden, L->ln1_g, L->ln1_b, h, d->hidden, rows, d->hidden, d->eps, 0, stream));
    LP_TRY(lp_gemm_bf16(h, d->hidden, L->qkv_w, d->hidden, qkv, qkv_w, rows, qkv_w, d->hidden, L->qkv_b, nullptr, 0, 0,
                        LP_ACT_NONE, 0, 0, stream));
    LP_TRY(vit_attention(d, qkv, att, cu, n_seqs, uniform_T, stream));
    LP_TRY(lp_gemm_bf16(att, d->hidden, L->out_w, d->hidden, x, d->hidden, rows, d->hidden, d->hidden, L->out_b, x,
                        d->hidden, 0, LP_ACT_NONE, 0, 0, stream));
    LP_TRY(lp_layernorm(x, d->hidden, L->ln2_g, L->ln2_b, h, d->hidden, rows, d->hidden, d->eps, 0, stream));
    LP_TRY(lp_gemm_bf16(h, d->hidden, L->fc1_w, d->hidden, mlp, d->ffn, rows, d->ffn, d->hidden, L->fc1_b, nullptr, 0,
                        0, LP_ACT_GELU_TANH, 0, 0, stream));
    LP_TRY(lp_gemm_bf16(mlp, d->ffn, L->fc2_w, d->ffn, x, d->hidden, rows, d->hidden, d->ffn, L->fc2_b, x, d->hidden,
                        0, LP_ACT_NONE, 0, 0, stream));
  }
  // ln_post; shuffle_side != 0 writes the pixel-shuffled [rows/4, 4*hidden] layout directly
  LP_TRY(lp_layernorm(x, d->hidden, w->post_g, w->post_b, out, ldo, rows, d->hidden, d->eps, shuffle_side, stream));
  return 0;
}

extern "C" size_t lp_vit_workspace_bytes(const lp_vit_dims* d, int N) {
  return vit_ws(d, N, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int lp_vit_forward(const lp_vit_dims* d, const lp_vit_weights* w, const void* pixels, int pixels_fp32, int N,
                              void* out, long long ldo, int shuffle, void* workspace, size_t ws_bytes, void* stream) {
  LP_REQUIRE(d && w && pixels && out && N > 0, "lp_vit_forward: null argument or N=%d", N);
  LP_REQUIRE(d->hidden == d->heads * d->head_dim && d->head_dim == 72 && d->head_stride == 80,
             "lp_vit_forward: SigLIP-SO400M head geometry expected (16 x 72, stride 80)");
  LP_REQUIRE(d->grid * d->grid == d->tokens && d->patch_k_pad % 8 == 0 && d->patch_k_pad >= d->patch_k,
             "lp_vit_forward: bad patch geometry");
  LP_REQUIRE(ws_bytes >= lp_vit_workspace_bytes(d, N), "lp_vit_forward: workspace too small (%zu < %zu)", ws_bytes,
             lp_vit_workspace_bytes(d, N));
  Carver cv(workspace, ws_bytes);
  void *A, *x, *h, *qkv, *att, *mlp;
  RowStats stA = {nullptr, 0}, stB = {nullptr, 0};
  vit_ws(d, N, &cv, &A, &x, &h, &qkv, &att, &mlp, &stA.buf, &stB.buf);
  const int rows = N * d->tokens;
  const int img = d->grid * 14;
  int rc;
  LP_TRY(lp_im2col_patch14(pixels, pixels_fp32, A, d->patch_k_pad, N, 3, img, img, stream));
  // patch embedding GEMM + bias + learned position embedding (row % tokens); with fused norms its epilogue also writes the
  // row statistics layer 0's LayerNorm needs
  lp_gemm_desc g = gemm_desc(A, d->patch_k_pad, w->patch_w, d->patch_k_pad, x, d->hidden, rows, d->hidden, d->patch_k_pad,
                             w->patch_b, w->pos, d->hidden, d->tokens, LP_ACT_NONE);
  if (vit_can_fuse(d, w)) produce_stats(g, stA);
  LP_TRY(lp_gemm_bf16_ex(&g, stream));
  return vit_layers_impl(d, w, x, h, qkv, att, mlp, rows, nullptr, N, d->tokens, out, ldo, shuffle ? d->grid : 0, stA, stB,
                         stream);
}

// Encoder layers + ln_post on an already embedded residual stream x [R, hidden] holding n_seqs images of possibly
// different lengths (NaViT, Idefics2: MP/model/idefics2/idefics_vision_tower.py:152-178 with padded patches REMOVED,
// which is what the reference's key-padding mask / flash-varlen unpad achieve).
extern "C" size_t lp_vit_layers_workspace_bytes(const lp_vit_dims* d, int total_rows) {
  const size_t r = (size_t)total_rows;
  return bf(r * d->hidden) * 2 + bf(r * 3 * d->heads * d->head_stride) + bf(r * d->ffn) + 2 * r * kMaxStatParts * 8 + 7 * 256 + 256;
}

extern "C" int lp_vit_layers(const lp_vit_dims* d, const lp_vit_weights* w, void* x, const int* cu_seqlens_host,
                             int n_seqs, void* out, long long ldo, void* workspace, size_t ws_bytes, void* stream) {
  LP_REQUIRE(d && w && x && out && cu_seqlens_host && n_seqs > 0, "lp_vit_layers: null argument");
  const int R = cu_seqlens_host[n_seqs];
  LP_REQUIRE(R > 0 && cu_seqlens_host[0] == 0, "lp_vit_layers: bad cu_seqlens");
  LP_REQUIRE(ws_bytes >= lp_vit_layers_workspace_bytes(d, R), "lp_vit_layers: workspace too small");
  Carver cv(workspace, ws_bytes);
  void* h = cv.take(bf((size_t)R * d->hidden));
  void* qkv = cv.take(bf((size_t)R * 3 * d->heads * d->head_stride));
  void* att = cv.take(bf((size_t)R * d->hidden));
  void* mlp = cv.take(bf((size_t)R * d->ffn));
  RowStats stA = {(float*)cv.take((size_t)R * kMaxStatParts * 8), 0}, stB = {(float*)cv.take((size_t)R * kMaxStatParts * 8), 0};
  return vit_layers_impl(d, w, x, h, qkv, att, mlp, R, cu_seqlens_host, n_seqs, 0, out, ldo, 0, stA, stB, stream);
}

// ------------------------------------------------------------------------------------------------ projector
namespace lp {
int launch_gemm_chain(const void* X, long long ldx, const void* W1, const float* b1, int act1, void* mid, long long ld_mid,
                      const void* W2, const float* b2, void* out, long long ldo, int M, int N1, int K1, int N2, cudaStream_t stream);
unsigned* gemm_flag_page();
}  // namespace lp
// both projector layers in ONE kernel (lp_gemm_chain.cu) when the GEMM flag workspace is present; 0 = two GEMM launches
static int g_chained_projector = 1;
extern "C" int lp_model_set_chained_projector(int enable) {
  const int old = g_chained_projector;
  g_chained_projector = enable ? 1 : 0;
  return old;
}

extern "C" size_t lp_project_workspace_bytes(int rows, int hidden) { return bf((size_t)rows * hidden) + 512; }

extern "C" int lp_project(const lp_proj_weights* w, const void* shuffled, int rows, int in_dim, int hidden, void* out,
                          long long ldo, void* workspace, size_t ws_bytes, void* stream) {
  LP_REQUIRE(w && shuffled && out && rows > 0, "lp_project: null argument");
  LP_REQUIRE(ws_bytes >= lp_project_workspace_bytes(rows, hidden), "lp_project: workspace too small");
  Carver cv(workspace, ws_bytes);
  void* mid = cv.take(bf((size_t)rows * hidden));
  int rc;
  if (g_chained_projector && gemm_flag_page() != nullptr && rows > 128 && in_dim % 8 == 0 && hidden % 8 == 0 && ldo % 8 == 0 &&
      rows <= 448 * 256)
    return launch_gemm_chain(shuffled, in_dim, w->w1, w->b1, w->act, mid, hidden, w->w2, w->b2, out, ldo, rows, hidden, in_dim, hidden,
                             reinterpret_cast<cudaStream_t>(stream));
  LP_TRY(lp_gemm_bf16(shuffled, in_dim, w->w1, in_dim, mid, hidden, rows, hidden, in_dim, w->b1, nullptr, 0, 0, w->act,
                      0, 0, stream));
  LP_TRY(lp_gemm_bf16(mid, hidden, w->w2, hidden, out, ldo, rows, hidden, hidden, w->b2, nullptr, 0, 0, LP_ACT_NONE, 0,
                      0, stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ decoder
static size_t dec_ws(const lp_dec_dims* d, int R, int n_seq, Carver* c, void** h, void** qkv, void** att, void** act,
                     float** cs, float** sn, void** last, float** st0 = nullptr, float** st1 = nullptr) {
  Carver tmp(nullptr, 0);
  Carver& cv = c ? *c : tmp;
  const int qkv_w = (d->heads + 2 * d->kv_heads) * d->head_dim;
  void* hh = cv.take(bf((size_t)R * d->hidden));
  void* q = cv.take(bf((size_t)R * qkv_w));
  void* at = cv.take(bf((size_t)R * d->hidden));
  void* ac = cv.take(bf((size_t)R * d->ffn));
  float* c1 = (float*)cv.take((size_t)R * (d->head_dim / 2) * 4);
  float* s1 = (float*)cv.take((size_t)R * (d->head_dim / 2) * 4);
  void* la = cv.take(bf((size_t)n_seq * d->hidden));
  float* sa = (float*)cv.take((size_t)R * kMaxStatParts * 8);
  float* sb = (float*)cv.take((size_t)R * kMaxStatParts * 8);
  if (h) { *h = hh; *qkv = q; *att = at; *act = ac; *cs = c1; *sn = s1; *last = la; }
  if (st0) { *st0 = sa; *st1 = sb; }
  return cv.used + 256;
}

extern "C" size_t lp_decoder_workspace_bytes(const lp_dec_dims* d, int total_rows, int n_seqs) {
  return dec_ws(d, total_rows, n_seqs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int lp_decoder_prefill(const lp_dec_dims* d, const lp_dec_weights* w, void* x, const int* pos_ids,
                                  const int* cu_seqlens_host, int n_seqs, void* kv_cache, long long kv_capacity,
                                  void* logits, long long ld_logits, int logits_mode, int logits_fp32,
                                  void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(d && w && x && pos_ids && cu_seqlens_host && n_seqs > 0, "lp_decoder_prefill: null argument");
  LP_REQUIRE(d->head_dim == 128 && d->heads % d->kv_heads == 0, "lp_decoder_prefill: head_dim must be 128");
  const int R = cu_seqlens_host[n_seqs];
  LP_REQUIRE(R > 0 && cu_seqlens_host[0] == 0, "lp_decoder_prefill: bad cu_seqlens");
  LP_REQUIRE(ws_bytes >= lp_decoder_workspace_bytes(d, R, n_seqs), "lp_decoder_prefill: workspace too small (%zu < %zu)",
             ws_bytes, lp_decoder_workspace_bytes(d, R, n_seqs));
  LP_REQUIRE(logits_mode == LP_LOGITS_NONE || logits != nullptr, "lp_decoder_prefill: logits buffer missing");
  LP_REQUIRE(kv_cache == nullptr || kv_capacity >= R, "lp_decoder_prefill: kv_capacity %lld < %d rows", kv_capacity, R);
  Carver cv(workspace, ws_bytes);
  void *h, *qkv, *att, *act, *last;
  float *cs, *sn;
  RowStats stA = {nullptr, 0}, stB = {nullptr, 0};
  dec_ws(d, R, n_seqs, &cv, &h, &qkv, &att, &act, &cs, &sn, &last, &stA.buf, &stB.buf);
  const int H = d->hidden, hd = d->head_dim;
  const int q_w = d->heads * hd, kv_w = d->kv_heads * hd, qkv_w = q_w + 2 * kv_w;
  const float scale = 1.0f / sqrtf((float)hd);
  bool uniform = true;
  for (int s = 1; s < n_seqs; ++s)
    uniform &= (cu_seqlens_host[s + 1] - cu_seqlens_host[s]) == (cu_seqlens_host[1] - cu_seqlens_host[0]);
  int rc;
  LP_TRY(lp_rope_table(pos_ids, w->inv_freq, cs, sn, R, hd / 2, stream));
  bool fused = g_fused_norm != 0 && d->layers > 0;
  for (int l = 0; l < d->layers; ++l) fused = fused && w->layers[l].qkv_wf && w->layers[l].gateup_wf;
  if (fused) {  // RMSNorm of the merged embeddings entering layer 0: no GEMM wrote these rows
    LP_TRY(lp_row_stats(x, H, R, H, stA.buf, stream));
    stA.parts = 1;
  }
  for (int l = 0; l < d->layers; ++l) {
    const lp_dec_layer_weights* L = &w->layers[l];
    __nv_bfloat16* kc = kv_cache ? (__nv_bfloat16*)kv_cache + (size_t)l * 2 * kv_capacity * kv_w : nullptr;
    if (fused) {
      // RMSNorm folded into the QKV GEMM (gain in the weights, rstd in the epilogue) together with rotate-half RoPE on the
      // q and k heads
      lp_gemm_desc g = gemm_desc(x, H, L->qkv_wf, H, qkv, qkv_w, R, qkv_w, H, nullptr, nullptr, 0, 0, LP_ACT_NONE);
      consume_stats(g, LP_NORM_RMS, stA, nullptr, d->eps);
      g.rope_cos = cs; g.rope_sin = sn; g.rope_heads = d->heads + d->kv_heads;
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
      // with a cache: K (already rotated) and V rows are appended by a pure copy, so cached and cache-less prefills agree bit for bit
      if (kc) LP_TRY(rope_kv_append(qkv, qkv_w, R, d->heads, d->kv_heads, hd, nullptr, nullptr, kc, kc + (size_t)kv_capacity * kv_w, kv_w, stream));
    } else {
      LP_TRY(lp_rmsnorm(x, H, L->ln1_g, h, H, R, H, d->eps, stream));
      LP_TRY(lp_gemm_bf16(h, H, L->qkv_w, H, qkv, qkv_w, R, qkv_w, H, nullptr, nullptr, 0, 0, LP_ACT_NONE, 0, 0, stream));
      // q heads and k heads are contiguous column-wise: one RoPE launch covers heads + kv_heads; with a cache the same
      // launch appends K (after RoPE) and V of rows 0..R-1: [layers][2][capacity][kv_w]
      if (kc) {
        LP_TRY(rope_kv_append(qkv, qkv_w, R, d->heads, d->kv_heads, hd, cs, sn, kc, kc + (size_t)kv_capacity * kv_w, kv_w,
                              stream));
      } else {
        LP_TRY(lp_rope_inplace(qkv, qkv_w, R, 0, d->heads + d->kv_heads, hd, cs, sn, stream));
      }
    }
    if (uniform) {
      LP_TRY(lp_attention_fwd(qkv, qkv_w, n_seqs, R / n_seqs, d->heads, d->kv_heads, hd, hd, 0, q_w, q_w + kv_w, 1,
                              scale, att, H, g_attn_variant, stream));
    } else {
      for (int s = 0; s < n_seqs; ++s) {
        const int r0 = cu_seqlens_host[s], T = cu_seqlens_host[s + 1] - r0;
        LP_TRY(lp_attention_fwd((const __nv_bfloat16*)qkv + (size_t)r0 * qkv_w, qkv_w, 1, T, d->heads, d->kv_heads, hd,
                                hd, 0, q_w, q_w + kv_w, 1, scale, (__nv_bfloat16*)att + (size_t)r0 * H, H, g_attn_variant,
                                stream));
      }
    }
    if (fused) {
      lp_gemm_desc g = gemm_desc(att, H, L->o_w, H, x, H, R, H, H, nullptr, x, H, 0, LP_ACT_NONE);
      produce_stats(g, stB);
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
      g = gemm_desc(x, H, L->gateup_wf, H, act, d->ffn, R, 2 * d->ffn, H, nullptr, nullptr, 0, 0, LP_ACT_SWIGLU);
      consume_stats(g, LP_NORM_RMS, stB, nullptr, d->eps);
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
      g = gemm_desc(act, d->ffn, L->down_w, d->ffn, x, H, R, H, d->ffn, nullptr, x, H, 0, LP_ACT_NONE);
      if (l + 1 < d->layers) produce_stats(g, stA);
      LP_TRY(lp_gemm_bf16_ex(&g, stream));
    } else {
      LP_TRY(lp_gemm_bf16(att, H, L->o_w, H, x, H, R, H, H, nullptr, x, H, 0, LP_ACT_NONE, 0, 0, stream));
      LP_TRY(lp_rmsnorm(x, H, L->ln2_g, h, H, R, H, d->eps, stream));
      LP_TRY(lp_gemm_bf16(h, H, L->gateup_w, H, act, d->ffn, R, 2 * d->ffn, H, nullptr, nullptr, 0, 0, LP_ACT_SWIGLU, 0,
                          0, stream));
      LP_TRY(lp_gemm_bf16(act, d->ffn, L->down_w, d->ffn, x, H, R, H, d->ffn, nullptr, x, H, 0, LP_ACT_NONE, 0, 0,
                          stream));
    }
  }
  if (logits_mode == LP_LOGITS_ALL) {
    LP_TRY(lp_rmsnorm(x, H, w->final_g, h, H, R, H, d->eps, stream));
    LP_TRY(lp_gemm_bf16(h, H, w->lm_head, H, logits, ld_logits, R, d->vocab, H, nullptr, nullptr, 0, 0, LP_ACT_NONE,
                        logits_fp32, 0, stream));
  } else if (logits_mode == LP_LOGITS_LAST) {
    for (int s = 0; s < n_seqs; ++s) {
      const int r = cu_seqlens_host[s + 1] - 1;
      LP_TRY(lp_rmsnorm((const __nv_bfloat16*)x + (size_t)r * H, H, w->final_g, (__nv_bfloat16*)last + (size_t)s * H, H,
                        1, H, d->eps, stream));
    }
    LP_TRY(lp_gemm_bf16(last, H, w->lm_head, H, logits, ld_logits, n_seqs, d->vocab, H, nullptr, nullptr, 0, 0,
                        LP_ACT_NONE, logits_fp32, 0, stream));
  }
  return 0;
}
