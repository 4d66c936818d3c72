/* leopard_b200 — C ABI of the B200-native Leopard multi-image prefill path.
 *
 * This is the drop-in boundary under the reference's Python plugin surface: the shims in
 * leopard_b200/hf.py and leopard_b200/megatron.py keep the reference's class/function names and call ONLY the
 * functions below through ctypes.  The reference has no FFI of its own (it calls torch / transformers /
 * TransformerEngine / flash-attn); each entry point cites the reference call site whose arithmetic it replaces
 * (paths relative to the reference checkout: EV/ = evaluations/models/, MP/ = Pai-Megatron-Patch/megatron_patch/,
 * MLM/ = Pai-Megatron-Patch/Megatron-LM-240603/megatron/).
 *
 * Conventions: plain pointers and sizes, no torch types.  Every pointer is a DEVICE pointer unless it says host.
 * bf16 tensors are row-major with the last dimension contiguous.  Nothing allocates or synchronises; all work is
 * enqueued on `stream` (a cudaStream_t passed as void*).  Return 0 on success, a negative lp_status on failure
 * (never throws); lp_last_error() gives the message of the calling thread's last failure.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with LP_ERR_CUDA.
 */
#ifndef LEOPARD_B200_H
#define LEOPARD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LP_ABI_VERSION 3

enum lp_status { LP_OK = 0, LP_ERR_ARG = -1, LP_ERR_CUDA = -2 };

/* GEMM epilogue activations */
enum lp_act {
  LP_ACT_NONE = 0,
  LP_ACT_GELU_TANH = 1, /* "gelu_pytorch_tanh": SigLIP MLP (MP/model/llava/clip_encoder.py:331), Megatron projector (D8) */
  LP_ACT_GELU_ERF = 2,  /* exact GELU: HF projector (EV/llava_multiimg_siglip_anyres.py:184) */
  LP_ACT_SWIGLU = 3     /* silu(gate)*up: MP/model/llava/transformer.py:129-139 */
};

const char* lp_last_error(void);
int lp_abi_version(void);

/* C[M,N] = act(A[M,K] · W[N,K]^T + bias) + residual        (tcgen05 / TMEM / TMA GEMM, fp32 accumulate)
 *   replaces torch.nn.functional.linear / Column-/RowParallelLinear / TELayerNormColumnParallelLinear GEMMs:
 *   EV/llava_multiimg_siglip_anyres.py:187-192 (projector), MP/model/llava/transformer.py:163-175,577-624 (decoder),
 *   MP/model/llava/clip_encoder.py:51-72 (ViT), MP/model/llava/vlm_model.py:32-60 (lm_head).
 * A: bf16 [M,K] lda; W: bf16 [N,K] ldw (nn.Linear layout); C: bf16 or fp32 (out_fp32) [M,N'] ldc.
 * bias: fp32 [N] or NULL.  residual: bf16 [*,N'] ldr or NULL, row index = res_row_mod ? row % res_row_mod : row
 *   (res_row_mod = 676 adds the ViT position embedding, idefics_vision_tower.py:147-150).  residual may alias C.
 * act == LP_ACT_SWIGLU: W rows are interleaved in 32-row granules [gate_0..31, up_0..31, gate_32..63, ...] and the
 *   output has N' = N/2 columns; otherwise N' = N.
 * block_n: 128, 160, 192, 224, 256, or 0 = choose (wave-quantisation heuristic).  Requires N % 8 == 0, lda/ldw % 8 == 0, 16-byte aligned bases. */
int lp_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N,
                 int K, const float* bias, const void* residual, long long ldr, int res_row_mod, int act, int out_fp32,
                 int block_n, void* stream);

/* Extended GEMM: the same kernel with three epilogue fusions that remove the standalone HBM-bound glue kernels of the layer loops
 * (SURVEY.md §2.3 K2/K5 "LN prologue + GEMM", K9 "RMSNorm fused into the next GEMM", K10/K11 "RoPE in the QKV epilogue";
 * reference: TELayerNormColumnParallelLinear MP/model/llava/clip_encoder.py:51-72, RMSNorm MLM/legacy/model/rms_norm.py:26-31,
 * RoPE MLM/core/models/common/embeddings/rotary_pos_embedding.py:197-239).
 *   norm_mode != NONE (consumer): C = act(norm(A) W^T + b).  The norm's gain is folded into W offline (W' = W diag(gamma), what
 *     `W` must point at), so norm(A) W^T = rs[row] * (A W'^T - mu[row] * colsum[col]), with mu / rs from `stats_in`: per row
 *     `stats_in_parts` float2 partials (sum, sum of squares over the K features) written by the producing GEMM (or lp_row_stats),
 *     summed in index order.  LAYER mode: colsum[n] = sum_k W'[n,k] (fp32) and bias must already hold beta W^T + b.
 *   stats_out (producer): float2 [M, stats_out_parts], stats_out_parts == 2 * ceil(N / block_n) (lp_gemm_pick_block_n gives the
 *     block_n a block_n = 0 call will use): (sum, sum of squares) of the bf16-rounded outputs per (column block, epilogue half).
 *   rope_heads > 0: columns are 128-wide heads; heads [0, rope_heads) get rotate-half RoPE with cos / sin tables fp32 [M, 64]
 *     (lp_rope_table) applied to the fp32 accumulators before the single bf16 rounding.  bf16 output, no bias / act / residual.
 *   a_t / b_t (transposed operands, the SFT step's dgrad and wgrad; autograd's `grad_output @ weight` and `grad_output.t() @ input`
 *     of torch.nn.functional.linear): a_t: A is given as At [K, M] row-major (lda = its row stride); b_t: W is given as Wt [K, N]
 *     row-major (ldw).  The tiles are fed to the tensor core as MN-major operands — no transposed copy is made.  Needs the CTA-pair
 *     mode (M > 128 and at least one 256 x 256 tile per SM pair), M % 8 == 0 when a_t; plain epilogue (bias / out_fp32 only). */
enum lp_norm_mode { LP_NORM_NONE = 0, LP_NORM_RMS = 1, LP_NORM_LAYER = 2 };
typedef struct lp_gemm_desc {
  const void* A; long long lda;
  const void* W; long long ldw;
  void* C; long long ldc;
  int M, N, K;
  const float* bias;
  const void* residual; long long ldr; int res_row_mod;
  int act, out_fp32, block_n;
  int norm_mode; const float* stats_in; int stats_in_parts; const float* colsum; float norm_eps;
  float* stats_out; int stats_out_parts;
  const float* rope_cos; const float* rope_sin; int rope_heads;
  int a_t, b_t;
} lp_gemm_desc;
int lp_gemm_bf16_ex(const lp_gemm_desc* desc, void* stream);
/* the tile width a block_n = 0 launch of this problem uses under the current scheduling mode */
int lp_gemm_pick_block_n(int M, int N, int act);
/* (sum, sum of squares) per row of x bf16 [rows, width] -> stats float2 [rows] (one partial per row): the producer of
 * `stats_in` when the rows were not written by a GEMM (merged embeddings entering decoder layer 0). */
int lp_row_stats(const void* x, long long ldx, int rows, int width, float* stats, void* stream);
/* layer loops (lp_vit_forward / lp_vit_layers / lp_decoder_prefill): 1 (default) = norms and RoPE fused into the GEMM epilogues
 * when the weight structs carry the folded matrices, 0 = standalone norm / RoPE kernels.  Returns the previous setting. */
int lp_model_set_fused_norm(int enable);

/* Tensor parallel + sequence parallel launches (reference: ColumnParallelLinear / RowParallelLinear with
 * sequence_parallel, MLM/core/tensor_parallel/layers.py:387-454, mappings.py:119-199).  The M rows are tp_size segments of
 * seg_rows rows, segment r owned by rank r.  One kernel does the collective AND the GEMM over NVLink peer memory:
 *   LP_TP_ALLGATHER       C[M, N'] = act(A W^T): A is this rank's [M, lda] gather buffer of which only its own segment is valid
 *                         on entry; peer_A[r] is rank r's buffer (peer pointer).  The epilogue warps of all CTAs pull the peers'
 *                         segments in while the tiles of the own segment are multiplied; tiles of a peer segment wait for the copy.
 *   LP_TP_REDUCE_SCATTER  partial sums A W^T (K is this rank's shard) leave the epilogue as TMA stores into the OWNER's staging
 *                         buffer peer_stage[r] = rank r's bf16 [tp_size, seg_rows, ldc] at slot [tp_rank]; the last CTA posts
 *                         `epoch` to every owner; lp_tp_reduce on the owner sums the slots in rank order (+ residual).  C unused.
 * Cross-rank signalling: 16 u32 words per rank (signal_local; signal_peer[r] = rank r's words as peer pointers): word [r] = rank
 * r's all-gather readiness, word [8 + r] = rank r's reduce-scatter completion, compared with `epoch` (monotonic call number, the
 * same on every rank).  counter: a zeroed local u32 for this call; recycle: an older call's counter to re-zero (or NULL).
 * Waits are bounded (a missing peer traps the kernel instead of hanging the GPU). */
enum lp_tp_mode { LP_TP_ALLGATHER = 1, LP_TP_REDUCE_SCATTER = 2 };
typedef struct lp_tp_desc {
  int tp_size, tp_rank, seg_rows, mode;
  unsigned epoch;
  const void* peer_A[8];
  void* peer_stage[8];
  unsigned* counter;
  unsigned* recycle;
  unsigned* signal_local;
  unsigned* signal_peer[8];
} lp_tp_desc;
int lp_gemm_bf16_tp(const lp_gemm_desc* desc, const lp_tp_desc* tp, void* stream);
/* out[seg_rows, cols] = residual + sum_r stage[r] (fp32 sum in rank order, one bf16 rounding) once every rank's word [8 + r] of
 * signal_local has reached `epoch`.  stage: this rank's bf16 [tp_size, seg_rows, ld]. */
int lp_tp_reduce(const lp_tp_desc* tp, const void* stage, long long ld, const void* residual, long long ldr, void* out, long long ldo,
                 int cols, void* stream);

/* GEMM scheduling mode: 0 = one CTA per 128-row tile; 1 = CTA pairs (thread-block cluster of 2, tcgen05
 * cta_group::2, 256-row tiles) for problems with at least one tile per SM pair.  Returns the previous mode.
 * Results are identical up to fp32 summation order (same K order: bit-identical in practice).
 * 2 = 4-CTA clusters: two pairs on column-adjacent tiles share the A tile by TMA multicast (-25 % L2 reads per tile);
 *     opt-in, verified under the CPU emulator only so far. */
int lp_gemm_set_pair_mode(int enable);

/* Stream-K tail of lp_gemm_bf16 / lp_gemm_bf16_ex (large problems whose last wave of tiles would leave most SMs idle, e.g. the
 * decoder's N = 4096 projections at S = 4796: 304 tiles on 74 CTA pairs): the k loops of the last 1 + fraction waves of tiles are
 * cut into equal ranges, one per CTA pair; a tile shared by two pairs is started by one (fp32 accumulator dumped to the workspace)
 * and finished by the next (accumulator loaded back into TMEM, k loop continued) — the summation order of every output element is
 * that of the unsplit kernel, results are bit-identical with it on or off.  The library never allocates: the caller provides
 * `ws` = ZERO-FILLED device memory of lp_gemm_streamk_workspace_bytes() bytes, 256-byte aligned, used by one stream at a time and
 * valid until replaced (NULL = off, the default).  lp_gemm_set_streamk(mode) switches it without dropping the workspace and
 * returns the previous setting: 0 = off, 1 (default) = the tail only, 2 = additionally PURE stream-K (every CTA pair owns one
 * contiguous range of the (row block, column block, k) space) for norm-consuming launches whose operands are L2-resident — the
 * ViT's qkv / fc1 GEMMs: balanced at any crop count, the fused LayerNorm's row statistics fetched once per pair; bit-identical,
 * but measured slightly slower than mode 1 on the B200, hence opt-in. */
size_t lp_gemm_streamk_workspace_bytes(void);
int lp_gemm_set_streamk_workspace(void* ws, size_t bytes);
int lp_gemm_set_streamk(int enable);
/* Residual launches (C = A W^T + bias + residual, bf16, the o / down / out-proj / fc2 GEMMs of the layer loops) in CTA-pair mode:
 * 1 (default) = the residual tile is fetched by TMA into a per-warp ring one 32-column chunk ahead, added in place and stored
 * from the same buffer; 0 = the earlier register path (coalesced loads handed through the staging tile).  Same arithmetic:
 * results are bit-identical.  Returns the previous setting. */
int lp_gemm_set_tma_residual(int enable);
/* Tile raster of launches whose W fits L2 while A does not and whose last column block is ragged (the ViT fc2: N = 1152 = 4 x 256 +
 * 128): 1 (default) = row-major tile order, so that every persistent unit gets each column block — and so the cheap ragged ones —
 * equally often; 0 = the earlier groups of ~SMs / column-blocks row blocks (35 of 74 CTA pairs got all the half-width tiles).
 * Scheduling only: results are bit-identical.  Returns the previous setting. */
int lp_gemm_set_raster_balance(int enable);
/* Epilogue warps of lp_gemm_bf16(_ex) wait for a whole k loop before their tile is ready; `ns` > 0 makes them sleep that many
 * nanoseconds between probes of the accumulator-full barrier instead of polling back to back (0 = poll).  Timing only: results
 * are unaffected.  Returns the previous setting. */
int lp_gemm_set_epilogue_sleep(int ns);

/* Launch the hot prefill kernels (GEMM, attention v3, norms, RoPE) with programmatic dependent launch: the next kernel's
 * CTAs are scheduled as SMs free up and run their prologue (barrier init, TMEM allocation, descriptor prefetch) while the
 * previous kernel drains; each kernel executes griddepcontrol.wait before its first dependent global access.  Results
 * are unchanged.  Returns the previous setting (default set by the library). */
int lp_set_prefill_pdl(int enable);

/* O[b,t,h,:] = softmax(scale * Q K^T (+causal)) V          (tcgen05 flash attention, fp32 softmax, bf16 in/out)
 *   replaces flash_attn_varlen_func (MP/model/llava/transformer.py:476-515,869-877; the 4x KV repeat_interleave at
 *   :829-836 is not materialised) and the ViT's TEDotProductAttention / HF SiglipAttention
 *   (MP/model/llava/clip_encoder.py:58-63).
 * qkv: bf16 [B*T, ld]; for head h: q cols q_off + h*head_stride, k cols k_off + (h / (n_heads/n_kv_heads))*head_stride,
 *   v likewise from v_off.  Supported (head_dim, head_stride): (128,128) and (72,80) — the latter expects the 8 pad
 *   columns of q and k to be exact zeros (the ViT qkv weight is stored zero-padded).
 * out: bf16 [B*T, ldo], head h at cols h*head_dim.  scale = head_dim^-0.5 in the reference.
 * variant bit 0: keep P in TMEM (tcgen05.mma A-from-TMEM) instead of shared memory (one-tile kernel);
 * variant bit 1: two 128-row query tiles per CTA ping-ponged on the tensor pipe (P always in TMEM);
 * variant bits 2-3 (with bit 1 or 6): 0/1/2 of every 4 softmax exponentials evaluated on the FMA pipe (polynomial
 *   2^x, rel. error 1.1e-4) instead of MUFU.EX2;
 * variant bit 6 (64): v3 kernel (lp_attn3.cu) — two query tiles, S in 64-key halves, one MMA issuer per tile; the
 *   softmax row sum is taken in fp32 before P is rounded to bf16 (as flash-attention does);
 * variant bits 4 / 5 (16 / 32): force the softmax turn protocol on / off (see lp_attention_set_pingpong);
 * variant bit 7 (128): v4 kernel (lp_attn4.cu) — v3's dataflow, but ONE softmax warpgroup whose thread r owns row r of both
 *   query tiles and runs their exponentials as two independent instruction streams (no turn protocol);
 * variant bit 8 (256): short non-causal (72, 80) problems whose query-tile pairs are all complete (the SigLIP tower: T = 676) run
 *   the persistent form of v3 (lp_attn5.cu): one CTA per SM walks (batch, head, pair) items, barriers / TMEM set up once, the next
 *   item's loads and first QK^T overlap the epilogue; other shapes ignore the bit.  Same arithmetic as v3: bit-identical. */
int lp_attention_fwd(const void* qkv, long long ld, int B, int T, int n_heads, int n_kv_heads, int head_dim,
                     int head_stride, int q_off, int k_off, int v_off, int causal, float scale, void* out,
                     long long ldo, int variant, void* stream);

/* Softmax turn protocol of the two-tile kernel (variant bit 1): the two softmax warpgroups of a CTA alternate strictly
 * (named barriers), so one tile's exponentials overlap the other tile's MMAs instead of both tiles sharing the MUFU
 * unit and then both waiting for the tensor pipe.  Same arithmetic: results are bit-identical with it on or off.
 * lp_attention_set_pingpong sets the default (returns the previous one); variant bit 4 (16) forces it on and bit 5 (32)
 * forces it off for one call. */
int lp_attention_set_pingpong(int enable);

/* Attention kernel the model-level entry points (lp_vit_forward, lp_vit_layers, lp_decoder_prefill) pass to
 * lp_attention_fwd as `variant`; returns the previous value.  Default 336 = variant bit 6 (64): the v3 kernel — two
 * query tiles per CTA, S delivered in 64-key halves with their own TMEM columns so QK^T of the next half never waits
 * for the softmax of the current one, one MMA-issuing thread per query tile — plus bit 4 (16): the turn protocol on
 * its exp sections, plus bit 8 (256): the persistent form for the vision tower's short sequences.  80 = v3 everywhere,
 * 2 selects the v2 two-tile kernel. */
int lp_model_set_attention_variant(int variant);

/* Cross-attention with separate query and key/value buffers, non-causal (Idefics2 perceiver resampler: 64 latent
 * queries attend to concat(context, latents); MP/model/idefics2/perceiver_transformer.py:582-706, flash-varlen call
 * :347-420; HF Idefics2PerceiverAttention).  q: bf16 [B*Tq, ldq], head h at cols h*head_dim; kv: bf16 [B*Tk, ldkv],
 * k of kv-head g at cols k_off + g*head_dim, v at v_off + g*head_dim.  head_dim 96 or 128, Tq <= 128 per tile row
 * block (any Tq works; rows are tiled by 128).  out: bf16 [B*Tq, ldo]. */
int lp_attention_cross(const void* q, long long ldq, const void* kv, long long ldkv, int B, int Tq, int Tk, int n_heads,
                       int n_kv_heads, int head_dim, int k_off, int v_off, float scale, void* out, long long ldo,
                       void* stream);

/* out = LayerNorm(x) * gamma + beta, eps inside rsqrt; one row per CTA, fp32 statistics (two-pass).
 *   SigLIP layer_norm1/2 and post_layernorm / ln_post (MP/model/idefics2/idefics_vision_tower.py:77-81,176).
 * shuffle_side = 26 additionally applies Leopard's pixel_shuffle (EV/llava_multiimg_siglip_anyres.py:165-176 ==
 *   MP/model/llava/vlm_model.py:456-466) through the store address: input row (b, h, w) of a side x side grid goes to
 *   output row b*(side/2)^2 + (h/2)*(side/2) + (w/2), columns ((h%2)*2 + (w%2))*D .. +D, so `out` is [M/4, 4*D].
 * x/out bf16, gamma/beta fp32 [D]; D % 8 == 0, D <= 4096. */
int lp_layernorm(const void* x, long long ldx, const float* gamma, const float* beta, void* out, long long ldo, int M,
                 int D, float eps, int shuffle_side, void* stream);

/* out = x * rsqrt(mean(x^2) + eps) * gamma   (MLM/legacy/model/rms_norm.py:26-31; HF LlamaRMSNorm) */
int lp_rmsnorm(const void* x, long long ldx, const float* gamma, void* out, long long ldo, int M, int D, float eps,
               void* stream);

/* RMSNorm kernel for rows wider than 1280: 1 = packed register-resident (row read once, 16 loads per lane in flight;
 * default), 0 = two-pass re-read.  Bit-identical results.  Returns the previous value. */
int lp_norm_set_variant(int variant);

/* cos/sin[s, j] = cos/sin(pos[s] * inv_freq[j]), fp32 [S, half]
 *   (MLM/core/models/common/embeddings/rotary_pos_embedding.py:125-157; inv_freq incl. llama3.1 rescale :48-83 is
 *   computed by the host shim) */
int lp_rope_table(const int* pos, const float* inv_freq, float* cos_out, float* sin_out, int S, int half, void* stream);

/* in-place rotate-half RoPE on n_heads heads of width head_dim starting at column col0 of x[S, ld] (bf16):
 *   t*cos + cat(-t2, t1)*sin  (rotary_pos_embedding.py:197-239) */
int lp_rope_inplace(void* x, long long ld, int S, int col0, int n_heads, int head_dim, const float* cos_t,
                    const float* sin_t, void* stream);

/* Token embedding + image-feature merge for one sequence (B = 1).
 *   Every id == image_id is an image slot that receives rows_per_slot consecutive rows of feats (slots consume feature
 *   rows in order); every other id receives table[clamp(id, 0, vocab-1)].
 *   rows_per_slot = 169, image_id = 128200: HF expand convention (_merge_input_ids_with_image_features as called at
 *     EV/llava_multiimg_siglip_anyres.py:285-287; SURVEY.md App. A.1, B = 1 / no padding) -> S = L + 168 * slots.
 *   rows_per_slot = 1, image_id = -200: Megatron scatter convention (MP/model/llava/vlm_model.py:516-533).
 * plan: device int32 scratch [2*L + 2]; on completion plan[2*L] = number of slots, plan[2*L+1] = merged length S
 *   (the host shim reads them back to raise the reference's count-mismatch error). */
int lp_embed_merge(const long long* ids, int L, long long image_id, int rows_per_slot, const void* table, long long ldt,
                   long long vocab, const void* feats, long long ldf, int n_feat_rows, void* out, long long ldo,
                   int max_out_rows, int D, int* plan, void* stream);

/* out[(n*G+py)*G+px, c*196 + ky*14 + kx] = pixels[n, c, py*14+ky, px*14+kx] (bf16), columns [C*196, ldo) zeroed:
 *   A operand of the patch-embedding GEMM (conv1 k=14 s=14, MP/model/idefics2/idefics_vision_tower.py:57-64,118-122).
 *   pixels: [N, C, H, W] fp32 (HF eval, EV/...anyres.py:403-405) or bf16 (Megatron). */
int lp_im2col_patch14(const void* pixels, int pixels_fp32, void* out, long long ldo, int N, int C, int H, int W,
                      void* stream);

/* loss[r] = logsumexp(logits[r, :V]) - logits[r, labels[r]] in fp32 (0 where the label is negative): the per-token
 * cross entropy GPTModel.forward returns when `labels` is given (post_language_model_processing,
 * MP/model/llava/gpt_model.py:27-51 -> vocab_parallel_cross_entropy on `.float()` logits; TP = 1).  The caller shifts
 * labels and applies the loss mask / mean as loss_func does (EX/llava/pretrain_megatron_llava.py:101-121). */
int lp_cross_entropy(const void* logits, int logits_fp32, long long ld, const long long* labels, float* loss, int rows,
                     int V, void* stream);

/* Standalone pixel_shuffle(x, scale_factor=2) (EV/llava_multiimg_siglip_anyres.py:165-176): x bf16 [M, D] with
 * M = B*side*side rows -> out [M/4, 4*D].  (The fused path never materialises x: lp_layernorm's shuffle store.) */
int lp_pixel_shuffle(const void* x, long long ldx, void* out, long long ldo, int M, int D, int side, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Model-level entry points (the layer loops).  Weight structs hold DEVICE pointers in the kernel-native layout that
 * leopard_b200/native_weights.py produces from an HF- or Megatron-layout state dict:
 *   bf16 matrices in nn.Linear layout [out, in]; biases / norm gains as fp32 vectors (bf16-rounded values).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct lp_vit_dims {
  int layers, hidden, ffn, heads, head_dim, head_stride; /* 27, 1152, 4304, 16, 72, 80 */
  int grid, tokens, patch_k, patch_k_pad;                /* 26, 676, 588, 592 */
  float eps;                                             /* 1e-6 */
} lp_vit_dims;

typedef struct lp_vit_layer_weights {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const void* qkv_w;  /* [3*heads*head_stride, hidden]: per head 72 real rows + 8 zero rows; q block, k block, v block */
  const float* qkv_b; /* [3*heads*head_stride], zero in the pad rows */
  const void* out_w;  /* [hidden, hidden] */
  const float* out_b;
  const void* fc1_w;  /* [ffn, hidden] */
  const float* fc1_b;
  const void* fc2_w;  /* [hidden, ffn] */
  const float* fc2_b;
  /* optional (NULL = standalone LayerNorm kernels): LayerNorm folded into the consuming GEMM, see lp_gemm_bf16_ex.
   * qkv_wf = bf16(qkv_w * ln1_g[None, :]), qkv_c1[n] = sum_k qkv_wf[n, k], qkv_c2 = qkv_w ln1_b + qkv_b; fc1_* likewise with ln2 */
  const void* qkv_wf; const float* qkv_c1; const float* qkv_c2;
  const void* fc1_wf; const float* fc1_c1; const float* fc1_c2;
} lp_vit_layer_weights;

typedef struct lp_vit_weights {
  const void* patch_w;  /* [hidden, patch_k_pad]: conv1 weight flattened (c, ky, kx), zero padded */
  const float* patch_b; /* [hidden] */
  const void* pos;      /* bf16 [tokens, hidden] learned position embedding */
  const lp_vit_layer_weights* layers; /* HOST array of `layers` structs */
  const float *post_g, *post_b;
} lp_vit_weights;

typedef struct lp_proj_weights {
  const void* w1; /* [hidden_dec, 4*hidden_vit] */
  const float* b1;
  const void* w2; /* [hidden_dec, hidden_dec] */
  const float* b2;
  int act;        /* LP_ACT_GELU_ERF (HF eval) or LP_ACT_GELU_TANH (Megatron fused path) */
} lp_proj_weights;

typedef struct lp_dec_dims {
  int layers, hidden, ffn, heads, kv_heads, head_dim; /* 32, 4096, 14336, 32, 8, 128 */
  int vocab;                                          /* rows of lm_head (multiple of 8) */
  float eps;                                          /* 1e-5 */
} lp_dec_dims;

typedef struct lp_dec_layer_weights {
  const float* ln1_g;
  const void* qkv_w;    /* [(heads + 2*kv_heads)*head_dim, hidden]: all q heads, then k heads, then v heads */
  const void* o_w;      /* [hidden, hidden] */
  const float* ln2_g;
  const void* gateup_w; /* [2*ffn, hidden], 32-row granules interleaved gate/up (see lp_gemm_bf16, LP_ACT_SWIGLU) */
  const void* down_w;   /* [hidden, ffn] */
  /* optional (NULL = standalone RMSNorm kernels): qkv_wf = bf16(qkv_w * ln1_g[None, :]), gateup_wf = bf16(gateup_w * ln2_g[None, :]) */
  const void* qkv_wf;
  const void* gateup_wf;
} lp_dec_layer_weights;

typedef struct lp_dec_weights {
  const lp_dec_layer_weights* layers; /* HOST array */
  const float* final_g;
  const void* lm_head;   /* [vocab, hidden] */
  const float* inv_freq; /* device fp32 [head_dim/2] (llama3.1-rescaled where configured) */
} lp_dec_weights;

enum lp_logits_mode { LP_LOGITS_NONE = 0, LP_LOGITS_LAST = 1, LP_LOGITS_ALL = 2 };

/* SigLIP tower over N crops: pixels [N,3,364,364] (fp32 or bf16) -> ln_post output.
 *   shuffle = 0: out is last_hidden_state [N*676, 1152] (what `self.vision_tower(pixel_values).last_hidden_state`
 *     returns, EV/llava_multiimg_siglip_anyres.py:268,273; MP/model/idefics2/idefics_vision_tower.py:104-178);
 *   shuffle = 1: out is pixel_shuffle(last_hidden_state) [N*169, 4608] (EV/...anyres.py:165-176), written directly. */
size_t lp_vit_workspace_bytes(const lp_vit_dims* dims, int N);
int lp_vit_forward(const lp_vit_dims* dims, const lp_vit_weights* w, const void* pixels, int pixels_fp32, int N,
                   void* out, long long ldo, int shuffle, void* workspace, size_t ws_bytes, void* stream);

/* Encoder layers + ln_post on an ALREADY EMBEDDED residual stream x [R, hidden] (overwritten) that packs n_seqs images
 * of different lengths (rows cu_seqlens_host[s] .. cu_seqlens_host[s+1]-1): the NaViT / Idefics2 tower
 * (MP/model/idefics2/idefics_vision_tower.py:152-178).  Padded patches are removed by the caller instead of masked —
 * the result on valid patches is identical to the reference's key-padding mask.  out: bf16 [R, hidden]. */
size_t lp_vit_layers_workspace_bytes(const lp_vit_dims* dims, int total_rows);
int lp_vit_layers(const lp_vit_dims* dims, const lp_vit_weights* w, void* x, const int* cu_seqlens_host, int n_seqs,
                  void* out, long long ldo, void* workspace, size_t ws_bytes, void* stream);

/* linear_2(act(linear_1(x))) on shuffled rows [rows, in_dim] -> [rows, hidden]
 *   (myLlavaMultiModalProjector.forward EV/...anyres.py:187-192; Megatron MultimodalProjector
 *    MLM/core/models/vision/multimodal_projector.py:50-57 via MP/model/llava/vlm_model.py:468-475) */
size_t lp_project_workspace_bytes(int rows, int hidden);
int lp_project(const lp_proj_weights* w, const void* shuffled, int rows, int in_dim, int hidden, void* out,
               long long ldo, void* workspace, size_t ws_bytes, void* stream);
/* lp_project as ONE kernel (default, when the GEMM flag workspace of lp_gemm_set_streamk_workspace is present and rows > 128): the
 * tiles of both layers share one persistent launch, a layer-2 tile waits (dataflow counters, no launch boundary) until the layer-1
 * tiles of its row block have landed in the L2-resident intermediate; the pixel shuffle is ln_post's store, so shuffle + projector
 * cost this single launch.  0 = two GEMM launches.  Same arithmetic: results are bit-identical.  Returns the previous setting. */
int lp_model_set_chained_projector(int enable);

/* Decoder prefill over n_seqs packed sequences (rows cu_seqlens_host[s] .. cu_seqlens_host[s+1]-1 of x).
 *   x: bf16 [R, hidden] merged embeddings; used as the residual stream and overwritten with the last layer's output.
 *   pos_ids: device int32 [R].  kv_cache: NULL or bf16 [layers][2][kv_capacity][kv_heads*head_dim] (K after RoPE, then V; rows
 *     0..R-1 are written, kv_capacity >= R leaves room for decode steps).
 *   logits: LP_LOGITS_LAST -> [n_seqs, vocab] (final norm + lm_head on each sequence's last row only);
 *           LP_LOGITS_ALL  -> [R, vocab] (what the reference materialises, EV/...anyres.py:333); fp32 or bf16.
 *   Replaces HF LlamaForCausalLM.forward (EV/...anyres.py:322-333) / Megatron ParallelTransformer.forward +
 *   parallel_lm_logits (MP/model/llava/transformer.py:1747-1860, MP/model/llava/vlm_model.py:32-60). */
size_t lp_decoder_workspace_bytes(const lp_dec_dims* dims, int total_rows, int n_seqs);
int lp_decoder_prefill(const lp_dec_dims* dims, const lp_dec_weights* w, void* x, const int* pos_ids,
                       const int* cu_seqlens_host, int n_seqs, void* kv_cache, long long kv_capacity, void* logits,
                       long long ld_logits, int logits_mode, int logits_fp32, void* workspace, size_t ws_bytes,
                       void* stream);

/* One greedy-decode step after a prefill (cached branch of myLlavaForConditionalGeneration.forward,
 * EV/llava_multiimg_siglip_anyres.py:291-320; B = 1, no padding => position = number of cached rows, App. A.2).
 *   token_id: DEVICE int64[1] (the previously sampled token); its embedding row is gathered from embed_table.
 *   kv_cache: the prefill's cache, row `position` of every layer is written, rows 0..position are attended to.
 *   logits: fp32 [vocab] of the new position. */
size_t lp_decode_workspace_bytes(const lp_dec_dims* dims);
int lp_decode_step(const lp_dec_dims* dims, const lp_dec_weights* w, const void* embed_table, long long ld_embed,
                   const long long* token_id, int position, void* kv_cache, long long kv_capacity, void* logits,
                   void* workspace, size_t ws_bytes, void* stream);

/* The same step with the greedy loop state resident on the DEVICE, so that one step can be captured into a CUDA graph
 * and replayed per token without host work (GenerationMixin's greedy search driven by `generate(...,
 * max_new_tokens=128, use_cache=True)`, EV/llava_multiimg_siglip_anyres.py:448-452):
 *   token_io:    DEVICE int64[1]; in = previously selected token, out = argmax(logits[0:vocab_valid]) (lowest index
 *                among ties, as torch.argmax);
 *   position_io: DEVICE int32[1]; in = number of cached rows (row written by this step), out = in + 1.  The caller
 *                guarantees position < kv_capacity for every replay (out-of-range positions write nothing).
 *   logits: fp32 [vocab] of the new position (kept for callers that sample differently). */
int lp_decode_greedy_step(const lp_dec_dims* dims, const lp_dec_weights* w, const void* embed_table, long long ld_embed,
                          long long* token_io, int* position_io, void* kv_cache, long long kv_capacity, void* logits,
                          int vocab_valid, void* workspace, size_t ws_bytes, void* stream);

/* The greedy step above captured once into a CUDA graph (on a private stream) and replayed per token: one
 * cudaGraphLaunch instead of ~200 kernel launches.  All pointers are baked into the graph and must stay valid and
 * unchanged until lp_decode_graph_destroy; token_io / position_io are read and advanced on the device by every replay.
 * The handle owns only the cudaGraphExec_t. */
typedef struct lp_decode_graph lp_decode_graph;
int lp_decode_graph_create(const lp_dec_dims* dims, const lp_dec_weights* w, const void* embed_table,
                           long long ld_embed, long long* token_io, int* position_io, void* kv_cache,
                           long long kv_capacity, void* logits, int vocab_valid, void* workspace, size_t ws_bytes,
                           lp_decode_graph** out);
int lp_decode_graph_launch(lp_decode_graph* graph, void* stream);
void lp_decode_graph_destroy(lp_decode_graph* graph);

/* Decode implementation switch, returns the previous mode (default 7).
 *   bit 0: HBM-streaming GEMV kernels with fused RMSNorm / RoPE + cache append / SwiGLU / residual; 0 = M = 1 rows
 *          through lp_gemm_bf16 (first implementation, lp_decode_step only);
 *   bit 1: launch the step's kernels with programmatic dependent launch (griddepcontrol): the next kernel's CTAs request
 *          their first weight rows while the previous kernel drains;
 *   bit 2: block-cooperative GEMV (a CTA reduces 8 rows with K split over its threads, activations in registers)
 *          instead of warp-per-row-pair;
 *   bit 4: lp_decode_greedy_step runs the whole step as ONE persistent cooperative kernel with grid barriers (21). */
int lp_decode_set_mode(int mode);
/* MiB of the NEXT projections' weights (o_proj, head of gate/up) that the decode attention kernel pulls into L2 while it runs —
 * attention only reads the KV cache, so the HBM is nearly idle for its duration.  Results unchanged.  Defaults 32 / 24 (measured:
 * 3.14 ms per token against 3.18-3.25 without, inside the +-0.1 ms noise of the measurement); 0 / 0 = off. */
int lp_decode_set_prefetch(int o_proj_mb, int gateup_mb);

/* ------------------------------------------------------------------------------------------------------------------
 * GPU image front-end (SURVEY.md §8 a3 + a5, "next" row f-3): one RGB uint8 image -> the ViT inputs it contributes.
 *   rgb: DEVICE uint8 [H, W, 3] (what np.asarray(PIL.Image.convert('RGB')) holds), 16-byte aligned base.
 *   out_global: [3, 364, 364] — SiglipImageProcessor.preprocess(image) (EV/llava_multiimg_siglip_anyres.py:403-404):
 *     bicubic resize of the whole image to 364 x 364, x * rescale (float64, rounded to float32), (x - mean) / std.
 *   out_tiles (target_w, target_h > 0, both multiples of 364; NULL otherwise): [(target_h/364)*(target_w/364), 3, 364,
 *     364] — resize_and_pad_image(image, (target_w, target_h)) (EV/...anyres.py:102-140: aspect-preserving bicubic
 *     resize pasted centred on black), divide_to_patches (:143-162, row-major tiles), each tile through the same
 *     rescale / normalise (a 364 x 364 tile is not resized again).
 *   The resize is Pillow's ImagingResample (bicubic, a = -0.5, horizontal pass first, uint8 intermediate, 22-bit fixed
 *   point taps; vertical pass first for images more than 100x taller than wide that shrink vertically, as Pillow >= 11
 *   does): results are bit-identical to Pillow 12.2 + the PIL-backed SiglipImageProcessor.  Outputs are fp32
 *   (out_fp32 = 1, the reference's pixel_values dtype) or bf16 (what the patch-embedding GEMM consumes).
 *   rescale = 1/255, mean = std = {0.5, 0.5, 0.5} for SigLIP (HOST pointers).
 *   The target resolution comes from the host planner (allocate_patches / select_best_resolution, EV :26-99). */
size_t lp_image_frontend_workspace_bytes(int H, int W, int target_w, int target_h);
int lp_image_frontend(const void* rgb, int H, int W, int target_w, int target_h, void* out_global, void* out_tiles,
                      int out_fp32, double rescale, const float* mean, const float* std_, void* workspace,
                      size_t ws_bytes, void* stream);

/* Idefics2ImageProcessor.preprocess for ONE image on the GPU (the processor evaluations/models/idefics2_multiimg.py:22-25 builds:
 * size = {'longest_edge': R, 'shortest_edge': 0}, do_image_splitting = False; transformers' Idefics2ImageProcessor, PIL backend):
 * `Image.resize((out_w, out_h), resample)` with Pillow's ImagingResample arithmetic (filter 0 = BICUBIC, 1 = BILINEAR — the
 * processor's default), x * rescale in float64 rounded to float32, (x - mean) / std in float32, pasted at the top-left corner of a
 * planar [3, canvas_h, canvas_w] canvas (fp32 or bf16) whose other pixels are `pad_value` (the processor pads with 0.0 AFTER
 * normalising); mask (optional): int64 [canvas_h, canvas_w] = pixel_attention_mask.  (out_w, out_h) comes from the processor's own
 * integer rule (leopard_b200.frontend.idefics2_resize_size).  Bit-identical to the PIL-backed processor. */
size_t lp_image_resize_workspace_bytes(int H, int W, int out_w, int out_h, int filter);
int lp_image_resize(const void* rgb, int H, int W, int out_w, int out_h, int filter, int canvas_w, int canvas_h, float pad_value,
                    void* out, int out_fp32, long long* mask, double rescale, const float* mean, const float* std,
                    void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * SFT step (FIRST IMPLEMENTATION, specification = oracle/train.py): the HBM-bound half of the backward
 * pass, the masked LM loss and the optimizer.  In the reference all of this is torch autograd + apex inside Megatron's
 * train_step (Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py:101-160,
 * Megatron-LM-240424/megatron/core/optimizer/optimizer.py:300-370, clip_grads.py:103-141).  The dgrad / wgrad
 * contractions run on lp_gemm_bf16 over operands re-laid out by lp_transpose_bf16 (it contracts the contiguous
 * dimension): dX = gemm(dY, W^T), dW = gemm(dY^T, X^T) with fp32 output.
 *
 * lp_transpose_bf16: out[c, r] = in[r, c]; 4-byte aligned bases, even leading dimensions.
 * lp_rmsnorm_bwd:    dx = r (dy g) - x r^3 mean(dy g x) (+ add, the residual branch's gradient; may be NULL; dx may
 *                    alias add or dy), dgamma[D] = sum_rows dy x r (overwritten; fixed summation order).
 * lp_swiglu_fwd/bwd: m = silu(g) u on the [M, 2F] output of the fused gate/up weight (32-column granules, see
 *                    lp_gemm_bf16) and its gradient, written in the same interleaved layout.
 * lp_gelu_bwd:       dz = dy gelu'(z), act = LP_ACT_GELU_TANH or LP_ACT_GELU_ERF (projector, :187-192).
 * lp_colsum_bf16:    out[c] = sum_r dy[r, c] in fp32 (bias gradients).
 * lp_ce_fwd_bwd:     loss_func :101-121 on rows already shifted by the caller: loss[r] = lse - logits[r, target[r]],
 *                    dlogits[r] = (softmax - onehot) * weight[r] with weight = loss_mask / sum(loss_mask); weight-0 rows
 *                    cost one zero-fill; columns V..Vp-1 (vocabulary padding) receive 0.
 * lp_embed_bwd:      dE[id] += d[r] for ids >= 0; ids == image_id pass nothing (vlm_model.py:526-533 overwrites those
 *                    rows); other negative ids feed row 0 (the forward's clamp, vlm_model.py:226), ids >= vocab (rows of dE) feed row
 *                    vocab - 1 like the forward gather.  fp32 atomics.
 * lp_gather_rows_bf16: out[i] = in[idx[i]] — gradient of the image-feature scatter.
 * lp_softmax_bwd:    attention backward over materialised score rows: P = softmax(scale * scores, causal),
 *                    dS = scale * P * (dp - sum_j P dp); bf16 outputs, zeros beyond the causal limit.
 * lp_accumulate_f32:  dst += src (fp32) — gradient accumulation over micro-batches.
 * lp_sumsq_accum / lp_grad_norm_finish: global L2 norm over fp32 gradients (partial[] must be zeroed per step;
 *                    lp_grad_norm_workspace_bytes() bytes) -> out2 = {norm, clip coefficient} in device memory.
 * lp_adamw_step:     FusedAdam (adam_w_mode) on fp32 main params with the gradient scaled by out2[1]; also refreshes
 *                    the bf16 model copy (p_bf16 may be NULL).  step counts from 1.  A non-finite out2[0] (inf / NaN
 *                    gradients) makes the call a no-op: the iteration is skipped as in Megatron's found-inf path. */
/* Flash-style attention backward (tcgen05; (head_dim, head_stride) = (128, 128) causal or full, or (72, 80) full — the
 * vision tower's zero-padded head layout, whose padding columns of dqkv are written as zeros) — SECOND implementation,
 * compiled but not yet run on hardware; train.py selects it with LEOPARD_ATTN_BWD=flash.  Replaces autograd through
 * flash_attn_varlen_func (MP/model/llava/transformer.py:476-515,869-877).  qkv / dqkv: [B*T, ld] rows with the q heads at
 * q_off, k heads at k_off, v heads at v_off (head_stride columns each); out / dout: [B*T, ldo] forward output and its
 * gradient (head_dim columns per head).
 * dqkv receives dQ, dK, dV in the qkv column layout (dK / dV summed over the q heads of each kv group, scale applied).
 * Deterministic (no atomics).  The row log-sum-exp is recomputed inside (no forward change needed). */
size_t lp_attention_bwd_workspace_bytes(int B, int T, int n_heads);
int lp_attention_bwd(const void* qkv, long long ld, const void* out, const void* dout, long long ldo, int B, int T,
                     int n_heads, int n_kv_heads, int head_dim, int head_stride, int q_off, int k_off, int v_off, int causal,
                     float scale, void* dqkv, long long ldd, void* workspace, size_t ws_bytes, void* stream);
int lp_transpose_bf16(const void* in, long long ld_in, int rows, int cols, void* out, long long ld_out, void* stream);
/* Vision-tower backward pieces: LayerNorm backward (dx (+ add), dgamma, dbeta; fixed summation order), the inverse of the
 * pixel shuffle (gradient of lp_pixel_shuffle / lp_layernorm's shuffled store, EV :165-176), and the position-embedding
 * gradient out[t] = sum_n d[n * T + t] (the forward adds pos[row % T], idefics_vision_tower.py:147-150). */
size_t lp_layernorm_bwd_workspace_bytes(int D);
int lp_layernorm_bwd(const void* x, long long ldx, const float* gamma, const void* dy, long long lddy, const void* add,
                     long long ldadd, void* dx, long long lddx, float* dgamma, float* dbeta, int M, int D, float eps,
                     void* workspace, size_t workspace_bytes, void* stream);
int lp_pixel_unshuffle(const void* d, long long ldd, void* out, long long ldo, int M, int D, int side, void* stream);
int lp_rowmod_sum(const void* d, long long ldd, int n_rep, int T, int D, float* out, void* stream);
size_t lp_rmsnorm_bwd_workspace_bytes(int D);
int lp_rmsnorm_bwd(const void* x, long long ldx, const float* gamma, const void* dy, long long lddy, const void* add,
                   long long ldadd, void* dx, long long lddx, float* dgamma, int M, int D, float eps, void* workspace,
                   size_t workspace_bytes, void* stream);
int lp_swiglu_fwd(const void* gu, long long ldgu, void* out, long long ldo, int M, int F, void* stream);
int lp_swiglu_bwd(const void* gu, long long ldgu, const void* dm, long long lddm, void* dgu, long long lddgu, int M, int F,
                  void* stream);
int lp_gelu_bwd(const void* z, long long ldz, const void* dy, long long lddy, void* dz, long long lddz, int M, int N, int act,
                void* stream);
int lp_colsum_bf16(const void* dy, long long ld, int M, int N, float* out, void* stream);
int lp_ce_fwd_bwd(const void* logits, int logits_fp32, long long ld, const long long* target, const float* weight, float* loss,
                  void* dlogits, long long ldd, int rows, int V, int Vp, void* stream);
int lp_embed_bwd(const void* d, long long ldd, const long long* ids, int S, long long image_id, float* dE, long long ldE, int D,
                 long long vocab, void* stream);
int lp_gather_rows_bf16(const void* in, long long ld_in, const int* idx, int n, int D, void* out, long long ld_out,
                        void* stream);
int lp_softmax_bwd(const float* scores, long long lds, const float* dp, long long lddp, void* P, long long ldp, void* dS,
                   long long ldds, int Sq, int Sk, int causal, float scale, void* stream);
int lp_accumulate_f32(float* dst, const float* src, long long n, void* stream);
size_t lp_grad_norm_workspace_bytes(void);
int lp_sumsq_accum(const float* x, long long n, float* partial, void* stream);
int lp_grad_norm_finish(const float* partial, float max_norm, float* out2, void* stream);
int lp_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, const float* norm_and_clip, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Measurement hooks (used by bench.py only).
 * lp_launch_count: kernels this library has launched since load (the bench's "gpu_launches" claim).
 * lp_profile_enable(capacity > 0): from now on every lp_gemm_bf16 / lp_attention_fwd / norm / rope / merge launch is
 *   bracketed by a CUDA event pair on ITS stream, until `capacity` records are taken; capacity 0 disables.
 * lp_profile_read: synchronises on the recorded events, returns up to `capacity` records and disables profiling.
 *   kind: 1 gemm (a,b,c = M,N,K), 2 attention (a,b,c = B,T,heads*1000+head_dim), 3 norm (a,b = M,D), 4 rope, 5 merge,
 *   6 im2col, 7 pixel-shuffle, 8 decode attention, 10-13 decode GEMV (plain, qkv+rope, swiglu, logits),
 *   14 image front-end (a,b,c = H, W, tiles). */
typedef struct lp_profile_record { int kind, a, b, c; float ms; } lp_profile_record;
long long lp_launch_count(void);
/* Debug hook of the single-kernel decode step: when a DEVICE uint64 buffer (>= 2 * (5 * layers + 4) entries) is set,
 * CTA 0 records %globaltimer before and after every grid barrier (entries 2e, 2e+1 for barrier e = 1, 2, ...). */
int lp_decode_debug_stamps(void* device_u64_buffer);
int lp_profile_enable(int capacity);
int lp_profile_read(lp_profile_record* out, int capacity);

/* ------------------------------------------------------------------------------------------------------------------
 * PARITY MODE (leopard_b200/parity.py): the fp32-equivalent forward that answers BASELINE.json's "logits within 1e-3 max-abs of
 * the reference forward" — the reference's HF evaluation is fp32 end to end (EV/llava_multiimg_siglip_anyres.py:373).  Activations
 * stay fp32; every contraction runs on lp_gemm_bf16 with the fp32 operand split into three bf16 terms laid out along K
 * (lp_parity_split), fp32 accumulation in TMEM, fp32 output; the glue between the GEMMs are these fp32 kernels.
 *   lp_parity_split   x fp32 [M, K] -> bf16 [M, 3K] = [hi | mid | lo] (layout 0; pairs with a bf16 weight tiled 3x along K), or
 *                     [M, 6K] = [h h h m m l] (layout 1, A side) / [h m l h m h] (layout 2, B side) for fp32 x fp32 products
 *                     (attention): the six column blocks pair up as the products hh, hm, hl, mh, mm, lh.
 *   lp_parity_norm    LayerNorm (rms = 0, beta may be NULL) / RMSNorm (rms = 1) over fp32 rows.
 *   lp_parity_act     GELU tanh / GELU erf (out may alias x) or SwiGLU on 32-column gate / up granules (out [M, N/2]).
 *   lp_parity_add     y += x (x_row_mod > 0: x row = row % x_row_mod, the position embedding).
 *   lp_parity_rope    rotate-half RoPE in place on heads [0, n_heads) of x fp32 [S, ld].
 *   lp_parity_softmax softmax(scale * scores) per row in place, causal: keys 0 .. row only.
 * ------------------------------------------------------------------------------------------------------------------ */
int lp_parity_split(const float* x, long long ldx, int M, int K, void* out, long long ldo, int layout, void* stream);
int lp_parity_norm(const float* x, long long ldx, const float* gamma, const float* beta, float* out, long long ldo, int M, int D,
                   float eps, int rms, void* stream);
int lp_parity_act(const float* x, long long ldx, float* out, long long ldo, int M, int N, int act, void* stream);
int lp_parity_add(float* y, long long ldy, const float* x, long long ldx, int M, int N, int x_row_mod, void* stream);
int lp_parity_rope(float* x, long long ld, int S, int n_heads, int head_dim, const float* cs, const float* sn, void* stream);
int lp_parity_softmax(float* scores, long long ld, int Tq, int Tk, float scale, int causal, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LEOPARD_B200_H */
