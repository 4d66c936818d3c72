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

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LP_ABI_VERSION 1

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
 * block_n: 128, 256, or 0 = choose.  Requires N % 8 == 0, lda/ldw % 8 == 0, 16-byte aligned bases. */
int lp_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N,
                 int K, const float* bias, const void* residual, long long ldr, int res_row_mod, int act, int out_fp32,
                 int block_n, void* stream);

/* O[b,t,h,:] = softmax(scale * Q K^T (+causal)) V          (tcgen05 flash attention, fp32 softmax, bf16 in/out)
 *   replaces flash_attn_varlen_func (MP/model/llava/transformer.py:476-515,869-877; the 4x KV repeat_interleave at
 *   :829-836 is not materialised) and the ViT's TEDotProductAttention / HF SiglipAttention
 *   (MP/model/llava/clip_encoder.py:58-63).
 * qkv: bf16 [B*T, ld]; for head h: q cols q_off + h*head_stride, k cols k_off + (h / (n_heads/n_kv_heads))*head_stride,
 *   v likewise from v_off.  Supported (head_dim, head_stride): (128,128) and (72,80) — the latter expects the 8 pad
 *   columns of q and k to be exact zeros (the ViT qkv weight is stored zero-padded).
 * out: bf16 [B*T, ldo], head h at cols h*head_dim.  scale = head_dim^-0.5 in the reference.
 * variant bit 0: keep P in TMEM (tcgen05.mma A-from-TMEM) instead of shared memory. */
int lp_attention_fwd(const void* qkv, long long ld, int B, int T, int n_heads, int n_kv_heads, int head_dim,
                     int head_stride, int q_off, int k_off, int v_off, int causal, float scale, void* out,
                     long long ldo, int variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LEOPARD_B200_H */
