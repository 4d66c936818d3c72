"""Thin torch-tensor wrappers over the C ABI (one Python function per exported kernel entry point).

These add shape/dtype checks and output allocation only; no arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SWIGLU, NORM_NONE, NORM_RMS, NORM_LAYER  # noqa: F401


def _chk_bf16(t, name):
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA bf16 tensor, got {t.dtype} on {t.device}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must have a contiguous last dimension")


def gemm(a, w, bias=None, residual=None, act=ACT_NONE, out=None, out_fp32=False, res_row_mod=0, block_n=0):
    """out[M,N'] = act(a[M,K] @ w[N,K].T + bias) + residual   (see lp_gemm_bf16)."""
    _chk_bf16(a, "a")
    _chk_bf16(w, "w")
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise ValueError(f"K mismatch {K} vs {K2}")
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N):
        raise TypeError("bias must be fp32 [N]")
    if residual is not None:
        _chk_bf16(residual, "residual")
    rc = _lib.lib().lp_gemm_bf16(
        _lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(out), out.stride(0), M, N, K,
        _lib.ptr(bias), _lib.ptr(residual), residual.stride(0) if residual is not None else 0, res_row_mod,
        act, 1 if out.dtype == torch.float32 else 0, block_n, _lib.stream_ptr())
    _lib.check(rc, "lp_gemm_bf16")
    return out


def gemm_pick_block_n(M, N, act=ACT_NONE):
    return int(_lib.lib().lp_gemm_pick_block_n(M, N, act))


def row_stats(x, out=None):
    """(sum, sum of squares) per row of x bf16 [rows, width] -> fp32 [rows, 1, 2] (lp_row_stats): one partial per row."""
    _chk_bf16(x, "x")
    rows, width = x.shape
    if out is None:
        out = torch.empty((rows, 1, 2), device=x.device, dtype=torch.float32)
    rc = _lib.lib().lp_row_stats(_lib.ptr(x), x.stride(0), rows, width, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lp_row_stats")
    return out


def gemm_ex(a, w, bias=None, residual=None, act=ACT_NONE, out=None, out_fp32=False, res_row_mod=0, block_n=0, norm_mode=0,
            stats_in=None, colsum=None, eps=0.0, want_stats=False, rope=None, rope_heads=0, a_t=False, b_t=False):
    """lp_gemm_bf16_ex: lp_gemm_bf16 plus the epilogue fusions — `norm_mode` (NORM_RMS / NORM_LAYER) normalises the rows of
    `a` from `stats_in` fp32 [M, parts, 2] (w must be the gain-folded matrix; LAYER also needs `colsum` and the folded bias),
    `want_stats` returns (out, stats [M, parts, 2]) of the bf16 outputs, `rope` = (cos, sin) fp32 [M, 64] rotates the first
    `rope_heads` 128-column heads.  `a_t` / `b_t`: the operand is given transposed — a as At [K, M], w as Wt [K, N] — and is
    consumed as an MN-major tensor-core operand without a transposed copy (the SFT step's dgrad / wgrad)."""
    _chk_bf16(a, "a")
    _chk_bf16(w, "w")
    K, M = a.shape if a_t else a.shape[::-1]
    N = w.shape[1] if b_t else w.shape[0]
    if (w.shape[0] if b_t else w.shape[1]) != K:
        raise ValueError(f"K mismatch: a {tuple(a.shape)} (a_t={a_t}) vs w {tuple(w.shape)} (b_t={b_t})")
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    d = _lib.GemmDesc()
    d.A, d.lda, d.W, d.ldw, d.C, d.ldc = _lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(out), out.stride(0)
    d.M, d.N, d.K = M, N, K
    d.bias = _lib.ptr(bias)
    d.residual, d.ldr, d.res_row_mod = _lib.ptr(residual), (residual.stride(0) if residual is not None else 0), res_row_mod
    d.act, d.out_fp32, d.block_n = act, 1 if out.dtype == torch.float32 else 0, block_n
    d.norm_mode = norm_mode
    if stats_in is not None:
        d.stats_in, d.stats_in_parts = _lib.ptr(stats_in), stats_in.shape[1]
    d.colsum, d.norm_eps = _lib.ptr(colsum), float(eps)
    stats = None
    if want_stats:
        bn = block_n or gemm_pick_block_n(M, N, act)
        d.block_n = bn
        parts = 2 * (-(-N // bn))
        stats = torch.empty((M, parts, 2), device=a.device, dtype=torch.float32)
        d.stats_out, d.stats_out_parts = _lib.ptr(stats), parts
    if rope is not None:
        d.rope_cos, d.rope_sin, d.rope_heads = _lib.ptr(rope[0]), _lib.ptr(rope[1]), rope_heads
    d.a_t, d.b_t = int(bool(a_t)), int(bool(b_t))
    rc = _lib.lib().lp_gemm_bf16_ex(C.byref(d), _lib.stream_ptr())
    _lib.check(rc, "lp_gemm_bf16_ex")
    return (out, stats) if want_stats else out


def attention(qkv, B, T, n_heads, n_kv_heads, head_dim, head_stride, q_off, k_off, v_off, causal, scale=None,
              out=None, variant=0):
    """out[B*T, n_heads*head_dim] = softmax(scale q k^T (+causal)) v on column slices of `qkv` (lp_attention_fwd)."""
    _chk_bf16(qkv, "qkv")
    if qkv.shape[0] != B * T:
        raise ValueError(f"qkv has {qkv.shape[0]} rows, expected B*T = {B * T}")
    if scale is None:
        scale = head_dim ** -0.5
    if out is None:
        out = torch.empty((B * T, n_heads * head_dim), device=qkv.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_attention_fwd(_lib.ptr(qkv), qkv.stride(0), B, T, n_heads, n_kv_heads, head_dim, head_stride,
                                     q_off, k_off, v_off, 1 if causal else 0, float(scale), _lib.ptr(out),
                                     out.stride(0), variant, _lib.stream_ptr())
    _lib.check(rc, "lp_attention_fwd")
    return out


def _chk_f32(t, name):
    if t is not None and (t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous CUDA fp32 tensor")


def layernorm(x, gamma, beta, eps, shuffle_side=0, out=None):
    """LayerNorm over the last dim of x [M, D]; shuffle_side=26 stores the pixel-shuffled layout [M/4, 4D]."""
    _chk_bf16(x, "x")
    _chk_f32(gamma, "gamma")
    _chk_f32(beta, "beta")
    M, D = x.shape
    if out is None:
        shape = (M // 4, 4 * D) if shuffle_side else (M, D)
        out = torch.empty(shape, device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_layernorm(_lib.ptr(x), x.stride(0), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(out),
                                 out.stride(0), M, D, float(eps), shuffle_side, _lib.stream_ptr())
    _lib.check(rc, "lp_layernorm")
    return out


def rmsnorm(x, gamma, eps, out=None):
    _chk_bf16(x, "x")
    _chk_f32(gamma, "gamma")
    M, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    rc = _lib.lib().lp_rmsnorm(_lib.ptr(x), x.stride(0), _lib.ptr(gamma), _lib.ptr(out), out.stride(0), M, D,
                               float(eps), _lib.stream_ptr())
    _lib.check(rc, "lp_rmsnorm")
    return out


def rope_table(pos, inv_freq):
    """pos int32 [S], inv_freq fp32 [half] -> (cos, sin) fp32 [S, half]."""
    S, half = pos.numel(), inv_freq.numel()
    cs = torch.empty((S, half), device=pos.device, dtype=torch.float32)
    sn = torch.empty_like(cs)
    rc = _lib.lib().lp_rope_table(_lib.ptr(pos), _lib.ptr(inv_freq), _lib.ptr(cs), _lib.ptr(sn), S, half,
                                  _lib.stream_ptr())
    _lib.check(rc, "lp_rope_table")
    return cs, sn


def rope_inplace(x, col0, n_heads, head_dim, cs, sn):
    _chk_bf16(x, "x")
    rc = _lib.lib().lp_rope_inplace(_lib.ptr(x), x.stride(0), x.shape[0], col0, n_heads, head_dim, _lib.ptr(cs),
                                    _lib.ptr(sn), _lib.stream_ptr())
    _lib.check(rc, "lp_rope_inplace")
    return x


def im2col_patch14(pixels, ldo=592):
    """pixels [N,C,H,W] fp32/bf16 -> bf16 [N*(H/14)*(W/14), ldo]."""
    N, Cc, H, W = pixels.shape
    pixels = pixels.contiguous()
    out = torch.empty((N * (H // 14) * (W // 14), ldo), device=pixels.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_im2col_patch14(_lib.ptr(pixels), 1 if pixels.dtype == torch.float32 else 0, _lib.ptr(out), ldo,
                                      N, Cc, H, W, _lib.stream_ptr())
    _lib.check(rc, "lp_im2col_patch14")
    return out


def embed_merge(ids, table, feats, image_id, rows_per_slot):
    """ids int64 [L]; table bf16 [V, D]; feats bf16 [rows, D] or None -> (out [S, D], n_slots, S) (lp_embed_merge)."""
    L = ids.numel()
    D = table.shape[1]
    n_feat = 0 if feats is None else feats.shape[0]
    cap = L + (n_feat // rows_per_slot) * (rows_per_slot - 1)
    out = torch.zeros((cap, D), device=table.device, dtype=torch.bfloat16)
    plan = torch.empty(2 * L + 2, device=table.device, dtype=torch.int32)
    rc = _lib.lib().lp_embed_merge(_lib.ptr(ids), L, image_id, rows_per_slot, _lib.ptr(table), table.stride(0),
                                   table.shape[0], _lib.ptr(feats), feats.stride(0) if feats is not None else 0,
                                   n_feat, _lib.ptr(out), out.stride(0), cap, D, _lib.ptr(plan), _lib.stream_ptr())
    _lib.check(rc, "lp_embed_merge")
    n_slots, S = plan[2 * L:].tolist()
    return out, n_slots, S


def pixel_shuffle(x2d, side):
    """x2d bf16 [B*side*side, D] -> [B*side*side/4, 4*D] (lp_pixel_shuffle)."""
    _chk_bf16(x2d, "x")
    M, D = x2d.shape
    out = torch.empty((M // 4, 4 * D), device=x2d.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_pixel_shuffle(_lib.ptr(x2d), x2d.stride(0), _lib.ptr(out), out.stride(0), M, D, side,
                                     _lib.stream_ptr())
    _lib.check(rc, "lp_pixel_shuffle")
    return out


def attention_cross(q, kv, B, Tq, Tk, n_heads, n_kv_heads, head_dim, k_off, v_off, scale=None, out=None):
    """Cross-attention: q [B*Tq, heads*head_dim], kv [B*Tk, ...] with K at k_off, V at v_off (lp_attention_cross)."""
    _chk_bf16(q, "q")
    _chk_bf16(kv, "kv")
    if scale is None:
        scale = head_dim ** -0.5
    if out is None:
        out = torch.empty((B * Tq, n_heads * head_dim), device=q.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_attention_cross(_lib.ptr(q), q.stride(0), _lib.ptr(kv), kv.stride(0), B, Tq, Tk, n_heads,
                                       n_kv_heads, head_dim, k_off, v_off, float(scale), _lib.ptr(out), out.stride(0),
                                       _lib.stream_ptr())
    _lib.check(rc, "lp_attention_cross")
    return out


def cross_entropy(logits, labels):
    """per-row CE in fp32: logits [R, V] (bf16 or fp32), labels int64 [R] (negative = ignored -> 0) (lp_cross_entropy)."""
    R, V = logits.shape
    out = torch.empty((R,), device=logits.device, dtype=torch.float32)
    labels = labels.to(logits.device, torch.long).contiguous()
    rc = _lib.lib().lp_cross_entropy(_lib.ptr(logits), 1 if logits.dtype == torch.float32 else 0, logits.stride(0),
                                     _lib.ptr(labels), _lib.ptr(out), R, V, _lib.stream_ptr())
    _lib.check(rc, "lp_cross_entropy")
    return out


# ------------------------------------------------------------------------------------------------------------------
# SFT step (lp_train.cu): backward of the glue ops, masked LM loss, optimizer.  Same rule: allocation and checks only.
# ------------------------------------------------------------------------------------------------------------------
def transpose(x, out=None):
    """out[c, r] = x[r, c] for a bf16 matrix view (last dim contiguous, even leading dimension) (lp_transpose_bf16)."""
    _chk_bf16(x, "x")
    rows, cols = x.shape
    if out is None:   # leading dimension padded to 8 elements (the GEMM's TMA maps need 16-byte row strides)
        out = torch.empty((cols, -(-rows // 8) * 8), device=x.device, dtype=torch.bfloat16)[:, :rows]
    rc = _lib.lib().lp_transpose_bf16(_lib.ptr(x), x.stride(0), rows, cols, _lib.ptr(out), out.stride(0), _lib.stream_ptr())
    _lib.check(rc, "lp_transpose_bf16")
    return out


_rms_bwd_ws = {}


def rmsnorm_bwd(x, gamma, dy, eps, dgamma, add=None, out=None):
    """dx = rmsnorm'(x; gamma) dy (+ add); dgamma (fp32 [D]) is overwritten (lp_rmsnorm_bwd)."""
    _chk_bf16(x, "x")
    _chk_bf16(dy, "dy")
    _chk_f32(gamma, "gamma")
    _chk_f32(dgamma, "dgamma")
    M, D = x.shape
    L = _lib.lib()
    ws = _rms_bwd_ws.get((x.device, D))
    if ws is None:
        ws = torch.empty(L.lp_rmsnorm_bwd_workspace_bytes(D), dtype=torch.uint8, device=x.device)
        _rms_bwd_ws[(x.device, D)] = ws
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16)
    rc = L.lp_rmsnorm_bwd(_lib.ptr(x), x.stride(0), _lib.ptr(gamma), _lib.ptr(dy), dy.stride(0), _lib.ptr(add),
                          add.stride(0) if add is not None else 0, _lib.ptr(out), out.stride(0), _lib.ptr(dgamma), M, D,
                          float(eps), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "lp_rmsnorm_bwd")
    return out


_ln_bwd_ws = {}


def layernorm_bwd(x, gamma, dy, eps, dgamma, dbeta, add=None, out=None):
    """dx = layernorm'(x; gamma) dy (+ add); dgamma, dbeta (fp32 [D]) are overwritten (lp_layernorm_bwd)."""
    _chk_bf16(x, "x")
    _chk_bf16(dy, "dy")
    for t, n in ((gamma, "gamma"), (dgamma, "dgamma"), (dbeta, "dbeta")):
        _chk_f32(t, n)
    M, D = x.shape
    L = _lib.lib()
    ws = _ln_bwd_ws.get((x.device, D))
    if ws is None:
        ws = torch.empty(L.lp_layernorm_bwd_workspace_bytes(D), dtype=torch.uint8, device=x.device)
        _ln_bwd_ws[(x.device, D)] = ws
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16)
    rc = L.lp_layernorm_bwd(_lib.ptr(x), x.stride(0), _lib.ptr(gamma), _lib.ptr(dy), dy.stride(0), _lib.ptr(add),
                            add.stride(0) if add is not None else 0, _lib.ptr(out), out.stride(0), _lib.ptr(dgamma),
                            _lib.ptr(dbeta), M, D, float(eps), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "lp_layernorm_bwd")
    return out


def pixel_unshuffle(d, side):
    """d bf16 [M/4, 4*D] (gradient of the shuffled layout) -> [M, D] (lp_pixel_unshuffle)."""
    _chk_bf16(d, "d")
    M4, D4 = d.shape
    out = torch.empty((M4 * 4, D4 // 4), device=d.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_pixel_unshuffle(_lib.ptr(d), d.stride(0), _lib.ptr(out), out.stride(0), M4 * 4, D4 // 4, side,
                                       _lib.stream_ptr())
    _lib.check(rc, "lp_pixel_unshuffle")
    return out


def rowmod_sum(d, T, out):
    """out[t] = sum_n d[n*T + t] (fp32 [T, D], overwritten) (lp_rowmod_sum)."""
    _chk_bf16(d, "d")
    _chk_f32(out, "out")
    rc = _lib.lib().lp_rowmod_sum(_lib.ptr(d), d.stride(0), d.shape[0] // T, T, d.shape[1], _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lp_rowmod_sum")
    return out


def swiglu_fwd(gu, out=None):
    """gu bf16 [M, 2F] in the fused weight's interleaved column order -> silu(gate) * up [M, F] (lp_swiglu_fwd)."""
    _chk_bf16(gu, "gu")
    M, F2 = gu.shape
    if out is None:
        out = torch.empty((M, F2 // 2), device=gu.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_swiglu_fwd(_lib.ptr(gu), gu.stride(0), _lib.ptr(out), out.stride(0), M, F2 // 2, _lib.stream_ptr())
    _lib.check(rc, "lp_swiglu_fwd")
    return out


def swiglu_bwd(gu, dm, out=None):
    """gradient w.r.t. gu (same interleaved layout) given dm [M, F] (lp_swiglu_bwd)."""
    _chk_bf16(gu, "gu")
    _chk_bf16(dm, "dm")
    M, F2 = gu.shape
    if out is None:
        out = torch.empty((M, F2), device=gu.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_swiglu_bwd(_lib.ptr(gu), gu.stride(0), _lib.ptr(dm), dm.stride(0), _lib.ptr(out), out.stride(0), M,
                                  F2 // 2, _lib.stream_ptr())
    _lib.check(rc, "lp_swiglu_bwd")
    return out


def gelu_bwd(z, dy, act, out=None):
    _chk_bf16(z, "z")
    _chk_bf16(dy, "dy")
    M, N = z.shape
    if out is None:
        out = torch.empty((M, N), device=z.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_gelu_bwd(_lib.ptr(z), z.stride(0), _lib.ptr(dy), dy.stride(0), _lib.ptr(out), out.stride(0), M, N,
                                act, _lib.stream_ptr())
    _lib.check(rc, "lp_gelu_bwd")
    return out


def colsum(dy, out):
    """out[c] = sum_r dy[r, c] (fp32 [N], overwritten) (lp_colsum_bf16)."""
    _chk_bf16(dy, "dy")
    _chk_f32(out, "out")
    M, N = dy.shape
    rc = _lib.lib().lp_colsum_bf16(_lib.ptr(dy), dy.stride(0), M, N, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lp_colsum_bf16")
    return out


def ce_fwd_bwd(logits, target, weight, vocab):
    """logits [R, Vp] (fp32 or bf16; columns >= vocab are padding), target int64 [R], weight fp32 [R] ->
    (loss [R] fp32, dlogits bf16 [R, Vp]) (lp_ce_fwd_bwd)."""
    R, Vp = logits.shape
    loss = torch.empty((R,), device=logits.device, dtype=torch.float32)
    d = torch.empty((R, Vp), device=logits.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_ce_fwd_bwd(_lib.ptr(logits), 1 if logits.dtype == torch.float32 else 0, logits.stride(0),
                                  _lib.ptr(target), _lib.ptr(weight), _lib.ptr(loss), _lib.ptr(d), d.stride(0), R, vocab, Vp,
                                  _lib.stream_ptr())
    _lib.check(rc, "lp_ce_fwd_bwd")
    return loss, d


def embed_bwd(d, ids, image_id, dE):
    """dE[id] += d[r] for text rows (lp_embed_bwd); dE fp32 [V, D] must hold the running gradient (zeroed per step)."""
    _chk_bf16(d, "d")
    rc = _lib.lib().lp_embed_bwd(_lib.ptr(d), d.stride(0), _lib.ptr(ids), ids.numel(), image_id, _lib.ptr(dE), dE.stride(0),
                                 d.shape[1], dE.shape[0], _lib.stream_ptr())
    _lib.check(rc, "lp_embed_bwd")


def gather_rows(x, idx):
    """x bf16 [R, D], idx int32 [n] -> [n, D] (lp_gather_rows_bf16)."""
    _chk_bf16(x, "x")
    out = torch.empty((idx.numel(), x.shape[1]), device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_gather_rows_bf16(_lib.ptr(x), x.stride(0), _lib.ptr(idx), idx.numel(), x.shape[1], _lib.ptr(out),
                                        out.stride(0), _lib.stream_ptr())
    _lib.check(rc, "lp_gather_rows_bf16")
    return out


def softmax_bwd(scores, dp, causal, scale, P=None, dS=None):
    """scores, dp fp32 [Sq, Sk] -> (P, dS) bf16 [Sq, Sk] (lp_softmax_bwd)."""
    Sq, Sk = scores.shape
    if P is None:
        P = torch.empty((Sq, Sk), device=scores.device, dtype=torch.bfloat16)
    if dS is None:
        dS = torch.empty((Sq, Sk), device=scores.device, dtype=torch.bfloat16)
    rc = _lib.lib().lp_softmax_bwd(_lib.ptr(scores), scores.stride(0), _lib.ptr(dp), dp.stride(0), _lib.ptr(P), P.stride(0),
                                   _lib.ptr(dS), dS.stride(0), Sq, Sk, 1 if causal else 0, float(scale), _lib.stream_ptr())
    _lib.check(rc, "lp_softmax_bwd")
    return P, dS


def accumulate(dst, src):
    """dst += src for contiguous fp32 tensors of equal size (lp_accumulate_f32)."""
    _chk_f32(dst, "dst")
    _chk_f32(src, "src")
    if dst.numel() != src.numel():
        raise ValueError("size mismatch")
    _lib.check(_lib.lib().lp_accumulate_f32(_lib.ptr(dst), _lib.ptr(src), dst.numel(), _lib.stream_ptr()),
               "lp_accumulate_f32")
    return dst


def grad_norm(flat_grads, max_norm, partial=None, out2=None):
    """Global L2 norm of the fp32 tensors in `flat_grads` -> device tensor [norm, clip coefficient]."""
    L = _lib.lib()
    dev = flat_grads[0].device
    if partial is None:
        partial = torch.empty(L.lp_grad_norm_workspace_bytes() // 4, dtype=torch.float32, device=dev)
    if out2 is None:
        out2 = torch.empty(2, dtype=torch.float32, device=dev)
    partial.zero_()
    for g in flat_grads:
        _chk_f32(g, "grad")
        _lib.check(L.lp_sumsq_accum(_lib.ptr(g), g.numel(), _lib.ptr(partial), _lib.stream_ptr()), "lp_sumsq_accum")
    _lib.check(L.lp_grad_norm_finish(_lib.ptr(partial), float(max_norm), _lib.ptr(out2), _lib.stream_ptr()),
               "lp_grad_norm_finish")
    return out2


def adamw_step(p, g, m, v, p16, step, lr, beta1, beta2, eps, weight_decay, norm_and_clip):
    """In-place FusedAdam (adam_w_mode) update of fp32 p / m / v from fp32 g; refreshes the bf16 copy p16 (or None)."""
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _chk_f32(t, n)
    rc = _lib.lib().lp_adamw_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.ptr(p16), p.numel(), float(lr),
                                  float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                                  _lib.ptr(norm_and_clip), _lib.stream_ptr())
    _lib.check(rc, "lp_adamw_step")


_attn_bwd_ws = {}


def attention_bwd(qkv, out, dout, B, T, n_heads, n_kv_heads, head_dim, q_off, k_off, v_off, causal, scale=None, dqkv=None,
                  head_stride=None):
    """Flash-style backward of `attention` (lp_attention_bwd): returns dqkv in the qkv column layout."""
    _chk_bf16(qkv, "qkv")
    _chk_bf16(out, "out")
    _chk_bf16(dout, "dout")
    if out.stride(0) != dout.stride(0):
        raise ValueError("out and dout must share their leading dimension")
    if scale is None:
        scale = head_dim ** -0.5
    if dqkv is None:
        dqkv = torch.zeros_like(qkv)
    L = _lib.lib()
    need = L.lp_attention_bwd_workspace_bytes(B, T, n_heads)
    ws = _attn_bwd_ws.get(qkv.device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=qkv.device)
        _attn_bwd_ws[qkv.device] = ws
    rc = L.lp_attention_bwd(_lib.ptr(qkv), qkv.stride(0), _lib.ptr(out), _lib.ptr(dout), out.stride(0), B, T, n_heads,
                            n_kv_heads, head_dim, head_stride or head_dim, q_off, k_off, v_off, 1 if causal else 0,
                            float(scale), _lib.ptr(dqkv),
                            dqkv.stride(0), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "lp_attention_bwd")
    return dqkv


# ---- parity mode (csrc/lp_parity.cu; sequencing in leopard_b200/parity.py) ---------------------------------------------------
def _chk_f32_rows(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 2 or t.stride(1) != 1:
        raise TypeError(f"{name} must be a CUDA fp32 matrix with contiguous rows")


def parity_split(x, out, layout):
    """x fp32 [M, K] (row stride arbitrary: column slices allowed) -> out bf16 [>= M, >= n K] three-term split along K
    (lp_parity_split; layouts in the header)."""
    _chk_f32_rows(x, "x")
    M, K = x.shape
    _lib.check(_lib.lib().lp_parity_split(_lib.ptr(x), x.stride(0), M, K, _lib.ptr(out), out.stride(0), layout, _lib.stream_ptr()),
               "lp_parity_split")
    return out


def parity_norm(x, gamma, beta, eps, rms):
    _chk_f32(x, "x")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().lp_parity_norm(_lib.ptr(x), x.stride(0), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(out), out.stride(0),
                                         x.shape[0], x.shape[1], float(eps), 1 if rms else 0, _lib.stream_ptr()), "lp_parity_norm")
    return out


def parity_act(x, kind):
    _chk_f32(x, "x")
    n_out = x.shape[1] // 2 if kind == ACT_SWIGLU else x.shape[1]
    out = torch.empty((x.shape[0], n_out), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().lp_parity_act(_lib.ptr(x), x.stride(0), _lib.ptr(out), out.stride(0), x.shape[0], x.shape[1], kind,
                                        _lib.stream_ptr()), "lp_parity_act")
    return out


def parity_add_(y, x, row_mod=0):
    _chk_f32(y, "y")
    _chk_f32(x, "x")
    _lib.check(_lib.lib().lp_parity_add(_lib.ptr(y), y.stride(0), _lib.ptr(x), x.stride(0), y.shape[0], y.shape[1], row_mod,
                                        _lib.stream_ptr()), "lp_parity_add")
    return y


def parity_rope_(x, n_heads, head_dim, cs, sn):
    _chk_f32(x, "x")
    _lib.check(_lib.lib().lp_parity_rope(_lib.ptr(x), x.stride(0), x.shape[0], n_heads, head_dim, _lib.ptr(cs), _lib.ptr(sn),
                                         _lib.stream_ptr()), "lp_parity_rope")
    return x


def parity_softmax_(scores, Tk, scale, causal):
    _chk_f32(scores, "scores")
    _lib.check(_lib.lib().lp_parity_softmax(_lib.ptr(scores), scores.stride(0), scores.shape[0], Tk, float(scale),
                                            1 if causal else 0, _lib.stream_ptr()), "lp_parity_softmax")
    return scores
