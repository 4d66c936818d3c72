"""Thin torch-tensor wrappers over the C ABI (one Python function per exported kernel entry point).

These add shape/dtype checks and output allocation only; no arithmetic happens in Python.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SWIGLU  # noqa: F401


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
