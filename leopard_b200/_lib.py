"""ctypes binding of the C ABI declared in include/leopard_b200.h.

There is no CPU fallback: if the shared library is missing or a call fails, a `NativeError` is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libleopard_b200.so"

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SWIGLU = 0, 1, 2, 3
LOGITS_NONE, LOGITS_LAST, LOGITS_ALL = 0, 1, 2


class NativeError(RuntimeError):
    pass


_vp, _ll, _i, _f, _sz = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_size_t


class VitDims(C.Structure):
    _fields_ = [(n, _i) for n in ("layers", "hidden", "ffn", "heads", "head_dim", "head_stride", "grid", "tokens",
                                  "patch_k", "patch_k_pad")] + [("eps", _f)]


class VitLayerWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("ln1_g", "ln1_b", "ln2_g", "ln2_b", "qkv_w", "qkv_b", "out_w", "out_b", "fc1_w",
                                   "fc1_b", "fc2_w", "fc2_b", "qkv_wf", "qkv_c1", "qkv_c2", "fc1_wf", "fc1_c1", "fc1_c2")]


class VitWeights(C.Structure):
    _fields_ = [("patch_w", _vp), ("patch_b", _vp), ("pos", _vp), ("layers", C.POINTER(VitLayerWeights)),
                ("post_g", _vp), ("post_b", _vp)]


class ProjWeights(C.Structure):
    _fields_ = [("w1", _vp), ("b1", _vp), ("w2", _vp), ("b2", _vp), ("act", _i)]


class DecDims(C.Structure):
    _fields_ = [(n, _i) for n in ("layers", "hidden", "ffn", "heads", "kv_heads", "head_dim", "vocab")] + [("eps", _f)]


class DecLayerWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("ln1_g", "qkv_w", "o_w", "ln2_g", "gateup_w", "down_w", "qkv_wf", "gateup_wf")]


class DecWeights(C.Structure):
    _fields_ = [("layers", C.POINTER(DecLayerWeights)), ("final_g", _vp), ("lm_head", _vp), ("inv_freq", _vp)]


class GemmDesc(C.Structure):
    """lp_gemm_desc (include/leopard_b200.h)."""
    _fields_ = [("A", _vp), ("lda", _ll), ("W", _vp), ("ldw", _ll), ("C", _vp), ("ldc", _ll), ("M", _i), ("N", _i), ("K", _i),
                ("bias", _vp), ("residual", _vp), ("ldr", _ll), ("res_row_mod", _i), ("act", _i), ("out_fp32", _i),
                ("block_n", _i), ("norm_mode", _i), ("stats_in", _vp), ("stats_in_parts", _i), ("colsum", _vp),
                ("norm_eps", _f), ("stats_out", _vp), ("stats_out_parts", _i), ("rope_cos", _vp), ("rope_sin", _vp),
                ("rope_heads", _i), ("a_t", _i), ("b_t", _i)]


NORM_NONE, NORM_RMS, NORM_LAYER = 0, 1, 2
TP_ALLGATHER, TP_REDUCE_SCATTER = 1, 2


class TpDesc(C.Structure):
    """lp_tp_desc (include/leopard_b200.h)."""
    _fields_ = [("tp_size", _i), ("tp_rank", _i), ("seg_rows", _i), ("mode", _i), ("epoch", C.c_uint), ("peer_A", _vp * 8),
                ("peer_stage", _vp * 8), ("counter", _vp), ("recycle", _vp), ("signal_local", _vp), ("signal_peer", _vp * 8)]


class ProfileRecord(C.Structure):
    _fields_ = [("kind", _i), ("a", _i), ("b", _i), ("c", _i), ("ms", _f)]


_P = C.POINTER

# name -> (restype, argtypes); must list EVERY symbol include/leopard_b200.h declares (tests/test_abi.py checks)
_SIGS = {
    "lp_last_error": (C.c_char_p, []),
    "lp_abi_version": (_i, []),
    "lp_gemm_bf16": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "lp_gemm_bf16_ex": (_i, [C.POINTER(GemmDesc), _vp]),
    "lp_gemm_pick_block_n": (_i, [_i, _i, _i]),
    "lp_row_stats": (_i, [_vp, _ll, _i, _i, _vp, _vp]),
    "lp_model_set_fused_norm": (_i, [_i]),
    "lp_model_set_chained_projector": (_i, [_i]),
    "lp_gemm_set_pair_mode": (_i, [_i]),
    "lp_gemm_streamk_workspace_bytes": (_sz, []),
    "lp_gemm_set_streamk_workspace": (_i, [_vp, _sz]),
    "lp_gemm_set_streamk": (_i, [_i]),
    "lp_gemm_set_tma_residual": (_i, [_i]),
    "lp_gemm_set_raster_balance": (_i, [_i]),
    "lp_gemm_set_epilogue_sleep": (_i, [_i]),
    "lp_gemm_bf16_tp": (_i, [C.POINTER(GemmDesc), C.POINTER(TpDesc), _vp]),
    "lp_tp_reduce": (_i, [C.POINTER(TpDesc), _vp, _ll, _vp, _ll, _vp, _ll, _i, _vp]),
    "lp_parity_split": (_i, [_vp, _ll, _i, _i, _vp, _ll, _i, _vp]),
    "lp_parity_norm": (_i, [_vp, _ll, _vp, _vp, _vp, _ll, _i, _i, _f, _i, _vp]),
    "lp_parity_act": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "lp_parity_add": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "lp_parity_rope": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _vp]),
    "lp_parity_softmax": (_i, [_vp, _ll, _i, _i, _f, _i, _vp]),
    "lp_set_prefill_pdl": (_i, [_i]),
    "lp_attention_fwd": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _ll, _i, _vp]),
    "lp_attention_set_pingpong": (_i, [_i]),
    "lp_model_set_attention_variant": (_i, [_i]),
    "lp_attention_cross": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _ll, _vp]),
    "lp_layernorm": (_i, [_vp, _ll, _vp, _vp, _vp, _ll, _i, _i, _f, _i, _vp]),
    "lp_rmsnorm": (_i, [_vp, _ll, _vp, _vp, _ll, _i, _i, _f, _vp]),
    "lp_norm_set_variant": (_i, [_i]),
    "lp_rope_table": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "lp_rope_inplace": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lp_embed_merge": (_i, [_vp, _i, _ll, _i, _vp, _ll, _ll, _vp, _ll, _i, _vp, _ll, _i, _i, _vp, _vp]),
    "lp_im2col_patch14": (_i, [_vp, _i, _vp, _ll, _i, _i, _i, _i, _vp]),
    "lp_cross_entropy": (_i, [_vp, _i, _ll, _vp, _vp, _i, _i, _vp]),
    "lp_pixel_shuffle": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "lp_vit_workspace_bytes": (_sz, [_P(VitDims), _i]),
    "lp_vit_forward": (_i, [_P(VitDims), _P(VitWeights), _vp, _i, _i, _vp, _ll, _i, _vp, _sz, _vp]),
    "lp_vit_layers_workspace_bytes": (_sz, [_P(VitDims), _i]),
    "lp_vit_layers": (_i, [_P(VitDims), _P(VitWeights), _vp, _P(_i), _i, _vp, _ll, _vp, _sz, _vp]),
    "lp_project_workspace_bytes": (_sz, [_i, _i]),
    "lp_project": (_i, [_P(ProjWeights), _vp, _i, _i, _i, _vp, _ll, _vp, _sz, _vp]),
    "lp_decoder_workspace_bytes": (_sz, [_P(DecDims), _i, _i]),
    "lp_decoder_prefill": (_i, [_P(DecDims), _P(DecWeights), _vp, _vp, _P(_i), _i, _vp, _ll, _vp, _ll, _i, _i, _vp,
                                _sz, _vp]),
    "lp_decode_workspace_bytes": (_sz, [_P(DecDims)]),
    "lp_decode_step": (_i, [_P(DecDims), _P(DecWeights), _vp, _ll, _vp, _i, _vp, _ll, _vp, _vp, _sz, _vp]),
    "lp_decode_greedy_step": (_i, [_P(DecDims), _P(DecWeights), _vp, _ll, _vp, _vp, _vp, _ll, _vp, _i, _vp, _sz, _vp]),
    "lp_decode_graph_create": (_i, [_P(DecDims), _P(DecWeights), _vp, _ll, _vp, _vp, _vp, _ll, _vp, _i, _vp, _sz,
                                    _P(_vp)]),
    "lp_decode_graph_launch": (_i, [_vp, _vp]),
    "lp_decode_graph_destroy": (None, [_vp]),
    "lp_decode_set_mode": (_i, [_i]),
    "lp_decode_set_prefetch": (_i, [_i, _i]),
    "lp_image_frontend_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "lp_image_frontend": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, C.c_double, _P(_f), _P(_f), _vp, _sz, _vp]),
    "lp_image_resize_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "lp_image_resize": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _vp, C.c_double, _P(_f), _P(_f), _vp, _sz, _vp]),
    "lp_decode_debug_stamps": (_i, [_vp]),
    "lp_attention_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "lp_attention_bwd": (_i, [_vp, _ll, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _ll, _vp, _sz, _vp]),
    "lp_transpose_bf16": (_i, [_vp, _ll, _i, _i, _vp, _ll, _vp]),
    "lp_layernorm_bwd_workspace_bytes": (_sz, [_i]),
    "lp_layernorm_bwd": (_i, [_vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _i, _i, _f, _vp, _sz, _vp]),
    "lp_pixel_unshuffle": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "lp_rowmod_sum": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp]),
    "lp_rmsnorm_bwd_workspace_bytes": (_sz, [_i]),
    "lp_rmsnorm_bwd": (_i, [_vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _i, _i, _f, _vp, _sz, _vp]),
    "lp_swiglu_fwd": (_i, [_vp, _ll, _vp, _ll, _i, _i, _vp]),
    "lp_swiglu_bwd": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _vp]),
    "lp_gelu_bwd": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "lp_colsum_bf16": (_i, [_vp, _ll, _i, _i, _vp, _vp]),
    "lp_ce_fwd_bwd": (_i, [_vp, _i, _ll, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "lp_embed_bwd": (_i, [_vp, _ll, _vp, _i, _ll, _vp, _ll, _i, _ll, _vp]),
    "lp_gather_rows_bf16": (_i, [_vp, _ll, _vp, _i, _i, _vp, _ll, _vp]),
    "lp_softmax_bwd": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _f, _vp]),
    "lp_accumulate_f32": (_i, [_vp, _vp, _ll, _vp]),
    "lp_grad_norm_workspace_bytes": (_sz, []),
    "lp_sumsq_accum": (_i, [_vp, _ll, _vp, _vp]),
    "lp_grad_norm_finish": (_i, [_vp, _f, _vp, _vp]),
    "lp_adamw_step": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _vp, _vp]),
    "lp_launch_count": (_ll, []),
    "lp_profile_enable": (_i, [_i]),
    "lp_profile_read": (_i, [_vp, _i]),
}

_lib = None


def declared_symbols():
    return sorted(_SIGS)


def lib():
    """Load (once) and return the native library; raises NativeError when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -m leopard_b200.build` "
                "(or __graft_entry__.build()); leopard_b200 has no CPU fallback")
        try:
            L = C.CDLL(str(LIB_PATH))
        except OSError as e:  # pragma: no cover
            raise NativeError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (restype, argtypes) in _SIGS.items():
            try:
                fn = getattr(L, name)
            except AttributeError as e:
                raise NativeError(f"{LIB_PATH} does not export {name}: rebuild the library") from e
            fn.argtypes = argtypes
            fn.restype = restype
        import os
        if os.environ.get("LEOPARD_PREFILL_PDL") is not None:  # A/B switch for the prefill kernels' launch mode
            L.lp_set_prefill_pdl(int(os.environ["LEOPARD_PREFILL_PDL"]))
        _lib = L
    return _lib


_streamk_ws = None   # the device buffer handed to lp_gemm_set_streamk_workspace (kept alive here)


def ensure_streamk_workspace(device=None):
    """Give the GEMM its stream-K workspace (include/leopard_b200.h: lp_gemm_set_streamk_workspace) on the process's GPU — one
    process per GPU, one compute stream, as everywhere in this package.  LEOPARD_STREAMK=0 leaves the stream-K tail off."""
    global _streamk_ws
    import os
    import torch
    L = lib()
    if os.environ.get("LEOPARD_STREAMK", "1") == "0" or not torch.cuda.is_available():
        return None
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.index is None:                   # "cuda" and "cuda:<current>" are the same device
        dev = torch.device("cuda", torch.cuda.current_device())
    if _streamk_ws is not None:
        if _streamk_ws.device != dev:      # a second device in one process: the single workspace cannot serve both
            check(L.lp_gemm_set_streamk_workspace(None, 0), "lp_gemm_set_streamk_workspace")
        return _streamk_ws
    with torch.cuda.device(dev):
        ws = torch.zeros(int(L.lp_gemm_streamk_workspace_bytes()), dtype=torch.uint8, device=dev)
    check(L.lp_gemm_set_streamk_workspace(C.c_void_p(ws.data_ptr()), ws.numel()), "lp_gemm_set_streamk_workspace")
    _streamk_ws = ws
    return ws


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().lp_last_error()
        raise NativeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
