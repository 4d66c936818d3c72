"""ctypes binding of the C ABI declared in include/leopard_b200.h.

There is no CPU fallback: if the shared library is missing or a call fails, a `NativeError` is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libleopard_b200.so"

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SWIGLU = 0, 1, 2, 3
MERGE_HF_EXPAND, MERGE_MEGATRON_SCATTER = 0, 1


class NativeError(RuntimeError):
    pass


_lib = None

_vp, _ll, _i, _fp = C.c_void_p, C.c_longlong, C.c_int, C.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGS = {
    "lp_last_error": [],
    "lp_abi_version": [],
    "lp_gemm_bf16": [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _vp, _ll, _i, _i, _i, _i, _vp],
}
_RESTYPES = {"lp_last_error": C.c_char_p}


def declared_symbols():
    return sorted(_SIGS)


def register(name, argtypes, restype=None):
    _SIGS[name] = argtypes
    if restype is not None:
        _RESTYPES[name] = restype


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
        for name, argtypes in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


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

register("lp_attention_fwd", [_vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, C.c_float, _vp, _ll, _i, _vp])
