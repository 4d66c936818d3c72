"""not gpu: the C-ABI library builds, loads and exports exactly what include/leopard_b200.h declares; ctypes
signatures cover every symbol; argument errors are reported as status codes (no compute without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    txt = (ROOT / "include" / "leopard_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lp_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound(native_lib):
    from leopard_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 17
    assert syms == _lib.declared_symbols(), "ctypes table and header disagree"
    for s in syms:
        assert hasattr(native_lib, s), s
    assert native_lib.lp_abi_version() == 3


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors have the size and field offsets a C compiler gives the header's structs (gcc on include/leopard_b200.h)."""
    import subprocess
    from leopard_b200 import _lib
    pairs = {"lp_vit_dims": _lib.VitDims, "lp_vit_layer_weights": _lib.VitLayerWeights, "lp_vit_weights": _lib.VitWeights,
             "lp_proj_weights": _lib.ProjWeights, "lp_dec_dims": _lib.DecDims, "lp_dec_layer_weights": _lib.DecLayerWeights,
             "lp_dec_weights": _lib.DecWeights, "lp_gemm_desc": _lib.GemmDesc, "lp_tp_desc": _lib.TpDesc,
             "lp_profile_record": _lib.ProfileRecord}
    lines = []
    for cname, ct in pairs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "leopard_b200.h"\nint main(void) {\n' + "\n".join(lines)
                   + "\nreturn 0; }\n")
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, ct in pairs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_argument_errors_are_status_codes(native_lib):
    rc = native_lib.lp_gemm_bf16(None, 8, None, 8, None, 8, 0, 8, 8, None, None, 0, 0, 0, 0, 0, None)
    assert rc == -1 and b"empty problem" in native_lib.lp_last_error()
    rc = native_lib.lp_attention_fwd(None, 64, 1, 4, 1, 1, 64, 64, 0, 64, 128, 1, 0.125, None, 64, 0, None)
    assert rc == -1 and b"head_dim" in native_lib.lp_last_error()
    rc = native_lib.lp_layernorm(None, 8, None, None, None, 8, 4, 5000, 1e-6, 0, None)
    assert rc == -1


def test_workspace_queries_are_host_only(native_lib):
    from leopard_b200 import _lib
    from leopard_b200.config import leopard_llava
    from leopard_b200.native_weights import HEAD_STRIDE, PATCH_K_PAD
    v, d = leopard_llava().vit, leopard_llava().dec
    vd = _lib.VitDims(v.layers, v.hidden, v.ffn, v.heads, v.head_dim, HEAD_STRIDE, v.grid, v.tokens, v.patch_k,
                      PATCH_K_PAD, v.eps)
    n28 = native_lib.lp_vit_workspace_bytes(C.byref(vd), 28)
    base = 28 * 676 * (592 + 1152 * 3 + 3840 + 4304) * 2          # activations; + two row-statistics buffers (32 float2 per row)
    assert base <= n28 < 1.01 * (base + 2 * 28 * 676 * 32 * 8) + 4096
    dd = _lib.DecDims(d.layers, d.hidden, d.ffn, d.heads, d.kv_heads, d.head_dim, d.vocab, d.eps)
    n = native_lib.lp_decoder_workspace_bytes(C.byref(dd), 4796, 1)
    assert n >= 4796 * (4096 * 2 + 6144 + 14336) * 2


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from leopard_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.NativeError, match="no CPU fallback"):
        _lib.lib()


def test_engine_refuses_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from leopard_b200 import _lib
    from leopard_b200.config import leopard_llava, reduced
    from leopard_b200.engine import PrefillEngine
    from leopard_b200.native_weights import NativeWeights
    cfg = reduced(leopard_llava(), 1, 1, 1024)
    with pytest.raises(_lib.NativeError, match="no CPU fallback"):
        PrefillEngine(cfg, NativeWeights(cfg, "cpu"))


def _header_prototypes():
    """name -> list of C parameter type strings, parsed from the header (comments stripped)."""
    txt = (ROOT / "include" / "leopard_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(lp_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", txt, flags=re.S):
        params = " ".join(m.group(2).split())
        protos[m.group(1)] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    return protos


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and argument class (pointer / 64-bit / int / float / double / size_t) of every binding against the C prototype: a
    mismatch would not fail to load — it would corrupt arguments on the GPU box."""
    from leopard_b200 import _lib

    def klass(ctype_decl: str) -> str:
        d = ctype_decl
        if "*" in d:
            return "ptr"
        base = d.rsplit(" ", 1)[0] if " " in d else d
        if "long long" in base:
            return "ll"
        if "size_t" in base:
            return "sz"
        if "double" in base:
            return "f64"
        if "float" in base:
            return "f32"
        return "int"

    def klass_ct(t) -> str:
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) not in (None, "q", "i", "f", "d", "Q", "L", "l"):
            return "ptr"
        return {C.c_longlong: "ll", C.c_int: "int", C.c_float: "f32", C.c_double: "f64", C.c_size_t: "sz"}[t]

    protos = _header_prototypes()
    assert set(protos) == set(_lib._SIGS)
    for name, (_, argtypes) in _lib._SIGS.items():
        want = [klass(p) for p in protos[name]]
        got = [klass_ct(t) for t in argtypes]
        assert got == want, (name, got, want)
