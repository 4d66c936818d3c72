"""The decode step (leopard_b200/csrc/lp_decode.cu — embed, fused RMSNorm + QKV GEMV + RoPE + cache append, split-K attention merged
by atomic tickets, o-proj / SwiGLU / down GEMVs, final norm + logits, arg-max — AND the host functions lp_decode_step /
lp_decode_greedy_step that sequence them) executed on the CPU: launches rewritten by tests/emu/launch_rewrite.py, runtime names from
tests/emu/host_emu.h, the file's one-line inline-PTX helpers (streaming loads, cp.async, ex2) replaced by plain C++.  Checked
against the decoder oracle over a KV cache the oracle filled.  Parity-green on the B200 (tests/test_gpu_decode.py); this is the
CPU regression check.  Narrow widths keep it to about a second; LEOPARD_EMU_FULL_WIDTH=1 runs the real Llama widths (the
block-cooperative GEMV kernels, ~1 minute)."""
import ctypes as C
import os
import re
import subprocess
import sys
from dataclasses import replace
from pathlib import Path

import pytest
import torch

from leopard_b200 import _lib
from leopard_b200.config import leopard_llava, reduced
from leopard_b200.native_weights import load_hf
from leopard_b200.weights import SyntheticCheckpoint
from oracle import model as OM

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu_dec")
    sys.path.insert(0, str(ROOT / "tests/emu"))
    from launch_rewrite import rewrite_launches
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// numerics helpers"))
    (out / "elem_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    src = (ROOT / "leopard_b200/csrc/lp_decode.cu").read_text()
    for line in ('#include "lp_common.cuh"\n', '#include "lp_host.h"\n', '#include "../../include/leopard_b200.h"\n',
                 "int rope_table_base(int pos_base, const float* inv_freq, float* cs, float* sn, int S, int half, cudaStream_t stream);\n"):
        assert line in src
        src = src.replace(line, "")
    ptx = [  # the file's inline-PTX one-liners -> plain C++
        (r'asm volatile\("ld\.global\.nc\.L1::no_allocate\.v4\.u32.*?: "l"\(p\)\);', "r = *reinterpret_cast<const uint4*>(p);"),
        (r'asm volatile\("prefetch\.global\.L2 \[%0\];" ::"l"\(p\)\);', "(void)p;"),
        (r'asm volatile\("cp\.async\.cg\.shared\.global.*?: "memory"\);',
         "*reinterpret_cast<uint4*>(smem_dst) = *reinterpret_cast<const uint4*>(gmem_src);"),
        (r'asm volatile\("cp\.async\.commit_group;" ::: "memory"\);', ""),
        (r'asm volatile\("cp\.async\.wait_group %0;" ::"n"\(N\) : "memory"\);', ""),
        (r'asm volatile\("ld\.global\.cg\.v4\.u32.*?: "l"\(p\)\);', "r = *reinterpret_cast<const uint4*>(p);"),
        (r'asm\("ex2\.approx\.ftz\.f32 %0, %1;" : "=f"\(y\) : "f"\(x\)\);', "y = exp2f(x);"),
        (r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\(t\)\);', "t = 0;"),
        (r'asm volatile\("ld\.acquire\.gpu\.global\.u32 %0, \[%1\];" : "=r"\(v\) : "l"\(p\) : "memory"\);',
         "v = std::atomic_ref<unsigned>(*const_cast<unsigned*>(p)).load();"),
    ]
    for pat, rep in ptx:
        src, n = re.subn(pat, rep, src, flags=re.S)
        assert n >= 1, pat
    for decl in ("  extern __shared__ uint4 xs[];", "  extern __shared__ __align__(16) uint8_t da_smem[];", "  extern __shared__ uint4 xs3[];"):
        assert decl in src, decl            # -> the static buffers of tests/emu/decode_emu.cpp
        src = src.replace(decl, "")
    assert "asm" not in src
    src = rewrite_launches(src)
    assert "<<<" not in src
    (out / "decode_src.inc").write_text(src)
    lib = out / "libdec.so"
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/decode_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(str(lib))
    L.lp_decode_workspace_bytes.restype = C.c_size_t
    L.lp_decode_workspace_bytes.argtypes = [C.POINTER(_lib.DecDims)]
    return L


def _setup(full_width: bool):
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=2 if not full_width else 1, vocab=256)
    if not full_width:
        cfg = replace(cfg, dec=replace(cfg.dec, hidden=512, heads=4, kv_heads=1, ffn=512), vit=replace(cfg.vit, hidden=288, heads=4, ffn=512))
    ck = SyntheticCheckpoint(cfg, seed=3)
    return cfg, ck, load_hf(cfg, lambda n: ck[n], device="cpu")


@pytest.mark.parametrize("full_width", [False] + ([True] if os.environ.get("LEOPARD_EMU_FULL_WIDTH") == "1" else []))
def test_decode_steps_under_emulation(emu, full_width):
    cfg, ck, nw = _setup(full_width)
    d = cfg.dec
    S, cap = 45, 64                                              # 45 cached rows: two 32-key tiles in the first split
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, d.vocab, (S + 2,), generator=g)
    emb = ck.f32(OM.LM + "embed_tokens.weight")[ids]
    kv = []
    ref = OM.decoder_forward(ck, cfg, emb, torch.arange(S + 2), kv_out=kv)       # logits of every position, K/V of every layer
    kv_w = d.kv_heads * d.head_dim
    cache = torch.zeros((d.layers, 2, cap, kv_w), dtype=torch.bfloat16)
    for l, (k, v) in enumerate(kv):
        cache[l, 0, :S] = k[:S].reshape(S, kv_w).to(torch.bfloat16)
        cache[l, 1, :S] = v[:S].reshape(S, kv_w).to(torch.bfloat16)
    logits = torch.full((nw.vocab_padded,), float("nan"))
    ws = torch.empty(emu.lp_decode_workspace_bytes(C.byref(nw.dec_dims)), dtype=torch.uint8)
    args = (C.byref(nw.dec_dims), C.byref(nw.dec), C.c_void_p(nw.embed.data_ptr()), C.c_longlong(nw.embed.stride(0)))

    def rel(a, b):
        return float((a.float() - b.float()).abs().max() / b.float().abs().max())

    # host-positioned step: token S at position S
    tok = ids[S:S + 1].clone()
    rc = emu.lp_decode_step(*args, C.c_void_p(tok.data_ptr()), S, C.c_void_p(cache.data_ptr()), C.c_longlong(cap),
                            C.c_void_p(logits.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), None)
    assert rc == 0
    def same_choice(got, want):          # arg-max up to near-ties: random weights give flat logits, bf16 noise can swap neighbours
        return float(got[int(want.argmax())]) >= float(got.max()) - 0.03 * float(want.abs().max())

    assert rel(logits[: d.vocab], ref[S]) < 3e-2 and same_choice(logits[: d.vocab], ref[S])
    for l, (k, v) in enumerate(kv):                              # the step appended this token's K (after RoPE) and V
        assert rel(cache[l, 0, S], k[S].reshape(-1)) < 2e-2 and rel(cache[l, 1, S], v[S].reshape(-1)) < 2e-2
        assert float(cache[l, :, S + 1:].abs().max()) == 0.0
    # device-resident greedy step: token S + 1 at position S + 1 -> arg-max written back, position advanced
    tok_io, pos_io = ids[S + 1:S + 2].clone(), torch.tensor([S + 1], dtype=torch.int32)
    rc = emu.lp_decode_greedy_step(*args, C.c_void_p(tok_io.data_ptr()), C.c_void_p(pos_io.data_ptr()), C.c_void_p(cache.data_ptr()),
                                   C.c_longlong(cap), C.c_void_p(logits.data_ptr()), d.vocab, C.c_void_p(ws.data_ptr()),
                                   C.c_size_t(ws.numel()), None)
    assert rc == 0
    assert rel(logits[: d.vocab], ref[S + 1]) < 3e-2
    assert int(tok_io) == int(logits[: d.vocab].argmax()) and same_choice(logits[: d.vocab], ref[S + 1]) and int(pos_io) == S + 2
