"""The SOURCE of the prefill path's HBM-bound glue kernels (leopard_b200/csrc/lp_elem.cu: norms, RoPE, KV append, embedding merge,
im2col, pixel shuffle, cross entropy) executed on the CPU under tests/emu/cuda_emu.h, against the oracle / op contracts.  The
kernels are parity-green on the B200 (tests/test_gpu_ops.py); this is their CPU regression check."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import model as OM
from oracle import native_ops as ref

ROOT = Path(__file__).resolve().parents[1]
BF = torch.bfloat16
LL, I, FL = C.c_longlong, C.c_int, C.c_float


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu_elem")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// numerics helpers"))
    (out / "elem_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    src = (ROOT / "leopard_b200/csrc/lp_elem.cu").read_text()
    body = src[src.index("namespace lp {"):src.index("}  // namespace lp")]
    body = body[:body.index("static int g_norm_variant")] + body[body.index("// cos/sin tables"):]     # drop the host launcher
    assert "<<<" not in body
    (out / "elem_kernels.inc").write_text(body + "\n}  // namespace lp\n")
    lib = out / "libelem.so"
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/elem_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def rnd(*shape, seed=0, scale=1.0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


def err(got, want):
    return float((got.float() - want.float()).abs().max() / (want.float().abs().max() + 1e-12))


def test_norm_kernels(emu):
    for M, D in ((7, 1152), (9, 4096), (3, 64)):
        x = rnd(M, D, seed=1, scale=2.0)
        g = 1 + rnd(D, seed=2, scale=0.1, dtype=torch.float32)
        b = rnd(D, seed=3, scale=0.1, dtype=torch.float32)
        want_ln = ref.layernorm(x, g, b, 1e-6)
        got = torch.empty((M, D), dtype=BF)
        emu.emu_norm(P(x), LL(D), P(g), P(b), P(got), LL(D), I(M), I(D), FL(1e-6), I(0), I(0), I(0))
        assert err(got, want_ln) < 1e-2, ("layernorm", M, D)
        want_rms = ref.rmsnorm(x, g, 1e-5)
        variants = (0, 1, 2) if D == 4096 else (0,)
        outs = []
        for v in variants:
            got = torch.empty((M, D), dtype=BF)
            emu.emu_norm(P(x), LL(D), P(g), None, P(got), LL(D), I(M), I(D), FL(1e-5), I(0), I(1), I(v))
            assert err(got, want_rms) < 1e-2, ("rmsnorm", v, M, D)
            outs.append(got)
        for o in outs[1:]:
            assert torch.equal(o, outs[0])          # the three RMSNorm kernels round identically (documented in the source)
    # ln_post's fused pixel-shuffle store
    x = rnd(2 * 16, 64, seed=4)
    g, b = torch.ones(64), torch.zeros(64)
    got = torch.empty((2 * 4, 4 * 64), dtype=BF)
    emu.emu_norm(P(x), LL(64), P(g), P(b), P(got), LL(256), I(32), I(64), FL(1e-6), I(4), I(0), I(0))
    assert err(got, ref.layernorm(x, g, b, 1e-6, shuffle_side=4)) < 1e-2


def test_rope_and_kv_append(emu):
    S, n_q, n_kv, hd = 11, 4, 2, 128
    pos = torch.arange(3, 3 + S, dtype=torch.int32)
    inv = OM.rope_inv_freq(__import__("leopard_b200.config", fromlist=["x"]).leopard_llava().dec)
    cs, sn = torch.empty((S, hd // 2)), torch.empty((S, hd // 2))
    emu.emu_rope_table(P(pos), P(inv), P(cs), P(sn), I(S), I(hd // 2))
    wc, ws = ref.rope_table(pos, inv)
    assert torch.allclose(cs, wc, atol=1e-6) and torch.allclose(sn, ws, atol=1e-6)
    qkv = rnd(S, (n_q + 2 * n_kv) * hd, seed=1)
    want = ref.rope_inplace(qkv.clone(), 0, n_q + n_kv, hd, cs, sn)
    got = qkv.clone()
    emu.emu_rope(P(got), LL(got.stride(0)), I(S), I(0), I(n_q + n_kv), I(hd), P(cs), P(sn))
    assert torch.equal(got, want)
    got2, kc, vc = qkv.clone(), torch.zeros((S + 5, n_kv * hd), dtype=BF), torch.zeros((S + 5, n_kv * hd), dtype=BF)
    emu.emu_rope_kv(P(got2), LL(got2.stride(0)), I(S), I(n_q), I(n_kv), I(hd), P(cs), P(sn), P(kc), P(vc), LL(n_kv * hd))
    assert torch.equal(got2, want)
    assert torch.equal(kc[:S], want[:, n_q * hd:(n_q + n_kv) * hd]) and torch.equal(vc[:S], qkv[:, (n_q + n_kv) * hd:])
    assert float(kc[S:].abs().max()) == 0.0


def test_embedding_merge_both_conventions(emu):
    D, V = 64, 50
    table = rnd(V, D, seed=1)
    # HF convention: one placeholder id expands to rows_per_slot feature rows
    ids = torch.tensor([4, 7, 40, 9, 40, 2])
    feats = rnd(2 * 3, D, seed=2)
    out = torch.zeros((6 + 2 * 2, D), dtype=BF)
    plan = torch.zeros(2 * 6 + 2, dtype=torch.int32)
    emu.emu_embed_merge(P(ids), I(6), LL(40), I(3), P(table), LL(D), LL(V), P(feats), LL(D), I(6), P(out), LL(D), I(D), P(plan))
    want = torch.cat([table[[4, 7]], feats[:3], table[[9]], feats[3:], table[[2]]])
    assert torch.equal(out, want) and plan[-2:].tolist() == [2, 10]
    # Megatron convention: one -200 per feature row, other negative ids read row 0
    ids = torch.tensor([5, -200, -200, 8, -100, -200])
    feats = rnd(3, D, seed=3)
    out = torch.zeros((6, D), dtype=BF)
    emu.emu_embed_merge(P(ids), I(6), LL(-200), I(1), P(table), LL(D), LL(V), P(feats), LL(D), I(3), P(out), LL(D), I(D), P(plan))
    want = torch.stack([table[5], feats[0], feats[1], table[8], table[0], feats[2]])
    assert torch.equal(out, want)


def test_im2col_pixel_shuffle_cross_entropy(emu):
    pix = rnd(2, 3, 28, 42, seed=1, dtype=torch.float32)
    for t, flag in ((pix, 1), (pix.to(BF), 0)):
        out = torch.full((2 * 2 * 3, 592), 5.0).to(BF)
        emu.emu_im2col(P(t), I(flag), P(out), LL(592), I(2), I(3), I(28), I(42))
        want = torch.zeros((12, 592), dtype=BF)
        p = t.float().view(2, 3, 2, 14, 3, 14).permute(0, 2, 4, 1, 3, 5).reshape(12, 588)
        want[:, :588] = p.to(BF)
        assert torch.equal(out, want)                      # incl. the zeroed K padding 588..591
    x = rnd(2 * 36, 24, seed=2)
    out = torch.empty((2 * 9, 96), dtype=BF)
    emu.emu_pixel_shuffle(P(x), LL(24), P(out), LL(96), I(72), I(24), I(6))
    assert torch.equal(out, OM.pixel_shuffle(x.float().view(2, 36, 24)).reshape(18, 96).to(BF))
    logits = rnd(5, 301, seed=3, scale=3.0, dtype=torch.float32)
    labels = torch.tensor([3, 300, -100, 17, 0])
    loss = torch.empty(5)
    emu.emu_cross_entropy(P(logits), I(1), LL(301), P(labels), P(loss), I(5), I(301))
    want = F.cross_entropy(logits, labels.clamp(min=0), reduction="none")
    want[2] = 0.0
    assert torch.allclose(loss, want, atol=1e-5)
