"""The SOURCE of the SFT step's element-wise / reduction kernels (leopard_b200/csrc/lp_train.cu) compiled with g++ against a
thread-per-CUDA-thread emulator (tests/emu/cuda_emu.h) and executed on the CPU, against the op contracts of
oracle/native_ops.py.  This checks the kernels' logic (indexing, masks, reductions, layouts, ragged edges) without a GPU;
the same contracts are checked on the device by tests/test_zz_gpu_train.py.  Nothing here is on the product path."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest
import torch

from oracle import native_ops as ref
from oracle import train as OT

ROOT = Path(__file__).resolve().parents[1]
BF = torch.bfloat16


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// numerics helpers"))
    helpers = common[a:common.rindex("}  // namespace lp")]
    train = (ROOT / "leopard_b200/csrc/lp_train.cu").read_text()
    kernels = train[train.index("namespace lp {"):train.index("}  // namespace lp") + len("}  // namespace lp")]
    (out / "kernels.inc").write_text("namespace lp {\n" + helpers + "\n}\n" + kernels + "\n")
    lib = out / "libemu.so"
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/train_kernels_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def rnd(*shape, seed=0, scale=1.0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


def err(got, want):
    return float((got.float() - want.float()).abs().max() / (want.float().abs().max() + 1e-12))


LL, I, F = C.c_longlong, C.c_int, C.c_float


def test_transpose_ragged_and_strided(emu):
    for rows, cols in ((64, 64), (130, 72), (7, 250)):
        x = rnd(rows, cols, seed=rows)
        out = torch.zeros((cols, rows + 2), dtype=BF)
        emu.emu_transpose(P(x), LL(x.stride(0)), I(rows), I(cols), P(out), LL(out.stride(0)))
        assert torch.equal(out[:, :rows], x.t()) and float(out[:, rows:].abs().max()) == 0.0
    big = rnd(40, 96, seed=1)
    view = big[:, 32:64]
    out = torch.zeros((32, 80), dtype=BF)
    emu.emu_transpose(P(view), LL(96), I(40), I(32), P(out[:, 40:]), LL(80))
    assert torch.equal(out[:, 40:], view.t()) and float(out[:, :40].abs().max()) == 0.0


def test_rmsnorm_and_layernorm_backward(emu):
    for M, D, nb in ((9, 4096, 4), (5, 1152, 5), (3, 64, 2), (2, 8192, 1)):
        x, dy, add = rnd(M, D, seed=1, scale=2.0), rnd(M, D, seed=2), rnd(M, D, seed=3)
        gamma = 1 + rnd(D, seed=4, scale=0.1, dtype=torch.float32)
        for use_add in (True, False):
            a = add if use_add else None
            wdg, gdg = torch.empty(D), torch.zeros(D)
            want = ref.rmsnorm_bwd(x, gamma, dy, 1e-5, wdg, add=a)
            got = torch.empty((M, D), dtype=BF)
            emu.emu_rmsnorm_bwd(P(x), LL(D), P(gamma), P(dy), LL(D), P(a), LL(D), P(got), LL(D), P(gdg), I(M), I(D), F(1e-5), I(nb))
            assert err(got, want) < 1e-2 and err(gdg, wdg) < 1e-4, (M, D, use_add)
        if D <= 4096:
            wdg, wdb, gdg, gdb = torch.empty(D), torch.empty(D), torch.zeros(D), torch.zeros(D)
            want = ref.layernorm_bwd(x, gamma, dy, 1e-6, wdg, wdb, add=add)
            got = torch.empty((M, D), dtype=BF)
            emu.emu_layernorm_bwd(P(x), LL(D), P(gamma), P(dy), LL(D), P(add), LL(D), P(got), LL(D), P(gdg), P(gdb), I(M), I(D),
                                  F(1e-6), I(nb))
            assert err(got, want) < 1e-2 and err(gdg, wdg) < 1e-4 and err(gdb, wdb) < 1e-4, (M, D)


def test_swiglu_gelu_colsum(emu):
    M, Fd = 5, 96
    gu, dm = rnd(M, 2 * Fd, seed=1, scale=2.0), rnd(M, Fd, seed=2)
    out = torch.empty((M, Fd), dtype=BF)
    emu.emu_swiglu(I(0), P(gu), LL(2 * Fd), None, LL(0), P(out), LL(Fd), I(M), I(Fd), I(2))
    assert err(out, ref.swiglu_fwd(gu)) < 1e-2
    dgu = torch.empty((M, 2 * Fd), dtype=BF)
    emu.emu_swiglu(I(1), P(gu), LL(2 * Fd), P(dm), LL(Fd), P(dgu), LL(2 * Fd), I(M), I(Fd), I(2))
    assert err(dgu, ref.swiglu_bwd(gu, dm)) < 1e-2
    z, dy = rnd(7, 64, seed=3, scale=2.0), rnd(7, 64, seed=4)
    for act in (1, 2):
        dz = torch.empty((7, 64), dtype=BF)
        emu.emu_gelu_bwd(P(z), LL(64), P(dy), LL(64), P(dz), LL(64), I(7), I(64), I(act), I(1))
        assert err(dz, ref.gelu_bwd(z, dy, act)) < 1e-2, act
    d = rnd(37, 130, seed=5)
    cs = torch.zeros(130)
    emu.emu_colsum(P(d), LL(130), I(37), I(130), P(cs))
    assert err(cs, d.float().sum(0)) < 1e-5


def test_masked_loss_forward_backward(emu):
    R, V, Vp = 6, 301, 304
    g = torch.Generator().manual_seed(0)
    target = torch.randint(0, V, (R,), generator=g)
    target[1], target[2] = -100, -1
    weight = torch.tensor([0.3, 0.2, 0.0, 0.1, 0.0, 0.4])
    for dtype in (torch.float32, BF):
        logits = rnd(R, Vp, seed=1, scale=3.0, dtype=dtype)
        want_loss, want_d = ref.ce_fwd_bwd(logits, target, weight, V)
        loss, d = torch.full((R,), -1.0), torch.full((R, Vp), 7.0).to(BF)
        emu.emu_ce(P(logits), I(1 if dtype == torch.float32 else 0), LL(Vp), P(target), P(weight), P(loss), P(d), LL(Vp), I(R), I(V),
                   I(Vp))
        assert err(loss, want_loss) < 1e-5 and err(d, want_d) < 1e-2
        assert float(d[:, V:].abs().max()) == 0.0 and float(d[weight == 0].abs().max()) == 0.0


def test_embedding_scatter_gather_unshuffle_rowsum(emu):
    ids = torch.tensor([3, -200, 3, 7, -100, 0, -200, 12345])       # 12345: beyond the table -> last row, like the forward
    d = rnd(8, 64, seed=2)
    want, got = torch.zeros(12, 64), torch.zeros(12, 64)
    ref.embed_bwd(d, ids, -200, want)
    emu.emu_embed_bwd(P(d), LL(64), P(ids), I(8), LL(-200), P(got), LL(64), I(64), LL(12))
    assert err(got, want) < 1e-6 and float(got[1].abs().max()) == 0.0
    src = rnd(20, 64, seed=3)
    idx = torch.tensor([5, 0, 5, 19], dtype=torch.int32)
    out = torch.empty((4, 64), dtype=BF)
    emu.emu_gather_rows(P(src), LL(64), P(idx), I(4), I(64), P(out), LL(64))
    assert torch.equal(out, src[idx.long()])
    dsh = rnd(2 * 4, 4 * 16, seed=4)                       # side 4: 16 tokens -> 4 shuffled rows per image
    un = torch.empty((2 * 16, 16), dtype=BF)
    emu.emu_pixel_unshuffle(P(dsh), LL(64), P(un), LL(16), I(32), I(16), I(4))
    assert torch.equal(un, ref.pixel_unshuffle(dsh, 4)) and torch.equal(ref._shuffle(un.float(), 4).to(BF), dsh)
    dx = rnd(3 * 10, 34, seed=5)
    rs = torch.zeros((10, 34))
    emu.emu_rowmod_sum(P(dx), LL(34), I(3), I(10), I(34), P(rs))
    assert err(rs, dx.float().view(3, 10, 34).sum(0)) < 1e-6


def test_softmax_backward(emu):
    for S, causal in ((40, True), (33, False)):
        sc, dp = rnd(S, S, seed=1, scale=4.0, dtype=torch.float32), rnd(S, S, seed=2, dtype=torch.float32)
        wp, wd = ref.softmax_bwd(sc, dp, causal, 0.088)
        ld = -(-S // 8) * 8
        gp, gd = torch.full((S, ld), 9.0).to(BF), torch.full((S, ld), 9.0).to(BF)
        emu.emu_softmax_bwd(P(sc), LL(S), P(dp), LL(S), P(gp), LL(ld), P(gd), LL(ld), I(S), I(S), I(1 if causal else 0), F(0.088))
        assert err(gp[:, :S], wp) < 1e-2 and err(gd[:, :S], wd) < 1e-2
        if causal:
            assert float(gp[:, :S].float().triu(1).abs().max()) == 0.0


def test_grad_norm_adamw_accumulate(emu):
    x = rnd(10_003, seed=1, scale=2.0, dtype=torch.float32)
    out = torch.zeros(2)
    emu.emu_grad_norm(P(x), LL(x.numel()), I(3), F(1.0), P(out))               # the harness feeds the tensor twice
    norm = OT.global_grad_norm([x, x])
    assert abs(float(out[0]) - norm) < 1e-5 * norm and abs(float(out[1]) - OT.clip_coefficient(norm, 1.0)) < 1e-7
    small = x * 1e-4                                                            # (kept alive: P() only takes the address)
    emu.emu_grad_norm(P(small), LL(small.numel()), I(2), F(1.0), P(out))
    assert float(out[1]) == 1.0
    p = rnd(1001, seed=3, dtype=torch.float32)
    pr, m, v = p.clone(), torch.zeros(1001), torch.zeros(1001)
    pd, md, vd, p16 = p.clone(), torch.zeros(1001), torch.zeros(1001), torch.empty(1001, dtype=BF)
    clip = torch.tensor([5.0, 0.2])
    for step in (1, 2, 3):
        g = rnd(1001, seed=10 + step, dtype=torch.float32)
        OT.adamw_update(pr, g * 0.2, m, v, step, 1e-3, 0.9, 0.999, 1e-8, 0.01)
        emu.emu_adamw(P(pd), P(g), P(md), P(vd), P(p16), LL(1001), F(1e-3), F(0.9), F(0.999), F(1e-8), F(0.01),
                      F(1 / (1 - 0.9 ** step)), F(1 / (1 - 0.999 ** step)), P(clip), I(2))
    assert torch.allclose(pd, pr, rtol=1e-5, atol=1e-6) and torch.equal(p16, pd.to(BF))
    # a non-finite gradient norm turns the update into a no-op (found-inf path)
    keep = (pd.clone(), md.clone(), vd.clone(), p16.clone())
    for bad in (float("inf"), float("nan")):
        bad_norm = torch.tensor([bad, 1.0])
        emu.emu_adamw(P(pd), P(g), P(md), P(vd), P(p16), LL(1001), F(1e-3), F(0.9), F(0.999), F(1e-8), F(0.01), F(1.0), F(1.0),
                      P(bad_norm), I(2))
        assert all(torch.equal(x, y) for x, y in zip((pd, md, vd, p16), keep))
    # the 16-byte variant (what lp_adamw_step launches for aligned tensors of 4 k elements) == the scalar kernel, bit for bit
    n4 = 2052
    p0, g0 = rnd(n4, seed=21, dtype=torch.float32), rnd(n4, seed=22, dtype=torch.float32)
    res = []
    for fn in (emu.emu_adamw, emu.emu_adamw_vec4):
        pp, mm, vv, pb = p0.clone(), torch.full((n4,), 0.01), torch.full((n4,), 0.02), torch.empty(n4, dtype=BF)
        for step in (1, 2):
            fn(P(pp), P(g0), P(mm), P(vv), P(pb), LL(n4), F(1e-3), F(0.9), F(0.999), F(1e-8), F(0.01), F(1 / (1 - 0.9 ** step)),
               F(1 / (1 - 0.999 ** step)), P(clip), I(2))
        res.append((pp, mm, vv, pb))
    assert all(torch.equal(x, y) for x, y in zip(*res))
    a, b = rnd(1003, seed=7, dtype=torch.float32), rnd(1003, seed=8, dtype=torch.float32)
    want = a + b
    emu.emu_accumulate(P(a), P(b), LL(1003), I(2))
    assert torch.equal(a, want)
