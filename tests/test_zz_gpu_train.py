"""-m gpu: the SFT step's kernels (csrc/lp_train.cu) and the whole step (train.py) against the oracle.

First run on a B200 in round 2 (gpurun_out/r02a_train_tests.log -> profiles/r02_train_tests.log): all green except the
trainable-tower step on the materialised route (T = 676 is not a multiple of 8: fixed with zero-padded key / value
operands); the flash-style backward is the default route since then.  Hard passes: no xfail markers.
Checker: oracle/native_ops.py per kernel (same signatures), oracle/train.py (autograd) for the whole step."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from leopard_b200 import ops  # noqa: E402
from leopard_b200.config import leopard_llava, reduced  # noqa: E402
from leopard_b200.native_weights import load_hf  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint  # noqa: E402
from oracle import native_ops as ref_ops  # noqa: E402
from oracle import train as OT  # noqa: E402

DEV = "cuda:0"


def _rand(*shape, seed=0, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


def _err(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return float((got - want).abs().max() / (want.abs().max() + 1e-12))


def test_transpose_including_ragged_edges_and_strided_views(native_lib):
    for rows, cols in ((64, 64), (130, 72), (7, 250), (4096, 128)):
        x = _rand(rows, cols, seed=rows)
        assert torch.equal(ops.transpose(x.to(DEV)).cpu(), x.t())
    big = _rand(96, 512, seed=1).to(DEV)
    view = big[:, 128:256]                                   # a head's column slice
    out = torch.zeros((128, 4 * 96), dtype=torch.bfloat16, device=DEV)
    ops.transpose(view, out=out[:, 96:192])
    assert torch.equal(out[:, 96:192].cpu(), view.cpu().t()) and float(out[:, :96].abs().max()) == 0.0


def test_rmsnorm_backward(native_lib):
    for M, D in ((37, 4096), (300, 1152), (5, 8192), (1000, 64)):
        x, dy, add = _rand(M, D, seed=1), _rand(M, D, seed=2), _rand(M, D, seed=3)
        gamma = 1 + _rand(D, seed=4, scale=0.1, dtype=torch.float32)
        want_dg, got_dg = torch.empty(D), torch.empty(D, device=DEV)
        want = ref_ops.rmsnorm_bwd(x, gamma, dy, 1e-5, want_dg, add=add)
        got = ops.rmsnorm_bwd(x.to(DEV), gamma.to(DEV), dy.to(DEV), 1e-5, got_dg, add=add.to(DEV))
        assert _err(got, want) < 1e-2 and _err(got_dg, want_dg) < 2e-3, (M, D)
        got = ops.rmsnorm_bwd(x.to(DEV), gamma.to(DEV), dy.to(DEV), 1e-5, got_dg)
        assert _err(got, ref_ops.rmsnorm_bwd(x, gamma, dy, 1e-5, want_dg)) < 1e-2


def test_swiglu_and_gelu_backward(native_lib):
    gu, dm = _rand(50, 2 * 14336, seed=1, scale=2.0), _rand(50, 14336, seed=2)
    assert _err(ops.swiglu_fwd(gu.to(DEV)), ref_ops.swiglu_fwd(gu)) < 1e-2
    assert _err(ops.swiglu_bwd(gu.to(DEV), dm.to(DEV)), ref_ops.swiglu_bwd(gu, dm)) < 1e-2
    # the materialised forward agrees with the GEMM's fused epilogue on the same pre-activations
    h, w = _rand(64, 256, seed=3), _rand(512, 256, seed=4, scale=0.1)
    fused = ops.gemm(h.to(DEV), w.to(DEV), act=ops.ACT_SWIGLU)
    assert _err(ops.swiglu_fwd(ops.gemm(h.to(DEV), w.to(DEV))), fused) < 1e-2
    z, dy = _rand(33, 4096, seed=5, scale=2.0), _rand(33, 4096, seed=6)
    for act in (ops.ACT_GELU_TANH, ops.ACT_GELU_ERF):
        assert _err(ops.gelu_bwd(z.to(DEV), dy.to(DEV), act), ref_ops.gelu_bwd(z, dy, act)) < 1e-2


def test_colsum_gather_and_embedding_scatter(native_lib):
    dy = _rand(3380, 4096, seed=1)
    got = ops.colsum(dy.to(DEV), torch.empty(4096, device=DEV))
    assert _err(got, dy.float().sum(0)) < 1e-4
    idx = torch.tensor([5, 0, 5, 3379], dtype=torch.int32)
    assert torch.equal(ops.gather_rows(dy.to(DEV), idx.to(DEV)).cpu(), dy[idx.long()])
    ids = torch.tensor([3, -200, 3, 7, -100, 0, -200, 12345])      # 12345 >= 12 rows: clamped like the forward gather
    d = _rand(8, 256, seed=2)
    want, got = torch.zeros(12, 256), torch.zeros(12, 256, device=DEV)
    ref_ops.embed_bwd(d, ids, -200, want)
    ops.embed_bwd(d.to(DEV), ids.to(DEV), -200, got)
    assert _err(got, want) < 1e-6 and float(got[1].abs().max()) == 0.0


def test_masked_loss_forward_backward(native_lib):
    R, V, Vp = 40, 1003, 1008
    g = torch.Generator().manual_seed(0)
    target = torch.randint(0, V, (R,), generator=g)
    target[3], target[4] = -100, -1
    weight = (torch.rand(R, generator=g) < 0.5).float()
    weight[3] = 0.25                                           # weighted row with an out-of-range label
    weight = weight / weight.sum()
    for dtype in (torch.float32, torch.bfloat16):
        logits = _rand(R, Vp, seed=1, scale=3.0, dtype=dtype)
        want_loss, want_d = ref_ops.ce_fwd_bwd(logits, target, weight, V)
        loss, d = ops.ce_fwd_bwd(logits.to(DEV), target.to(DEV), weight.to(DEV), V)
        assert _err(loss, want_loss) < 1e-4 and _err(d, want_d) < 1e-2
        assert float(d[:, V:].abs().max()) == 0.0 and float(d[weight.to(DEV) == 0].abs().max()) == 0.0


def test_softmax_backward_causal_and_full(native_lib):
    for S, causal in ((200, True), (77, False), (1024, True)):
        sc = _rand(S, S, seed=1, scale=4.0, dtype=torch.float32)
        dp = _rand(S, S, seed=2, dtype=torch.float32)
        wp, wd = ref_ops.softmax_bwd(sc, dp, causal, 0.088)
        gp, gd = ops.softmax_bwd(sc.to(DEV), dp.to(DEV), causal, 0.088)
        assert _err(gp, wp) < 1e-2 and _err(gd, wd) < 1e-2
        if causal:
            assert float(gp.cpu().triu(1).abs().max()) == 0.0 and float(gd.cpu().triu(1).abs().max()) == 0.0


def test_grad_norm_clip_and_adamw(native_lib):
    g1, g2 = _rand(1_000_003, seed=1, dtype=torch.float32), _rand(4096, seed=2, scale=3.0, dtype=torch.float32)
    out = ops.grad_norm([g1.to(DEV), g2.to(DEV)], 1.0).cpu()
    norm = OT.global_grad_norm([g1, g2])
    assert abs(float(out[0]) - norm) < 1e-4 * norm and abs(float(out[1]) - OT.clip_coefficient(norm, 1.0)) < 1e-6
    assert float(ops.grad_norm([g2.to(DEV) * 1e-4], 1.0).cpu()[1]) == 1.0
    p = _rand(10_000, seed=3, dtype=torch.float32)
    pr, m, v = p.clone(), torch.zeros_like(p), torch.zeros_like(p)
    pd, md, vd = p.to(DEV), torch.zeros(10_000, device=DEV), torch.zeros(10_000, device=DEV)
    p16 = torch.empty(10_000, dtype=torch.bfloat16, device=DEV)
    clip = torch.tensor([5.0, 0.2], device=DEV)
    for step in (1, 2, 3):
        g = _rand(10_000, seed=10 + step, dtype=torch.float32)
        OT.adamw_update(pr, g * 0.2, m, v, step, 1e-3, 0.9, 0.999, 1e-8, 0.01)
        ops.adamw_step(pd, g.to(DEV), md, vd, p16, step, 1e-3, 0.9, 0.999, 1e-8, 0.01, clip)
    assert torch.allclose(pd.cpu(), pr, rtol=1e-5, atol=1e-6)
    assert torch.equal(p16.cpu(), pd.cpu().to(torch.bfloat16))


def test_whole_sft_step_against_autograd_oracle(native_lib):
    from leopard_b200.engine import PrefillEngine
    from leopard_b200.train import SFTTrainer
    from test_train_glue import _batch, _native_layout
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=2, vocab=324)
    ck = SyntheticCheckpoint(cfg, seed=11)
    eng = PrefillEngine(cfg, load_hf(cfg, lambda n: ck[n], DEV))
    tokens, labels, weights, images = _batch(cfg, b=2, n_text=15, seed=3)
    ock = OT.TrainableCheckpoint(ck, trainable=lambda n: not n.startswith("vision_tower."))
    ref_loss, ref = OT.sft_loss_and_grads(ock, cfg, tokens, labels, weights, images)
    ref_norm = OT.global_grad_norm(ref.values())
    tr = SFTTrainer(eng, lr=1e-5, clip_grad=1.0)
    before = eng.w.named["dec.1.o_w"].clone()
    loss, norm = tr.step(tokens, labels, weights, images.to(DEV))
    assert abs(loss - ref_loss) < 0.03 * abs(ref_loss), (loss, ref_loss)
    assert abs(norm - ref_norm) < 0.05 * ref_norm, (norm, ref_norm)
    bad = {}
    for n in tr.names:
        want, got = _native_layout(cfg, ref, n).float(), tr.grad(n).float().cpu()
        e = float((got - want).norm() / (want.norm() + 1e-12))
        if e > 0.06:
            bad[n] = e
    assert not bad, bad
    assert not torch.equal(eng.w.named["dec.1.o_w"], before)
    # a second step on the same batch lowers the loss
    loss2, _ = tr.step(tokens, labels, weights, images.to(DEV))
    assert loss2 < loss


def test_vision_tower_backward_kernels(native_lib):
    for M, D in ((676, 1152), (50, 4096), (3, 64)):
        x, dy, add = _rand(M, D, seed=1, scale=3.0), _rand(M, D, seed=2), _rand(M, D, seed=3)
        gamma = 1 + _rand(D, seed=4, scale=0.1, dtype=torch.float32)
        wg, wb = torch.empty(D), torch.empty(D)
        gg, gb = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
        want = ref_ops.layernorm_bwd(x, gamma, dy, 1e-6, wg, wb, add=add)
        got = ops.layernorm_bwd(x.to(DEV), gamma.to(DEV), dy.to(DEV), 1e-6, gg, gb, add=add.to(DEV))
        assert _err(got, want) < 1e-2 and _err(gg, wg) < 2e-3 and _err(gb, wb) < 2e-3, (M, D)
    d = _rand(2 * 169, 4 * 1152, seed=5)
    un = ops.pixel_unshuffle(d.to(DEV), 26)
    assert torch.equal(un.cpu(), ref_ops.pixel_unshuffle(d, 26))
    assert torch.equal(ops.pixel_shuffle(un, 26).cpu(), d)                      # inverse of the forward kernel
    dx = _rand(3 * 676, 1152, seed=6)
    got = ops.rowmod_sum(dx.to(DEV), 676, torch.empty((676, 1152), device=DEV))
    assert _err(got, dx.float().view(3, 676, 1152).sum(0)) < 1e-6
    a, b = _rand(1_000_003, seed=7, dtype=torch.float32), _rand(1_000_003, seed=8, dtype=torch.float32)
    assert torch.equal(ops.accumulate(a.to(DEV), b.to(DEV)).cpu(), a + b)


@pytest.mark.parametrize("attn_bwd", ["flash", "materialized"])
def test_whole_sft_step_with_trainable_vision_tower(native_lib, attn_bwd):
    """The reference's default: examples/llava/train_multiimg_llava_siglip.sh passes no --freeze-clip-vision-tower."""
    from leopard_b200.engine import PrefillEngine
    from leopard_b200.native_weights import named_to_hf
    from leopard_b200.train import SFTTrainer
    from test_train_glue import _batch
    cfg = reduced(leopard_llava(), vit_layers=2, dec_layers=1, vocab=320)
    ck = SyntheticCheckpoint(cfg, seed=13)
    eng = PrefillEngine(cfg, load_hf(cfg, lambda n: ck[n], DEV))
    tokens, labels, weights, images = _batch(cfg, b=1, n_text=15, seed=6)
    ref_loss, ref = OT.sft_loss_and_grads(OT.TrainableCheckpoint(ck), cfg, tokens, labels, weights, images)
    tr = SFTTrainer(eng, lr=0.0, freeze_vision_tower=False, attn_bwd=attn_bwd)
    loss, norm = tr.step(tokens, labels, weights, images.to(DEV))
    assert abs(loss - ref_loss) < 0.03 * abs(ref_loss)
    got = named_to_hf(cfg, {n: tr.grad(n) for n in tr.names})
    bad = {}
    for name, want in ref.items():
        if name.endswith("self_attn.k_proj.bias"):       # exactly zero in exact arithmetic (see test_train_glue.py)
            continue
        e = float((got[name].float().cpu() - want).norm() / (want.norm() + 1e-12))
        if e > 0.08:
            bad[name] = e
    assert not bad, bad


# ---- flash-style attention backward (lp_attn_bwd.cu): kept LAST — a wedged tcgen05 pipeline traps and poisons the context
def _attn_case(B, T, heads, kv_heads, seed):
    hd = 128
    q_w, kv_w = heads * hd, kv_heads * hd
    qkv = _rand(B * T, q_w + 2 * kv_w, seed=seed)
    dout = _rand(B * T, q_w, seed=seed + 1)
    return qkv, dout, q_w, kv_w


@pytest.mark.parametrize("B,T,heads,kv_heads,causal", [(1, 128, 4, 1, True), (2, 200, 8, 2, True), (1, 384, 4, 4, False),
                                                       (1, 1024, 8, 2, True)])
def test_flash_attention_backward_against_closed_form(native_lib, B, T, heads, kv_heads, causal):
    qkv, dout, q_w, kv_w = _attn_case(B, T, heads, kv_heads, seed=T)
    out = ops.attention(qkv.to(DEV), B, T, heads, kv_heads, 128, 128, 0, q_w, q_w + kv_w, causal)
    got = ops.attention_bwd(qkv.to(DEV), out, dout.to(DEV), B, T, heads, kv_heads, 128, 0, q_w, q_w + kv_w, causal)
    want = ref_ops.attention_bwd(qkv, out.cpu(), dout, B, T, heads, kv_heads, 128, 0, q_w, q_w + kv_w, causal)
    for name, sl in (("dq", slice(0, q_w)), ("dk", slice(q_w, q_w + kv_w)), ("dv", slice(q_w + kv_w, None))):
        assert _err(got[:, sl], want[:, sl]) < 3e-2, name
    again = ops.attention_bwd(qkv.to(DEV), out, dout.to(DEV), B, T, heads, kv_heads, 128, 0, q_w, q_w + kv_w, causal)
    assert torch.equal(got, again)                                  # deterministic: no atomics


def test_flash_attention_backward_vision_tower_layout(native_lib):
    """(head_dim, head_stride) = (72, 80), full attention over 676 tokens, several crops; zero padding columns in qkv."""
    B, T, heads = 2, 676, 16
    qkv = _rand(B * T, 3 * heads * 80, seed=3).view(B * T, 3 * heads, 80)
    qkv[:, :, 72:] = 0
    qkv = qkv.reshape(B * T, 3 * heads * 80).contiguous()
    dout = _rand(B * T, heads * 72, seed=4)
    out = ops.attention(qkv.to(DEV), B, T, heads, heads, 72, 80, 0, heads * 80, 2 * heads * 80, False)
    got = ops.attention_bwd(qkv.to(DEV), out, dout.to(DEV), B, T, heads, heads, 72, 0, heads * 80, 2 * heads * 80, False,
                            head_stride=80)
    want = ref_ops.attention_bwd(qkv, out.cpu(), dout, B, T, heads, heads, 72, 0, heads * 80, 2 * heads * 80, False,
                                 head_stride=80)
    assert _err(got, want) < 3e-2
    assert float(got.view(B * T, 3 * heads, 80)[:, :, 72:].abs().max()) == 0.0


def test_whole_sft_step_with_flash_backward_matches_materialised(native_lib):
    from leopard_b200.engine import PrefillEngine
    from leopard_b200.train import SFTTrainer
    from test_train_glue import _batch
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=2, vocab=324)
    ck = SyntheticCheckpoint(cfg, seed=11)
    eng = PrefillEngine(cfg, load_hf(cfg, lambda n: ck[n], DEV))
    tokens, labels, weights, images = _batch(cfg, b=2, n_text=15, seed=3)
    grads = {}
    for mode in ("materialized", "flash"):
        tr = SFTTrainer(eng, lr=0.0, attn_bwd=mode)
        tr.step(tokens, labels, weights, images.to(DEV))
        grads[mode] = tr.flat_grad.clone()
    rel = float((grads["flash"] - grads["materialized"]).norm() / grads["materialized"].norm())
    assert rel < 0.03, rel


# ---- 4-CTA-cluster GEMM with A multicast (lp_gemm4.cu, opt-in mode 2): the very last test — a wedged cluster traps
def test_four_cta_multicast_gemm_matches_pair_mode(native_lib):
    """Same K order as the pair kernel, so the results must be bit-identical; shapes large enough for pair mode (>= 74 tiles),
    one with an odd number of column blocks (the second pair idles on the last one), one through the SwiGLU epilogue."""
    from leopard_b200 import _lib
    L = _lib.lib()
    cases = [(4096, 4096, 1024, 0, False), (4096, 3840, 512, 0, True), (2048, 8192, 1024, ops.ACT_SWIGLU, True)]
    try:
        for M, N, K, act, use_res in cases:
            a, w = _rand(M, K, seed=M).to(DEV), (_rand(N, K, seed=N) * K ** -0.5).to(DEV)
            bias = _rand(N, seed=3, dtype=torch.float32).to(DEV)
            n_out = N // 2 if act == ops.ACT_SWIGLU else N
            res = _rand(M, n_out, seed=4).to(DEV) if use_res else None
            L.lp_gemm_set_pair_mode(1)
            want = ops.gemm(a, w, bias=bias, residual=res, act=act, block_n=256)
            L.lp_gemm_set_pair_mode(2)
            got = ops.gemm(a, w, bias=bias, residual=res, act=act, block_n=256)
            torch.cuda.synchronize()
            assert torch.equal(got, want), (M, N, K, act, float((got.float() - want.float()).abs().max()))
    finally:
        L.lp_gemm_set_pair_mode(1)
