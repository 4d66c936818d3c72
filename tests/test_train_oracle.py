"""CPU tests of oracle/train.py: the closed-form backward of every block against torch autograd, the optimizer arithmetic
against torch.optim / torch.nn.utils, and the end-to-end SFT step (loss goes down on a repeated batch)."""
import math

import torch
import torch.nn.functional as F

from leopard_b200.config import leopard_llava, reduced
from leopard_b200.weights import SyntheticCheckpoint
from oracle import model as OM
from oracle import train as OT


def _leaf(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g, dtype=torch.float64) * scale).requires_grad_(True)


def _close(a, b, tol=1e-9):
    assert torch.allclose(a, b, rtol=tol, atol=tol), float((a - b).abs().max())


def test_norm_and_activation_backward_match_autograd():
    x, w, dy = _leaf(7, 48, seed=1), _leaf(48, seed=2), _leaf(7, 48, seed=3).detach()
    OM.rmsnorm(x, w, 1e-5).backward(dy)
    dx, dw = OT.rmsnorm_bwd(x.detach(), w.detach(), 1e-5, dy)
    _close(dx, x.grad), _close(dw, w.grad)

    x, w, b = _leaf(5, 40, seed=4), _leaf(40, seed=5), _leaf(40, seed=6)
    F.layer_norm(x, (40,), w, b, 1e-6).backward(dy[:5, :40])
    dx, dw, db = OT.layernorm_bwd(x.detach(), w.detach(), 1e-6, dy[:5, :40])
    _close(dx, x.grad), _close(dw, w.grad), _close(db, b.grad)

    g, u = _leaf(6, 32, seed=7), _leaf(6, 32, seed=8)
    dm = _leaf(6, 32, seed=9).detach()
    (F.silu(g) * u).backward(dm)
    dg, du = OT.swiglu_bwd(g.detach(), u.detach(), dm)
    _close(dg, g.grad), _close(du, u.grad)

    for approx in ("none", "tanh"):
        x = _leaf(6, 32, seed=10, scale=2.0)
        F.gelu(x, approximate=approx).backward(dm)
        _close(OT.gelu_bwd(x.detach(), dm, approx), x.grad)


def test_linear_and_rope_backward_match_autograd():
    x, w, b = _leaf(9, 24, seed=1), _leaf(16, 24, seed=2), _leaf(16, seed=3)
    dy = _leaf(9, 16, seed=4).detach()
    F.linear(x, w, b).backward(dy)
    dx, dw, db = OT.linear_bwd(x.detach(), w.detach(), dy, has_bias=True)
    _close(dx, x.grad), _close(dw, w.grad), _close(db, b.grad)

    cfg = leopard_llava()
    cos, sin = OM.rope_cos_sin(cfg.dec, torch.arange(11))
    t = _leaf(11, 3, 128, seed=5)
    dt = _leaf(11, 3, 128, seed=6).detach()
    OM.apply_rope(t, cos.double(), sin.double()).backward(dt)
    _close(OT.rope_bwd(dt, cos.double(), sin.double()), t.grad)


def test_attention_backward_matches_autograd_gqa_causal_and_full():
    for causal, (S, H, Hkv, D) in ((True, (37, 8, 2, 16)), (False, (21, 4, 4, 24))):
        q, k, v = _leaf(S, H, D, seed=1), _leaf(S, Hkv, D, seed=2), _leaf(S, Hkv, D, seed=3)
        do = _leaf(S, H, D, seed=4).detach()
        scale = D ** -0.5
        o, lse = OT.attention_fwd(q, k, v, causal, scale)
        o.backward(do)
        # forward agrees with the prefill oracle's softmax(QK^T)V
        rep = H // Hkv
        kk = k.detach().repeat_interleave(rep, 1).transpose(0, 1)
        vv = v.detach().repeat_interleave(rep, 1).transpose(0, 1)
        sc = (q.detach().transpose(0, 1) @ kk.transpose(-1, -2)) * scale
        if causal:
            sc = sc.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
        _close(o.detach(), (torch.softmax(sc, -1) @ vv).transpose(0, 1))
        dq, dk, dv = OT.attention_bwd(q.detach(), k.detach(), v.detach(), o.detach(), lse.detach(), do, causal, scale)
        _close(dq, q.grad), _close(dk, k.grad), _close(dv, v.grad)


def test_masked_lm_loss_and_its_gradient():
    g = torch.Generator().manual_seed(3)
    logits = torch.randn((2, 12, 29), generator=g, dtype=torch.float64).requires_grad_(True)
    labels = torch.randint(0, 29, (2, 12), generator=g)
    w = (torch.rand((2, 12), generator=g) < 0.5).double()
    w[:, -1] = 1
    labels[0, 3] = -100
    w[0, 3] = 0
    # the reference's arithmetic (shift, CE, weighted mean) spelled with torch.nn.functional
    ce = F.cross_entropy(logits[:, :-1].reshape(-1, 29), labels[:, 1:].reshape(-1).clamp(min=0), reduction="none")
    ref = (ce * w[:, 1:].reshape(-1)).sum() / w[:, 1:].sum()
    ref.backward()
    loss, d = OT.masked_lm_loss(logits.detach(), labels, w)
    assert abs(float(loss) - float(ref.detach())) < 1e-6
    assert torch.allclose(d.double(), logits.grad, atol=1e-6)
    assert float(d[:, -1].abs().max()) == 0.0


def test_adamw_and_clip_match_torch():
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(257, generator=g)
    grads = [torch.randn(257, generator=g) * 3 for _ in range(4)]
    for wd in (0.0, 0.1):
        p_ref = p0.clone().requires_grad_(True)
        opt = torch.optim.AdamW([p_ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
        p, m, v = p0.clone(), torch.zeros(257), torch.zeros(257)
        for step, gr in enumerate(grads, 1):
            p_ref.grad = gr.clone()
            opt.step()
            OT.adamw_update(p, gr, m, v, step, 1e-2, 0.9, 0.999, 1e-8, wd)
        assert torch.allclose(p, p_ref.detach(), rtol=2e-5, atol=2e-6), float((p - p_ref.detach()).abs().max())

    ps = [torch.nn.Parameter(torch.zeros(5, 7)), torch.nn.Parameter(torch.zeros(11))]
    for q, s in zip(ps, (1, 2)):
        q.grad = torch.randn(q.shape, generator=torch.Generator().manual_seed(s)) * 4
    before = [q.grad.clone() for q in ps]
    norm = OT.global_grad_norm(before)
    ref_norm = float(torch.nn.utils.clip_grad_norm_(ps, 1.0))
    assert abs(norm - ref_norm) < 1e-5 * ref_norm
    c = OT.clip_coefficient(norm, 1.0)
    for q, b in zip(ps, before):
        assert torch.allclose(q.grad, b * c, rtol=1e-6, atol=1e-7)
    assert OT.clip_coefficient(0.5, 1.0) == 1.0

    assert OT.cosine_lr(5, 1e-3, 1e-4, 10, 100) == 1e-3 * 5 / 10
    assert abs(OT.cosine_lr(55, 1e-3, 1e-4, 10, 100) - (1e-4 + 9e-4 * 0.5 * (math.cos(math.pi * 0.5) + 1))) < 1e-12
    assert OT.cosine_lr(101, 1e-3, 1e-4, 10, 100) == 1e-4


def _tiny_batch(cfg, n_crops, n_text, seed=0):
    g = torch.Generator().manual_seed(seed)
    rows = n_crops * cfg.tokens_per_crop
    text = torch.randint(0, cfg.dec.vocab, (n_text,), generator=g)
    tokens = torch.cat([text[:3], torch.full((rows,), cfg.megatron_image_index), text[3:]])[None]
    labels = tokens.clone()
    labels[tokens < 0] = -100
    weights = torch.zeros(tokens.shape)
    weights[0, -(n_text - 3) // 2:] = 1.0          # "answer" tokens only
    images = torch.rand((n_crops, 3, 364, 364), generator=g) * 2 - 1
    return tokens, labels, weights, images


def test_sft_step_end_to_end_reduces_loss_and_respects_freezing():
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=320)
    frozen_vit = lambda name: not name.startswith("vision_tower.")   # --freeze-clip-vision-tower
    ck = OT.TrainableCheckpoint(SyntheticCheckpoint(cfg, seed=7), trainable=frozen_vit)
    tokens, labels, weights, images = _tiny_batch(cfg, n_crops=1, n_text=13)
    state = {}
    losses = []
    for _ in range(3):
        loss, norm = OT.sft_step(ck, cfg, state, tokens, labels, weights, images, lr=1e-5, clip_grad=1.0)
        losses.append(loss)
        assert norm > 0 and math.isfinite(norm)
    assert losses[2] < losses[0], losses
    grads = ck.grads()
    assert not any(k.startswith("vision_tower.") for k in grads)
    assert "multi_modal_projector.linear_1.weight" in grads and "language_model.lm_head.weight" in grads
    # text rows of the embedding table receive gradient, rows that never occur do not
    ge = grads["language_model.model.embed_tokens.weight"]
    used = torch.zeros(cfg.dec.vocab, dtype=torch.bool)
    used[tokens[tokens >= 0]] = True
    assert float(ge[~used].abs().max()) == 0.0 and float(ge[used].abs().max()) > 0.0
    assert state["step"] == 3
