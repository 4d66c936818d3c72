"""CPU: parity mode's host sequencing and split algebra (leopard_b200/parity.py) with every native op replaced by its contract
restatement (oracle/native_ops.py: the GEMM multiplies the bf16 operands exactly and accumulates in fp32, as the tcgen05 kernel
does) against the fp32 oracle — the three-term operand split, the [W | W | W] weight tiling, the six cross terms of attention,
head padding, GQA indexing and the merge reproduce fp32 arithmetic to ~1e-5, which is what lets the GPU run claim the
north-star 1e-3 tolerance.  The CUDA glue kernels themselves are checked on the GPU (tests/test_gpu_parity_mode.py)."""
import types

import pytest
import torch

from leopard_b200 import parity as PM
from leopard_b200.config import leopard_llava, reduced
from leopard_b200.native_weights import load_hf
from leopard_b200.weights import SyntheticCheckpoint, synth_inputs
from oracle import model as om
from oracle import native_ops


@pytest.fixture
def cpu_parity(monkeypatch):
    from dataclasses import replace
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=640)
    cfg = replace(cfg, dec=replace(cfg.dec, hidden=1024, heads=8, kv_heads=2, ffn=1024), vit=replace(cfg.vit, hidden=288, heads=4, ffn=512))
    ck = SyntheticCheckpoint(cfg, seed=21)
    nw = load_hf(cfg, lambda n: ck[n], device="cpu")
    monkeypatch.setattr(PM, "ops", native_ops)
    monkeypatch.setattr(PM._lib, "lib", lambda: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    eng = types.SimpleNamespace(cfg=cfg, w=nw, device=torch.device("cpu"))
    return cfg, ck, PM.ParityForward(eng)


def test_three_term_split_carries_fp32():
    x = torch.randn(64, 40) * torch.logspace(-3, 3, 40)
    out = torch.zeros(64, 120, dtype=torch.bfloat16)
    native_ops.parity_split(x, out, 0)
    back = out[:, :40].float() + out[:, 40:80].float() + out[:, 80:].float()
    assert float(((back - x).abs() / x.abs().clamp(min=1e-30)).max()) < 2.0 ** -22


def test_parity_forward_reproduces_fp32_oracle(cpu_parity):
    cfg, ck, pf = cpu_parity
    pix, ids = synth_inputs(cfg, n_crops=2, n_text=12)
    pix = pix.to(torch.bfloat16).float()
    stages, ref_stages = {}, {}
    got = pf.prefill_logits(ids, pix, stages=stages)
    ref = om.prefill_logits(ck, cfg, ids, pix, stages=ref_stages)[0]
    assert got.shape == ref.shape
    e_feat = float((stages["image_features"] - ref_stages["image_features"]).abs().max())
    e = float((got - ref).abs().max())
    assert e_feat < 1e-4 and e < 1e-4, (e_feat, e)                  # two orders below the north-star 1e-3
    assert torch.equal(got.argmax(-1), ref.argmax(-1))
