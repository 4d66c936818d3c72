"""CPU: bench.py's `library_gpu` competitor (benchmarks/library_gpu.py: F.linear + SDPA on the product's kernel-layout weight
tensors) computes the same function as the oracle — so its timing on the B200 is a fair library baseline and its logits
are an independent check of the weight re-layout (head padding to 80, fused qkv, 32-row gate/up granules)."""
import torch

from benchmarks.library_gpu import LibraryPrefill
from leopard_b200.config import leopard_llava, reduced
from leopard_b200.native_weights import load_hf
from leopard_b200.weights import SyntheticCheckpoint, synth_inputs
from oracle import model as om


def test_library_leg_matches_oracle_fp32_reduced_depth():
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=512)
    ck = SyntheticCheckpoint(cfg, seed=3)
    nw = load_hf(cfg, lambda n: ck[n], "cpu")
    lib = LibraryPrefill(cfg, nw.named, dtype=torch.float32)
    pix, ids = synth_inputs(cfg, n_crops=2, n_text=10)
    got = lib.prefill(ids[0], pix.to(torch.bfloat16).float(), last_only=False)
    want = om.prefill_logits(ck, cfg, ids, pix.to(torch.bfloat16).float())[0]
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 2e-4 * float(want.abs().max())
