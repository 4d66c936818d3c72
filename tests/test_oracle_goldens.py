"""not gpu: pin the CPU oracle (and the host crop planner) to golden vectors.

  ref_*.json/npz  were produced by executing the REFERENCE's own functions (tests/golden/gen_reference_goldens.py);
  hf_goldens.npz  by the installed transformers modules the reference calls as third-party code
                  (tests/golden/gen_hf_goldens.py).
"""
import json
from dataclasses import replace
from pathlib import Path

import numpy as np
import pytest
import torch

from leopard_b200 import preprocess as pp
from leopard_b200.config import leopard_idefics2_decoder, leopard_llava, reduced
from leopard_b200.native_weights import rope_inv_freq
from leopard_b200.weights import SyntheticCheckpoint, synth_inputs
from oracle import model as om

G = Path(__file__).resolve().parent / "golden"


def _tup(x):
    return None if x is None else tuple(x)


def test_crop_planner_matches_reference_goldens():
    ref = json.loads((G / "ref_crops.json").read_text())
    for row in ref["plans"]:
        sizes = [tuple(s) for s in row["sizes"]]
        if row["budget"] > 0:
            alloc = pp.allocate_patches(sizes, patch_budget=row["budget"])
            assert alloc == row["alloc"], sizes
            best = [pp.select_best_resolution(s, n) for s, n in zip(sizes, alloc)]
            assert best == [_tup(b) for b in row["best"]], sizes
            exp = [1 + (0 if b is None else (b[0] // 364) * (b[1] // 364)) for b in row["best"]]
        else:
            exp = [1] * len(sizes)
        assert pp.plan_crops(sizes) == exp
    for c in ref["best_sweep"]:
        assert pp.select_best_resolution(tuple(c["size"]), c["n"]) == _tup(c["best"]), c
    for c in ref["alloc_cases"]:
        assert pp.allocate_patches([tuple(s) for s in c["sizes"]], patch_budget=c["budget"]) == c["alloc"]


def test_baseline_config_crop_counts():
    # the allocator gives 17 / 4x7 / 8x5 ViT inputs for configs 2 / 3 / 4, 1 for config 1
    assert pp.plan_crops([(384, 384)]) == [1]
    assert pp.plan_crops([(1344, 1344)]) == [17]
    assert pp.plan_crops([(896, 1152)] * 4) == [7, 7, 7, 7]
    assert pp.plan_crops([(980, 980)] * 8) == [5] * 8


def test_resize_pad_divide_matches_reference_goldens():
    from PIL import Image
    ref = json.loads((G / "ref_crops.json").read_text())
    rng = np.random.RandomState(0)
    for c in ref["tiles"]:
        w, h = c["size"]
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
        res = pp.select_best_resolution(img.size, c["n"])
        assert res == _tup(c["res"])
        padded = pp.resize_and_pad_image(img, res)
        tiles = pp.divide_to_patches(padded, 364)
        assert len(tiles) == c["n_tiles"]
        assert int(np.asarray(padded, dtype=np.int64).sum()) == c["padded_sum"]
        assert [int(np.asarray(t, dtype=np.int64).sum()) for t in tiles] == c["tile_sums"]


def test_pixel_shuffle_and_projector_match_reference():
    z = np.load(G / "ref_shuffle_proj.npz")
    b, t, e = 2, 676, 3
    x = (torch.arange(b)[:, None, None] * 1_000_000 + torch.arange(t)[None, :, None] * 10 + torch.arange(e)[None, None, :]).float()
    assert np.array_equal(om.pixel_shuffle(x).numpy().astype(np.int64), z["shuffle_out"])

    class CK:
        def f32(self, n):
            return torch.from_numpy({"multi_modal_projector.linear_1.weight": z["w1"],
                                     "multi_modal_projector.linear_1.bias": z["b1"],
                                     "multi_modal_projector.linear_2.weight": z["w2"],
                                     "multi_modal_projector.linear_2.bias": z["b2"]}[n])
    cfg = leopard_llava()
    y = om.projector(CK(), cfg, torch.from_numpy(z["proj_in"]))
    assert torch.allclose(y, torch.from_numpy(z["proj_out"]), rtol=1e-5, atol=1e-5)


def test_rope_matches_megatron_reference():
    z = np.load(G / "ref_rope.npz")
    ll, mi = leopard_llava().dec, leopard_idefics2_decoder()
    for fn in (om.rope_inv_freq, rope_inv_freq):  # oracle AND the product's host-side table
        assert np.allclose(fn(ll).numpy(), z["inv_llama31"], rtol=1e-6, atol=0)
        assert np.allclose(fn(mi).numpy(), z["inv_mistral"], rtol=1e-6, atol=0)
        assert np.allclose(fn(replace(ll, rope_scaling=None)).numpy(), z["inv_base"], rtol=1e-6, atol=0)
    t = torch.from_numpy(z["t"])  # [s, b, h, d]
    cos, sin = om.rope_cos_sin(ll, torch.arange(t.shape[0]))
    out = om.apply_rope(t[:, 0], cos, sin)
    assert torch.allclose(out, torch.from_numpy(z["out"])[:, 0], rtol=1e-5, atol=1e-5)


@pytest.fixture(scope="module")
def hf():
    z = np.load(G / "hf_goldens.npz")
    vl, dl, vocab, S = [int(v) for v in z["meta"]]
    cfg = reduced(leopard_llava(), vl, dl, vocab)
    return z, cfg, SyntheticCheckpoint(cfg, seed=1234), S


def test_oracle_vit_matches_transformers_siglip(hf):
    z, cfg, ck, _ = hf
    pix, _ = synth_inputs(cfg, n_crops=2, n_text=8, seed=1)
    out = om.vit_forward(ck, cfg, pix.to(torch.bfloat16).float())
    got = out[:, ::13, ::9].numpy()
    assert np.abs(got - z["vit_out"]).max() < 2e-4 * float(z["vit_absmax"])


def test_oracle_decoder_matches_transformers_llama_and_mistral(hf):
    z, cfg, ck, S = hf
    g = torch.Generator().manual_seed(6)
    emb = (torch.randn(S, cfg.dec.hidden, generator=g) * 0.5).to(torch.bfloat16).float()
    pos = torch.from_numpy(z["dec_pos"])
    logits = om.decoder_forward(ck, cfg, emb, pos)
    assert np.abs(logits[:, ::3].numpy() - z["dec_logits"]).max() < 2e-4 * np.abs(z["dec_logits"]).max()
    cfg_m = replace(cfg, dec=replace(leopard_idefics2_decoder(), layers=cfg.dec.layers, vocab=cfg.dec.vocab))
    logits_m = om.decoder_forward(SyntheticCheckpoint(cfg_m, seed=1234), cfg_m, emb, pos)
    assert np.abs(logits_m[:, ::3].numpy() - z["dec_logits_mistral"]).max() < 2e-4 * np.abs(z["dec_logits_mistral"]).max()
    assert np.abs(logits_m[:, ::3].numpy() - z["dec_logits"]).max() > 1e-2  # the two RoPE parameterisations differ


def test_merge_conventions_agree():
    cfg = reduced(leopard_llava(), 1, 1, 2048)
    ck = SyntheticCheckpoint(cfg)
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(2, 169, 4096, generator=g)
    ids = torch.tensor([[5, 6, cfg.image_token_index, 7, cfg.image_token_index, 8]])
    emb, pos = om.merge_hf_expand(ck, cfg, ids, feats)
    assert emb.shape == (1, 4 + 2 * 169, 4096) and torch.equal(pos[0], torch.arange(342))
    # the same sequence written in the Megatron convention (169 x -200 per crop) must give the same rows
    meg = torch.cat([torch.tensor([5, 6]), torch.full((169,), -200), torch.tensor([7]), torch.full((169,), -200),
                     torch.tensor([8])])[None]
    assert torch.equal(om.merge_megatron_scatter(ck, cfg, meg, feats), emb)
    with pytest.raises(ValueError, match="number of image tokens"):
        om.merge_hf_expand(ck, cfg, ids, feats[:1])
    with pytest.raises(AssertionError, match="doesn't match"):
        om.merge_megatron_scatter(ck, cfg, meg, feats[:1])


def test_oracle_idefics2_matches_transformers():
    """NaViT tower with padding masks + dropped all-zero image, modality projection, perceiver, merger, Mistral."""
    import sys
    sys.path.insert(0, str(G))
    import gen_idefics2_goldens as gg
    from leopard_b200.config import leopard_idefics2, reduced_idefics2
    from oracle import idefics2 as oi
    z = np.load(G / "idefics2_goldens.npz")
    vl, pd, dl, vocab = [int(v) for v in z["meta"]]
    cfg = reduced_idefics2(leopard_idefics2(), vl, pd, dl, vocab)
    ck = SyntheticCheckpoint(cfg, seed=1234)
    ids, pv, pm, n_img = gg.make_inputs(cfg)
    st = {}
    logits = oi.prefill_logits(ck, cfg, ids, pv, pm, stages=st)
    feats = st["image_features"].reshape(-1, cfg.dec.hidden)
    assert feats.shape[0] == n_img * cfg.perceiver.latents
    assert np.abs(feats[:, ::7].numpy() - z["image_features"]).max() < 2e-4 * np.abs(z["image_features"]).max()
    assert np.abs(logits[0][:, ::2].numpy() - z["logits"]).max() < 2e-4 * np.abs(z["logits"]).max()
    # position ids: full image -> bucketised rows/cols of the 70 x 70 table
    mask = oi.patch_mask_from_pixels(pm[0, :2], 14)
    pos = oi.navit_position_ids(mask, 70)
    assert pos[0, 0] == 0 and pos[0, 1] == 70 // 13 * 0 + int(torch.bucketize(torch.tensor(1 / 13), torch.arange(1 / 70, 1.0, 1 / 70), right=True))
    assert int(mask[1].sum()) == 70 and bool((pos[1][~mask[1].reshape(-1)] == 0).all())
