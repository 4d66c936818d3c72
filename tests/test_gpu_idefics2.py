"""-m gpu: Leopard-Idefics2 rows (NaViT tower, modality projection, perceiver, merger, Mistral prefill) vs the oracle."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from leopard_b200 import ops  # noqa: E402
from leopard_b200.config import leopard_idefics2, reduced_idefics2  # noqa: E402
from leopard_b200.idefics2 import Idefics2ForConditionalGeneration, navit_position_ids, patch_rectangles  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint  # noqa: E402
from oracle import idefics2 as oi  # noqa: E402

G = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(G))


def rel_err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float((got - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("Tq,Tk,H,HK,dh", [(64, 134, 16, 4, 96), (64, 4964, 16, 4, 96), (64, 64, 16, 4, 96),
                                            (100, 300, 8, 2, 128)])
def test_cross_attention(native_lib, Tq, Tk, H, HK, dh):
    g = torch.Generator().manual_seed(Tq + Tk)
    q = torch.randn(Tq, H * dh, generator=g).to(torch.bfloat16)
    kv = torch.randn(Tk, 2 * HK * dh, generator=g).to(torch.bfloat16)
    qq = q.float().view(Tq, H, dh).transpose(0, 1)
    k = kv.float()[:, : HK * dh].view(Tk, HK, dh).transpose(0, 1).repeat_interleave(H // HK, 0)
    v = kv.float()[:, HK * dh:].view(Tk, HK, dh).transpose(0, 1).repeat_interleave(H // HK, 0)
    ref = (torch.softmax(qq @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).transpose(0, 1).reshape(Tq, H * dh)
    out = ops.attention_cross(q.cuda(), kv.cuda(), 1, Tq, Tk, H, HK, dh, 0, HK * dh)
    err = (out.float().cpu() - ref).abs()
    assert bool((err <= 2e-2 * ref.abs().clamp(min=0.5)).all()), float(err.max())


def test_host_position_ids_match_oracle():
    pm = torch.zeros(2, 140, 182, dtype=torch.bool)
    pm[0] = True
    pm[1, :98, :140] = True
    rects = patch_rectangles(pm)
    assert rects == [(10, 13), (7, 10)]
    mask = oi.patch_mask_from_pixels(pm, 14)
    ref = oi.navit_position_ids(mask, 70)
    for b, (nh, nw) in enumerate(rects):
        assert torch.equal(navit_position_ids(nh, nw, 70), ref[b][mask[b].reshape(-1)])
    bad = pm.clone()
    bad[1, :14, 154:168] = True
    with pytest.raises(ValueError, match="rectangle"):
        patch_rectangles(bad)


@pytest.fixture(scope="module")
def small(native_lib):
    cfg = reduced_idefics2(leopard_idefics2(), vit_layers=2, perceiver_depth=2, dec_layers=2, vocab=1024)
    ck = SyntheticCheckpoint(cfg, seed=1234)
    model = Idefics2ForConditionalGeneration(cfg, lambda n: ck[n], "cuda:0")
    return cfg, ck, model


def test_prefill_matches_oracle_with_padding_and_dropped_image(small):
    import gen_idefics2_goldens as gg
    cfg, ck, model = small
    ids, pv, pm, n_img = gg.make_inputs(cfg)
    st = {}
    ref = oi.prefill_logits(ck, cfg, ids, pv, pm, stages=st)
    feats = model.engine.image_features(pv, pm)
    e_f = rel_err(feats, st["image_features"].reshape(-1, 4096))
    out = model(input_ids=ids, pixel_values=pv, pixel_attention_mask=pm, attention_mask=torch.ones_like(ids))
    e = rel_err(out.logits, ref)
    print(f"\n[parity] idefics2 reduced: image_features rel err {e_f:.4f}, logits rel err {e:.4f}")
    assert feats.shape == (n_img * 64, 4096) and e_f < 3e-2
    assert out.logits.shape == ref.shape and e < 3e-2
    gen = model.generate(input_ids=ids, pixel_values=pv, pixel_attention_mask=pm, max_new_tokens=3)
    assert gen.shape == (1, ids.shape[1] + 3) and torch.equal(gen[:, : ids.shape[1]].cpu(), ids)
    margin = ref[0, -1].topk(2).values
    if float(margin[0] - margin[1]) > 0.15:
        assert int(gen[0, ids.shape[1]]) == int(ref[0, -1].argmax())
    with pytest.raises(ValueError, match="image tokens"):
        model(input_ids=ids[:, :-70], pixel_values=pv, pixel_attention_mask=pm)


def test_full_resolution_image_tile_counts(small):
    """cfg4 geometry: a 980 x 980 image is 70 x 70 = 4900 patches -> 64 rows, 1 image here."""
    cfg, ck, model = small
    g = torch.Generator().manual_seed(2)
    pv = (torch.rand(1, 1, 3, 980, 980, generator=g) * 2 - 1).to(torch.bfloat16)
    feats = model.engine.image_features(pv, None)
    assert feats.shape == (64, 4096) and bool(torch.isfinite(feats.float()).all())
    hid = oi.vision_tower(ck, cfg.vit, pv[0].float(), torch.ones(1, 70, 70, dtype=torch.bool))
    ref = oi.connector(ck, cfg.perceiver, hid, torch.ones(1, 4900, dtype=torch.bool))
    assert rel_err(feats, ref[0]) < 3e-2


def test_megatron_style_provider_logits_and_loss(small):
    import gen_idefics2_goldens as gg
    from leopard_b200.megatron import idefics2_model_provider
    cfg, ck, _ = small
    model = idefics2_model_provider(cfg, lambda n: ck[n], "cuda:0")
    ids, pv, pm, _ = gg.make_inputs(cfg)
    ref = oi.prefill_logits(ck, cfg, ids, pv, pm)[0]
    logits = model(ids, None, None, pixel_values=pv, pixel_attention_mask=pm)
    assert rel_err(logits[0], ref) < 3e-2
    labels = torch.randint(0, cfg.dec.vocab, ids.shape)
    labels[0, :5] = -100
    loss = model(ids, None, None, labels=labels, pixel_values=pv, pixel_attention_mask=pm)
    ref_loss = torch.nn.functional.cross_entropy(ref.float(), labels[0], ignore_index=-100, reduction="none")
    assert loss.shape == ids.shape and float((loss[0].cpu() - ref_loss).abs().max()) < 0.08


def test_gpu_image_processor_feeds_the_model(small):
    """raw RGB images -> Idefics2Frontend (the processor's image half on the GPU) -> forward: the logits equal those of the same
    call fed with the oracle-processed tensors (the front-end is bit-exact), and match the oracle forward.  Sizes are multiples
    of the 14-pixel patch after the longest-edge rule, as the reference's 980-pixel setting gives for its slide images."""
    import numpy as np
    from leopard_b200.frontend import Idefics2Frontend
    from oracle import frontend as ofe
    cfg, ck, model = small
    rng = np.random.default_rng(4)
    imgs = [rng.integers(0, 256, (280, 364, 3), dtype=np.uint8), rng.integers(0, 256, (98, 140, 3), dtype=np.uint8)]
    pv, pm = Idefics2Frontend("cuda:0", longest_edge=182)(imgs)          # 364 x 280 -> 182 x 140; the small image is kept
    assert pv.shape == (1, 2, 3, 140, 182) and pm.shape == (1, 2, 140, 182)
    want_pv, want_pm = ofe.idefics2_preprocess(imgs, 182, 0)
    assert np.array_equal(pv[0].cpu().numpy(), want_pv) and np.array_equal(pm[0].cpu().numpy(), want_pm)
    g = torch.Generator().manual_seed(5)
    text = torch.randint(0, cfg.image_token_id - 2, (10,), generator=g)
    img = torch.full((cfg.perceiver.latents,), cfg.image_token_id)
    ids = torch.cat([text[:4], img, text[4:7], img, text[7:]])[None]
    out = model(input_ids=ids, pixel_values=pv, pixel_attention_mask=pm.bool(), attention_mask=torch.ones_like(ids))
    ref = oi.prefill_logits(ck, cfg, ids, torch.from_numpy(want_pv)[None].to(torch.bfloat16).float(), torch.from_numpy(want_pm)[None].bool())
    assert out.logits.shape == ref.shape and rel_err(out.logits, ref) < 3e-2
