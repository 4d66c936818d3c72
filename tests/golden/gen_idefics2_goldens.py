"""Pin oracle/idefics2.py against the installed transformers Idefics2 (the third-party model the reference's Idefics2
eval driver loads, evaluations/models/idefics2_multiimg.py:22-30): full widths, reduced depth, seeded synthetic
checkpoint (leopard_b200.weights), variable-resolution images with padding masks and one all-zero padding image.
Writes idefics2_goldens.npz (sub-sampled outputs).  Run in the authoring container only."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from leopard_b200.config import leopard_idefics2, reduced_idefics2  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint  # noqa: E402

OUT = Path(__file__).resolve().parent
VOCAB = 1024


def make_inputs(cfg):
    g = torch.Generator().manual_seed(11)
    H, W = 140, 182
    pv = torch.zeros(1, 3, 3, H, W)
    pm = torch.zeros(1, 3, H, W, dtype=torch.bool)
    pv[0, 0] = torch.rand(3, H, W, generator=g) * 2 - 1
    pm[0, 0] = True                                   # 10 x 13 patches, all valid
    pv[0, 1, :, :98, :140] = torch.rand(3, 98, 140, generator=g) * 2 - 1
    pm[0, 1, :98, :140] = True                        # 7 x 10 valid patches, rest padding
    # image 2 stays all-zero: a padding image that must be dropped
    pv = pv.to(torch.bfloat16).float()
    n_img = 2
    text = torch.randint(0, cfg.image_token_id - 2, (12,), generator=g)
    img = torch.full((cfg.perceiver.latents,), cfg.image_token_id)
    ids = torch.cat([text[:3], img, text[3:6], img, text[6:]])[None]
    return ids, pv, pm, n_img


def main():
    from transformers import Idefics2Config, Idefics2ForConditionalGeneration
    cfg = reduced_idefics2(leopard_idefics2(), vit_layers=1, perceiver_depth=2, dec_layers=1, vocab=VOCAB)
    ck = SyntheticCheckpoint(cfg, seed=1234)
    v, pc, d = cfg.vit, cfg.perceiver, cfg.dec
    hc = Idefics2Config(
        vision_config=dict(hidden_size=v.hidden, intermediate_size=v.ffn, num_hidden_layers=v.layers,
                           num_attention_heads=v.heads, image_size=v.image_size, patch_size=v.patch_size,
                           hidden_act="gelu_pytorch_tanh", layer_norm_eps=v.eps),
        perceiver_config=dict(hidden_size=pc.hidden, resampler_n_latents=pc.latents, resampler_depth=pc.depth,
                              resampler_n_heads=pc.heads, resampler_head_dim=pc.head_dim,
                              num_key_value_heads=pc.kv_heads, rms_norm_eps=pc.eps, hidden_act="silu"),
        text_config=dict(model_type="mistral", hidden_size=d.hidden, intermediate_size=d.ffn,
                         num_hidden_layers=d.layers, num_attention_heads=d.heads, num_key_value_heads=d.kv_heads,
                         vocab_size=d.vocab, rms_norm_eps=d.eps, rope_theta=d.rope_theta, sliding_window=4096,
                         max_position_embeddings=32768, hidden_act="silu", tie_word_embeddings=False),
        image_token_id=cfg.image_token_id, tie_word_embeddings=False)
    hc._attn_implementation = "eager"
    hc.vision_config._attn_implementation = "eager"
    hc.perceiver_config._attn_implementation = "eager"
    hc.text_config._attn_implementation = "eager"
    m = Idefics2ForConditionalGeneration(hc).eval().float()
    sd = m.state_dict()
    new = {}
    for k in sd:
        assert k in ck, k
        new[k] = ck.f32(k).reshape(sd[k].shape)
    missing = [k for k in ck.keys() if k not in sd]
    assert not missing, missing
    m.load_state_dict(new, strict=True)
    ids, pv, pm, n_img = make_inputs(cfg)
    with torch.no_grad():
        feats = m.model.get_image_features(pv, pm).pooler_output
        out = m(input_ids=ids, pixel_values=pv, pixel_attention_mask=pm, attention_mask=torch.ones_like(ids))
    logits = out.logits[0]
    np.savez_compressed(OUT / "idefics2_goldens.npz", image_features=feats[:, ::7].numpy(),
                        logits=logits[:, ::2].numpy(), meta=np.array([v.layers, pc.depth, d.layers, VOCAB]))
    import transformers
    print("transformers", transformers.__version__, "feats", tuple(feats.shape), float(feats.abs().max()), "logits",
          tuple(logits.shape), float(logits.abs().max()))


if __name__ == "__main__":
    main()
