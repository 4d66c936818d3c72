"""Pin the oracle's restatement of the THIRD-PARTY blocks the reference calls: run the installed
transformers SiglipVisionModel / LlamaForCausalLM (the modules `self.vision_tower(...)` and `self.language_model(...)`
resolve to, evaluations/models/llava_multiimg_siglip_anyres.py:268,322-331) on the seeded synthetic checkpoint of
leopard_b200.weights at FULL width and reduced depth, and store sub-sampled outputs in hf_goldens.npz.

Run in the authoring container (transformers 5.5.0, CPU, fp32, eager attention); the fixture is committed, this
script is the record of how it was made.  tests/test_oracle_goldens.py replays the same inputs through oracle/.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from leopard_b200.config import leopard_llava, leopard_idefics2_decoder, reduced  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint, synth_inputs  # noqa: E402
from dataclasses import replace  # noqa: E402

OUT = Path(__file__).resolve().parent
VIT_LAYERS, DEC_LAYERS, VOCAB = 2, 2, 2048
ROWS, COLS = slice(None, None, 13), slice(None, None, 9)


def siglip(cfg, ck, pix):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    v = cfg.vit
    hc = SiglipVisionConfig(hidden_size=v.hidden, intermediate_size=v.ffn, num_hidden_layers=v.layers,
                            num_attention_heads=v.heads, image_size=v.image_size, patch_size=v.patch_size,
                            hidden_act="gelu_pytorch_tanh", layer_norm_eps=v.eps, attention_dropout=0.0)
    hc._attn_implementation = "eager"
    m = SiglipVisionModel(hc).eval().float()
    sd = m.state_dict()
    pre = "vision_tower."
    new = {}
    for k in sd:
        name = pre + k
        if name in ck:
            new[k] = ck.f32(name).reshape(sd[k].shape)
    missing = [k for k in sd if k not in new]
    assert all(".head." in k or "position_ids" in k for k in missing), missing
    m.load_state_dict(new, strict=False)
    with torch.no_grad():
        return m(pixel_values=pix).last_hidden_state


def llama(cfg, ck, embeds, pos):
    from transformers import LlamaConfig, LlamaForCausalLM
    d = cfg.dec
    kw = dict(hidden_size=d.hidden, intermediate_size=d.ffn, num_hidden_layers=d.layers, num_attention_heads=d.heads,
              num_key_value_heads=d.kv_heads, vocab_size=d.vocab, rms_norm_eps=d.eps, max_position_embeddings=131072,
              attention_bias=False, mlp_bias=False, tie_word_embeddings=False, hidden_act="silu")
    if d.rope_scaling is not None:
        rs = d.rope_scaling
        rope = dict(rope_type="llama3", rope_theta=d.rope_theta, factor=rs.factor, low_freq_factor=rs.low_freq_factor,
                    high_freq_factor=rs.high_freq_factor, original_max_position_embeddings=rs.original_max_position)
    else:
        rope = dict(rope_type="default", rope_theta=d.rope_theta)
    try:
        hc = LlamaConfig(rope_parameters=rope, **kw)
    except TypeError:
        hc = LlamaConfig(rope_theta=d.rope_theta, rope_scaling=None if d.rope_scaling is None else rope, **kw)
    hc._attn_implementation = "eager"
    m = LlamaForCausalLM(hc).eval().float()
    sd = m.state_dict()
    new = {}
    for k in sd:
        name = "language_model." + k
        assert name in ck, name
        new[k] = ck.f32(name)
    m.load_state_dict(new, strict=True)
    with torch.no_grad():
        out = m(inputs_embeds=embeds[None], position_ids=pos[None], use_cache=False)
    return out.logits[0]


def main():
    cfg = reduced(leopard_llava(), VIT_LAYERS, DEC_LAYERS, VOCAB)
    ck = SyntheticCheckpoint(cfg, seed=1234)
    pix, _ = synth_inputs(cfg, n_crops=2, n_text=8, seed=1)
    pix = pix.to(torch.bfloat16).float()
    vit_out = siglip(cfg, ck, pix)
    g = torch.Generator().manual_seed(6)
    S = 96
    emb = (torch.randn(S, cfg.dec.hidden, generator=g) * 0.5).to(torch.bfloat16).float()
    # contiguous positions straddling the llama3.1 8192 "original context" boundary (transformers >= 4.5x treats
    # non-unit position steps as packed-sequence boundaries, so the positions must stay contiguous)
    pos = torch.arange(S) + 8150
    logits = llama(cfg, ck, emb, pos)
    # Mistral parameterisation (Idefics2 decoder): theta 1e4, no rescale — same seeded weights
    cfg_m = replace(cfg, dec=replace(leopard_idefics2_decoder(), layers=DEC_LAYERS, vocab=VOCAB))
    ck_m = SyntheticCheckpoint(cfg_m, seed=1234)
    logits_m = llama(cfg_m, ck_m, emb, pos)
    np.savez_compressed(OUT / "hf_goldens.npz",
                        vit_out=vit_out[:, ROWS, COLS].numpy(), vit_absmax=float(vit_out.abs().max()),
                        dec_pos=pos.numpy(), dec_logits=logits[:, ::3].numpy(),
                        dec_logits_mistral=logits_m[:, ::3].numpy(),
                        meta=np.array([VIT_LAYERS, DEC_LAYERS, VOCAB, S]))
    import transformers
    print("transformers", transformers.__version__, "vit absmax", float(vit_out.abs().max()),
          "logits absmax", float(logits.abs().max()))


if __name__ == "__main__":
    main()
