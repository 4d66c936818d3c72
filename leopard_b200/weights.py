"""Synthetic checkpoints in the reference's HF (transformers <= 4.4x LLaVA) key layout, and the key inventory.

No Leopard checkpoint exists offline, so parity and performance work uses seeded synthetic
weights: every tensor is drawn from its own generator seeded by (base seed, tensor name), in fp32 on
the CPU, and rounded once to bf16 — the SAME bf16 values feed the CPU oracle and the CUDA path.  Because each tensor is
independent of every other, the oracle and the loaders can stream layer by layer.

Key layout = target of the reference converter
(Pai-Megatron-Patch/toolkits/model_checkpoints_convertor/llava/hf2megatron_llava.py:1100-1460).
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterator, Tuple

import torch

from .config import LeopardConfig

V = "vision_tower.vision_model."
P = "multi_modal_projector."
LM = "language_model.model."


def tensor_specs(cfg: LeopardConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (name, shape, kind) for every tensor of the model; kind in {linear, bias, norm, embed, pos, conv}."""
    v, d = cfg.vit, cfg.dec
    yield V + "embeddings.patch_embedding.weight", (v.hidden, 3, v.patch_size, v.patch_size), "conv"
    yield V + "embeddings.patch_embedding.bias", (v.hidden,), "bias"
    yield V + "embeddings.position_embedding.weight", (v.tokens, v.hidden), "pos"
    for i in range(v.layers):
        b = f"{V}encoder.layers.{i}."
        yield b + "layer_norm1.weight", (v.hidden,), "norm"
        yield b + "layer_norm1.bias", (v.hidden,), "bias"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield b + f"self_attn.{n}.weight", (v.hidden, v.hidden), "linear"
            yield b + f"self_attn.{n}.bias", (v.hidden,), "bias"
        yield b + "layer_norm2.weight", (v.hidden,), "norm"
        yield b + "layer_norm2.bias", (v.hidden,), "bias"
        yield b + "mlp.fc1.weight", (v.ffn, v.hidden), "linear"
        yield b + "mlp.fc1.bias", (v.ffn,), "bias"
        yield b + "mlp.fc2.weight", (v.hidden, v.ffn), "linear"
        yield b + "mlp.fc2.bias", (v.hidden,), "bias"
    yield V + "post_layernorm.weight", (v.hidden,), "norm"
    yield V + "post_layernorm.bias", (v.hidden,), "bias"
    yield P + "linear_1.weight", (d.hidden, cfg.proj_in), "linear"
    yield P + "linear_1.bias", (d.hidden,), "bias"
    yield P + "linear_2.weight", (d.hidden, d.hidden), "linear"
    yield P + "linear_2.bias", (d.hidden,), "bias"
    yield LM + "embed_tokens.weight", (d.vocab, d.hidden), "embed"
    kv = d.kv_heads * d.head_dim
    for i in range(d.layers):
        b = f"{LM}layers.{i}."
        yield b + "input_layernorm.weight", (d.hidden,), "norm"
        yield b + "self_attn.q_proj.weight", (d.hidden, d.hidden), "linear"
        yield b + "self_attn.k_proj.weight", (kv, d.hidden), "linear"
        yield b + "self_attn.v_proj.weight", (kv, d.hidden), "linear"
        yield b + "self_attn.o_proj.weight", (d.hidden, d.hidden), "linear"
        yield b + "post_attention_layernorm.weight", (d.hidden,), "norm"
        yield b + "mlp.gate_proj.weight", (d.ffn, d.hidden), "linear"
        yield b + "mlp.up_proj.weight", (d.ffn, d.hidden), "linear"
        yield b + "mlp.down_proj.weight", (d.hidden, d.ffn), "linear"
    yield LM + "norm.weight", (d.hidden,), "norm"
    yield "language_model.lm_head.weight", (d.vocab, d.hidden), "linear"


def idefics2_tensor_specs(cfg) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """HF Idefics2 key layout (`model.vision_model.*`, `model.connector.*`, `model.text_model.*`, `lm_head.weight`)."""
    v, pc, d = cfg.vit, cfg.perceiver, cfg.dec
    VM, CN, TM = "model.vision_model.", "model.connector.", "model.text_model."
    yield VM + "embeddings.patch_embedding.weight", (v.hidden, 3, v.patch_size, v.patch_size), "conv"
    yield VM + "embeddings.patch_embedding.bias", (v.hidden,), "bias"
    yield VM + "embeddings.position_embedding.weight", (v.tokens, v.hidden), "pos"
    for i in range(v.layers):
        b = f"{VM}encoder.layers.{i}."
        yield b + "layer_norm1.weight", (v.hidden,), "norm"
        yield b + "layer_norm1.bias", (v.hidden,), "bias"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield b + f"self_attn.{n}.weight", (v.hidden, v.hidden), "linear"
            yield b + f"self_attn.{n}.bias", (v.hidden,), "bias"
        yield b + "layer_norm2.weight", (v.hidden,), "norm"
        yield b + "layer_norm2.bias", (v.hidden,), "bias"
        yield b + "mlp.fc1.weight", (v.ffn, v.hidden), "linear"
        yield b + "mlp.fc1.bias", (v.ffn,), "bias"
        yield b + "mlp.fc2.weight", (v.hidden, v.ffn), "linear"
        yield b + "mlp.fc2.bias", (v.hidden,), "bias"
    yield VM + "post_layernorm.weight", (v.hidden,), "norm"
    yield VM + "post_layernorm.bias", (v.hidden,), "bias"
    mp = CN + "modality_projection."
    yield mp + "gate_proj.weight", (cfg.modality_ffn, v.hidden), "linear"
    yield mp + "up_proj.weight", (cfg.modality_ffn, v.hidden), "linear"
    yield mp + "down_proj.weight", (d.hidden, cfg.modality_ffn), "linear"
    pr = CN + "perceiver_resampler."
    yield pr + "latents", (pc.latents, pc.hidden), "latents"
    for i in range(pc.depth):
        b = f"{pr}layers.{i}."
        yield b + "input_latents_norm.weight", (pc.hidden,), "norm"
        yield b + "input_context_norm.weight", (pc.hidden,), "norm"
        yield b + "self_attn.q_proj.weight", (pc.heads * pc.head_dim, pc.hidden), "linear"
        yield b + "self_attn.k_proj.weight", (pc.kv_heads * pc.head_dim, pc.hidden), "linear"
        yield b + "self_attn.v_proj.weight", (pc.kv_heads * pc.head_dim, pc.hidden), "linear"
        yield b + "self_attn.o_proj.weight", (pc.hidden, pc.heads * pc.head_dim), "linear"
        yield b + "post_attention_layernorm.weight", (pc.hidden,), "norm"
        yield b + "mlp.gate_proj.weight", (pc.ffn, pc.hidden), "linear"
        yield b + "mlp.up_proj.weight", (pc.ffn, pc.hidden), "linear"
        yield b + "mlp.down_proj.weight", (pc.hidden, pc.ffn), "linear"
    yield pr + "norm.weight", (pc.hidden,), "norm"
    yield TM + "embed_tokens.weight", (d.vocab, d.hidden), "embed"
    kv = d.kv_heads * d.head_dim
    for i in range(d.layers):
        b = f"{TM}layers.{i}."
        yield b + "input_layernorm.weight", (d.hidden,), "norm"
        yield b + "self_attn.q_proj.weight", (d.hidden, d.hidden), "linear"
        yield b + "self_attn.k_proj.weight", (kv, d.hidden), "linear"
        yield b + "self_attn.v_proj.weight", (kv, d.hidden), "linear"
        yield b + "self_attn.o_proj.weight", (d.hidden, d.hidden), "linear"
        yield b + "post_attention_layernorm.weight", (d.hidden,), "norm"
        yield b + "mlp.gate_proj.weight", (d.ffn, d.hidden), "linear"
        yield b + "mlp.up_proj.weight", (d.ffn, d.hidden), "linear"
        yield b + "mlp.down_proj.weight", (d.hidden, d.ffn), "linear"
    yield TM + "norm.weight", (d.hidden,), "norm"
    yield "lm_head.weight", (d.vocab, d.hidden), "linear"


def _seed_for(name: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:7], "little")


def synth_tensor(name: str, shape, kind: str, seed: int = 1234) -> torch.Tensor:
    """One synthetic tensor: fp32 draw on the CPU, rounded once to bf16 (returned as bf16)."""
    g = torch.Generator(device="cpu").manual_seed(_seed_for(name, seed))
    if kind in ("linear", "conv"):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        t = torch.randn(shape, generator=g) * fan_in ** -0.5
    elif kind == "bias":
        t = torch.randn(shape, generator=g) * 0.02
    elif kind == "norm":
        t = 1.0 + torch.randn(shape, generator=g) * 0.02
    elif kind == "embed":
        t = torch.randn(shape, generator=g) * 0.05
    elif kind == "pos":
        t = torch.randn(shape, generator=g) * 0.02
    elif kind == "latents":
        t = 1.0 + torch.randn(shape, generator=g) * 0.1   # the reference initialises latents to ones
    else:
        raise ValueError(kind)
    return t.to(torch.bfloat16)


class SyntheticCheckpoint:
    """Lazy HF-layout state dict: `ckpt[name]` materialises one bf16 tensor; nothing is cached."""

    def __init__(self, cfg, seed: int = 1234):
        self.cfg = cfg
        self.seed = seed
        specs = idefics2_tensor_specs(cfg) if hasattr(cfg, "perceiver") else tensor_specs(cfg)
        self._specs: Dict[str, Tuple[Tuple[int, ...], str]] = {n: (s, k) for n, s, k in specs}

    def keys(self):
        return self._specs.keys()

    def __contains__(self, name):
        return name in self._specs

    def shape(self, name):
        return self._specs[name][0]

    def __getitem__(self, name: str) -> torch.Tensor:
        shape, kind = self._specs[name]
        return synth_tensor(name, shape, kind, self.seed)

    def f32(self, name: str) -> torch.Tensor:
        """bf16-rounded values as fp32 (what the oracle computes with)."""
        return self[name].float()


def synth_inputs(cfg: LeopardConfig, n_crops: int, n_text: int, seed: int = 1, placeholder_per_crop: bool = True):
    """Seeded synthetic inputs in the HF-eval convention: pixels U(-1,1) fp32 [N,3,364,364] and a
    prompt with ONE placeholder id per crop (EV/llava_multiimg_siglip_anyres.py:426-435) surrounded by text ids."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    pix = torch.rand((n_crops, 3, cfg.vit.image_size, cfg.vit.image_size), generator=g) * 2.0 - 1.0
    g2 = torch.Generator(device="cpu").manual_seed(seed + 1)
    hi = min(cfg.dec.vocab, 128000)
    hi = min(hi, cfg.image_token_index)  # text ids never collide with the placeholder
    text = torch.randint(0, hi, (n_text,), generator=g2, dtype=torch.long)
    n_head = n_text // 2
    ids = torch.cat([text[:n_head], torch.full((n_crops,), cfg.image_token_index, dtype=torch.long), text[n_head:]])
    return pix, ids.unsqueeze(0)
