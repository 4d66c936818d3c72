"""Model hyper-parameters of the two Leopard variants on the prefill path (SURVEY.md §8 header line).

Leopard-LLaVA  = SigLIP-SO400M/14-364 + pixel-shuffle/2-layer MLP projector + Llama-3.1-8B
                 (Pai-Megatron-Patch/examples/llava/train_multiimg_llava_siglip.sh:56,73-93,179-180,220-228;
                  converter config Pai-Megatron-Patch/toolkits/model_checkpoints_convertor/llava/hf2megatron_llava.py:1026-1052)
Leopard-Idefics2 decoder = Mistral-7B (same widths; rope theta 1e4, no llama3 scaling, vocab 32003).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Optional


@dataclass(frozen=True)
class VitConfig:
    image_size: int = 364
    patch_size: int = 14
    hidden: int = 1152
    ffn: int = 4304
    heads: int = 16
    layers: int = 27
    eps: float = 1e-6

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size  # 26

    @property
    def tokens(self) -> int:
        return self.grid * self.grid  # 676

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads  # 72

    @property
    def patch_k(self) -> int:
        return 3 * self.patch_size * self.patch_size  # 588


@dataclass(frozen=True)
class RopeScaling:
    """llama3.1 frequency rescale (megatron/core/models/common/embeddings/rotary_pos_embedding.py:48-83)."""
    factor: float = 8.0
    low_freq_factor: float = 1.0
    high_freq_factor: float = 4.0
    original_max_position: int = 8192


@dataclass(frozen=True)
class DecoderConfig:
    hidden: int = 4096
    ffn: int = 14336
    heads: int = 32
    kv_heads: int = 8
    layers: int = 32
    vocab: int = 128256
    rope_theta: float = 500000.0
    rope_scaling: Optional[RopeScaling] = field(default_factory=RopeScaling)
    eps: float = 1e-5

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads  # 128


@dataclass(frozen=True)
class LeopardConfig:
    vit: VitConfig = field(default_factory=VitConfig)
    dec: DecoderConfig = field(default_factory=DecoderConfig)
    image_token_index: int = 128200        # <|reserved_special_token_195|>, EV/llava_multiimg_siglip_anyres.py:426-435
    megatron_image_index: int = -200       # IMAGE_TOKEN_INDEX, MP/model/llava/vlm_model.py:526
    projector_act: str = "gelu"            # HF: exact erf GELU; Megatron fused path: "gelu_pytorch_tanh" 
    shuffle: int = 2                       # pixel_shuffle scale factor

    @property
    def tokens_per_crop(self) -> int:
        return self.vit.tokens // (self.shuffle * self.shuffle)  # 169

    @property
    def proj_in(self) -> int:
        return self.vit.hidden * self.shuffle * self.shuffle  # 4608


@dataclass(frozen=True)
class PerceiverConfig:
    """Idefics2 perceiver resampler (Pai-Megatron-Patch/megatron_patch/model/idefics2/perceiver_transformer.py:1243-1723;
    HF Idefics2PerceiverConfig): 64 latents, 3 layers, 16 q / 4 kv heads x 96, gated MLP 4096 -> 16384 -> 4096."""
    latents: int = 64
    depth: int = 3
    heads: int = 16
    kv_heads: int = 4
    head_dim: int = 96
    hidden: int = 4096
    ffn: int = 16384
    eps: float = 1e-6


@dataclass(frozen=True)
class Idefics2Config:
    """Leopard-Idefics2 = NaViT SigLIP tower (<= 980 px, 70 x 70 position table) + modality projection (gated MLP
    1152 -> 14336 -> 4096) + perceiver + Mistral-7B (EX/idefics2/train_multiimg_idefics2.sh:93-98,234,239)."""
    vit: VitConfig = field(default_factory=lambda: VitConfig(image_size=980))
    perceiver: PerceiverConfig = field(default_factory=PerceiverConfig)
    dec: DecoderConfig = field(default_factory=lambda: DecoderConfig(vocab=32003, rope_theta=10000.0, rope_scaling=None))
    modality_ffn: int = 14336
    image_token_id: int = 32001


def leopard_idefics2() -> Idefics2Config:
    return Idefics2Config()


def reduced_idefics2(cfg: Idefics2Config, vit_layers: int, perceiver_depth: int, dec_layers: int,
                     vocab: Optional[int] = None) -> Idefics2Config:
    dec = replace(cfg.dec, layers=dec_layers, vocab=vocab or cfg.dec.vocab)
    return replace(cfg, vit=replace(cfg.vit, layers=vit_layers), perceiver=replace(cfg.perceiver, depth=perceiver_depth),
                   dec=dec, image_token_id=min(cfg.image_token_id, dec.vocab - 2))


def leopard_llava() -> LeopardConfig:
    return LeopardConfig()


def leopard_idefics2_decoder() -> DecoderConfig:
    """Mistral-7B parameterisation of the same decoder kernels."""
    return DecoderConfig(vocab=32003, rope_theta=10000.0, rope_scaling=None)


def reduced(cfg: LeopardConfig, vit_layers: int, dec_layers: int, vocab: Optional[int] = None) -> LeopardConfig:
    """Same widths, fewer layers / smaller vocabulary — for parity tests the CPU oracle finishes in seconds."""
    dec = replace(cfg.dec, layers=dec_layers, vocab=vocab or cfg.dec.vocab)
    return replace(cfg, vit=replace(cfg.vit, layers=vit_layers), dec=dec,
                   image_token_index=min(cfg.image_token_index, dec.vocab - 56))
