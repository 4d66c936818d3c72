"""CPU ORACLE, Idefics2 rows of the scope table — test infrastructure only.

Plain torch-CPU fp32 restatement of Leopard-Idefics2's prefill: NaViT-style SigLIP tower over variable-resolution
images (bucketised position ids, padded patches masked), gated-MLP modality projection, perceiver resampler
(64 latents, 3 layers, 16 q / 4 kv heads x 96), `inputs_merger`, Mistral-7B decoder.

Reference: the eval driver calls stock `AutoModelForVision2Seq` (evaluations/models/idefics2_multiimg.py:22-30,88-97),
i.e. third-party transformers code; the Megatron twin is
Pai-Megatron-Patch/megatron_patch/model/idefics2/idefics_vision_tower.py:104-178 (tower, NaViT position ids :129-150),
.../perceiver_transformer.py:582-706,1043-1143,1581-1723 (perceiver; K/V on concat(context, latents) :619-643) and
.../language_model.py:591-613 (inputs_merger).  PINNED against the installed transformers-5.5.0
Idefics2 modules with shared seeded weights (tests/golden/gen_idefics2_goldens.py -> tests/golden/idefics2_goldens.npz).

Weights: object with `f32(name)`; key layout = HF Idefics2 (`model.vision_model.*`, `model.connector.*`,
`model.text_model.*`, `lm_head.weight`).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import model as om

VM = "model.vision_model."
CN = "model.connector."
TM = "model.text_model."


def patch_mask_from_pixels(pixel_attention_mask: torch.Tensor, patch: int = 14) -> torch.Tensor:
    """[n, H, W] bool -> [n, H/14, W/14]: a patch is valid iff all of its pixels are (HF get_image_features;
    Megatron language_model.py:665-667)."""
    m = pixel_attention_mask.unfold(1, patch, patch).unfold(2, patch, patch)
    return m.sum(dim=(-1, -2)) == patch * patch


def navit_position_ids(patch_mask: torch.Tensor, side: int) -> torch.Tensor:
    """Bucketised fractional coordinates (idefics_vision_tower.py:129-150; HF Idefics2VisionEmbeddings):
    for an image with nh x nw valid patches, row i -> bucket(i / nh), col j -> bucket(j / nw) on a side x side grid."""
    n, gh, gw = patch_mask.shape
    boundaries = torch.arange(1 / side, 1.0, 1 / side)
    pos = torch.zeros(n, gh * gw, dtype=torch.long)
    for b in range(n):
        nh = int(patch_mask[b, :, 0].sum())
        nw = int(patch_mask[b, 0, :].sum())
        fh = torch.clamp(torch.arange(gh, dtype=torch.float32) * (1.0 / nh), max=1.0 - 1e-6)
        fw = torch.clamp(torch.arange(gw, dtype=torch.float32) * (1.0 / nw), max=1.0 - 1e-6)
        bh = torch.bucketize(fh, boundaries, right=True)
        bw = torch.bucketize(fw, boundaries, right=True)
        ids = (bh[:, None] * side + bw[None, :]).reshape(-1)
        flat = patch_mask[b].reshape(-1)
        pos[b][flat] = ids[flat]
    return pos


def vision_tower(ck, vit, pixel_values: torch.Tensor, patch_mask: torch.Tensor) -> torch.Tensor:
    """[n,3,H,W], mask [n,H/14,W/14] -> last_hidden_state [n, gh*gw, 1152] (values at masked patches are don't-care)."""
    n = pixel_values.shape[0]
    side = vit.image_size // vit.patch_size
    w = ck.f32(VM + "embeddings.patch_embedding.weight")
    b = ck.f32(VM + "embeddings.patch_embedding.bias")
    x = F.conv2d(pixel_values.float(), w, b, stride=vit.patch_size).flatten(2).transpose(1, 2)
    pos = navit_position_ids(patch_mask, side)
    x = x + ck.f32(VM + "embeddings.position_embedding.weight")[pos]
    key_ok = patch_mask.reshape(n, 1, 1, -1)
    for i in range(vit.layers):
        p = f"{VM}encoder.layers.{i}."
        d = x.shape[-1]
        h = F.layer_norm(x, (d,), ck.f32(p + "layer_norm1.weight"), ck.f32(p + "layer_norm1.bias"), vit.eps)
        q = F.linear(h, ck.f32(p + "self_attn.q_proj.weight"), ck.f32(p + "self_attn.q_proj.bias"))
        k = F.linear(h, ck.f32(p + "self_attn.k_proj.weight"), ck.f32(p + "self_attn.k_proj.bias"))
        v = F.linear(h, ck.f32(p + "self_attn.v_proj.weight"), ck.f32(p + "self_attn.v_proj.bias"))
        t = x.shape[1]
        q = q.view(n, t, vit.heads, vit.head_dim).transpose(1, 2)
        k = k.view(n, t, vit.heads, vit.head_dim).transpose(1, 2)
        v = v.view(n, t, vit.heads, vit.head_dim).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * vit.head_dim ** -0.5
        s = s.masked_fill(~key_ok, float("-inf"))
        a = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n, t, d)
        x = x + F.linear(a, ck.f32(p + "self_attn.out_proj.weight"), ck.f32(p + "self_attn.out_proj.bias"))
        h = F.layer_norm(x, (d,), ck.f32(p + "layer_norm2.weight"), ck.f32(p + "layer_norm2.bias"), vit.eps)
        h = F.gelu(F.linear(h, ck.f32(p + "mlp.fc1.weight"), ck.f32(p + "mlp.fc1.bias")), approximate="tanh")
        x = x + F.linear(h, ck.f32(p + "mlp.fc2.weight"), ck.f32(p + "mlp.fc2.bias"))
    return F.layer_norm(x, (x.shape[-1],), ck.f32(VM + "post_layernorm.weight"), ck.f32(VM + "post_layernorm.bias"),
                        vit.eps)


def gated_mlp(ck, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """down(silu(gate(x)) * up(x)) — Idefics2MLP (modality projection and perceiver MLP)."""
    g = F.linear(x, ck.f32(prefix + "gate_proj.weight"))
    u = F.linear(x, ck.f32(prefix + "up_proj.weight"))
    return F.linear(F.silu(g) * u, ck.f32(prefix + "down_proj.weight"))


def connector(ck, pc, image_hidden: torch.Tensor, patch_mask_flat: torch.Tensor) -> torch.Tensor:
    """[n, T, 1152] + key mask [n, T] -> [n, 64, 4096]: modality projection then `depth` perceiver layers
    (latents attend to concat(context, latents); padded context keys masked) and the final RMSNorm."""
    n = image_hidden.shape[0]
    ctx = gated_mlp(ck, CN + "modality_projection.", image_hidden)
    lat = ck.f32(CN + "perceiver_resampler.latents")[None].expand(n, -1, -1)
    key_ok = torch.cat([patch_mask_flat, torch.ones(n, lat.shape[1], dtype=torch.bool)], 1)[:, None, None, :]
    rep = pc.heads // pc.kv_heads
    for i in range(pc.depth):
        p = f"{CN}perceiver_resampler.layers.{i}."
        res = lat
        ln = om.rmsnorm(lat, ck.f32(p + "input_latents_norm.weight"), pc.eps)
        cn = om.rmsnorm(ctx, ck.f32(p + "input_context_norm.weight"), pc.eps)
        hs = torch.cat([cn, ln], 1)
        q = F.linear(ln, ck.f32(p + "self_attn.q_proj.weight")).view(n, -1, pc.heads, pc.head_dim).transpose(1, 2)
        k = F.linear(hs, ck.f32(p + "self_attn.k_proj.weight")).view(n, -1, pc.kv_heads, pc.head_dim).transpose(1, 2)
        v = F.linear(hs, ck.f32(p + "self_attn.v_proj.weight")).view(n, -1, pc.kv_heads, pc.head_dim).transpose(1, 2)
        k, v = k.repeat_interleave(rep, 1), v.repeat_interleave(rep, 1)
        s = (q @ k.transpose(-1, -2)) * pc.head_dim ** -0.5
        s = s.masked_fill(~key_ok, float("-inf"))
        a = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n, lat.shape[1], pc.heads * pc.head_dim)
        lat = res + F.linear(a, ck.f32(p + "self_attn.o_proj.weight"))
        h = om.rmsnorm(lat, ck.f32(p + "post_attention_layernorm.weight"), pc.eps)
        lat = lat + gated_mlp(ck, p + "mlp.", h)
    return om.rmsnorm(lat, ck.f32(CN + "perceiver_resampler.norm.weight"), pc.eps)


def inputs_merger(ck, cfg, input_ids: torch.Tensor, image_features: torch.Tensor) -> torch.Tensor:
    """rows where ids == image_token_id receive the image features in order, 64 per image
    (language_model.py:591-613; HF masked_scatter)."""
    emb = ck.f32(TM + "embed_tokens.weight")[input_ids]
    mask = input_ids == cfg.image_token_id
    flat = image_features.reshape(-1, image_features.shape[-1])
    assert int(mask.sum()) == flat.shape[0], f"image tokens {int(mask.sum())} != image feature rows {flat.shape[0]}"
    emb = emb.clone()
    emb[mask] = flat
    return emb


class _DecoderView:
    """Expose the Mistral decoder of an Idefics2 checkpoint under the Llama key names oracle.model expects."""

    def __init__(self, ck):
        self.ck = ck

    def f32(self, name):
        if name == "language_model.lm_head.weight":
            return self.ck.f32("lm_head.weight")
        return self.ck.f32(name.replace("language_model.model.", TM))


@torch.no_grad()
def prefill_logits(ck, cfg, input_ids: torch.Tensor, pixel_values: torch.Tensor, pixel_attention_mask: torch.Tensor,
                   stages=None) -> torch.Tensor:
    """input_ids [1, L] (64 image tokens per image), pixel_values [1, n, 3, H, W], pixel_attention_mask [1, n, H, W]
    -> logits [1, L, V].  All-zero padding images are dropped like the reference does (language_model.py:646,671-673)."""
    pv = pixel_values[0]
    pm = pixel_attention_mask[0].bool()
    real = (pv == 0).flatten(1).sum(1) != pv[0].numel()
    pv, pm = pv[real], pm[real]
    mask = patch_mask_from_pixels(pm, cfg.vit.patch_size)
    hid = vision_tower(ck, cfg.vit, pv, mask)
    feats = connector(ck, cfg.perceiver, hid, mask.flatten(1))
    emb = inputs_merger(ck, cfg, input_ids, feats)
    if stages is not None:
        stages.update(vision=hid, image_features=feats, embeds=emb, patch_mask=mask)

    class _C:  # oracle.model.decoder_forward reads cfg.dec only
        dec = cfg.dec
    return om.decoder_forward(_DecoderView(ck), _C, emb[0], torch.arange(emb.shape[1]))[None]
