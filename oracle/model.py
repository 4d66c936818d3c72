"""CPU ORACLE (test infrastructure — never imported by the product path).

Plain torch-CPU fp32 restatement of Leopard-LLaVA's multi-image prefill arithmetic, used ONLY as the checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.

PARITY PINNING: the reference has no tests or golden vectors for this path and neither of its two
implementations runs in this image (HF eval needs transformers <= 4.4x APIs; Megatron needs TransformerEngine).  The
blocks below are therefore pinned against what CAN run here:
  * pixel_shuffle / projector / crop planner: executed from the reference source itself
    (tests/golden/gen_reference_goldens.py -> tests/golden/ref_*.npz);
  * SigLIP tower, Llama decoder layer, RMSNorm, llama3.1 RoPE: the installed transformers-5.5.0 modules the reference
    calls as third-party code, with shared seeded weights (tests/golden/gen_hf_goldens.py -> tests/golden/hf_*.npz);
  * HF `_merge_input_ids_with_image_features` (transformers 4.38-4.4x, absent here): restated from the published
    algorithm — "parity unpinned" for that one function beyond the B=1 reduction the reference uses.

Every function cites the reference lines it follows (paths relative to the reference checkout).
Weights come from any object with `f32(name) -> fp32 tensor` using the HF key layout (leopard_b200.weights).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

V = "vision_tower.vision_model."
P = "multi_modal_projector."
LM = "language_model.model."


# ----------------------------------------------------------------------------------------------------------------
# SigLIP vision tower  (HF SiglipVisionModel as called at evaluations/models/llava_multiimg_siglip_anyres.py:268,273;
# Megatron twin Pai-Megatron-Patch/megatron_patch/model/idefics2/idefics_vision_tower.py:104-178 with the layer spec
# of megatron_patch/model/llava/clip_encoder.py:51-72,318-351)
# ----------------------------------------------------------------------------------------------------------------
def vit_embed(ck, cfg, pixel_values: torch.Tensor) -> torch.Tensor:
    """conv k=14 s=14 + bias, flatten row-major, + learned position embedding (ids 0..675)."""
    v = cfg.vit
    w = ck.f32(V + "embeddings.patch_embedding.weight")
    b = ck.f32(V + "embeddings.patch_embedding.bias")
    x = F.conv2d(pixel_values.float(), w, b, stride=v.patch_size)  # [N, hidden, 26, 26]
    x = x.flatten(2).transpose(1, 2)  # [N, 676, hidden]
    return x + ck.f32(V + "embeddings.position_embedding.weight")[None]


def vit_layer(ck, cfg, x: torch.Tensor, i: int) -> torch.Tensor:
    """pre-LN block: LN -> qkv(+bias) -> MHA (16 x 72, scale 72^-1/2, no mask) -> out(+bias) -> +res;
    LN -> fc1(+bias) -> GELU-tanh -> fc2(+bias) -> +res."""
    v = cfg.vit
    b = f"{V}encoder.layers.{i}."
    n, t, d = x.shape
    h = F.layer_norm(x, (d,), ck.f32(b + "layer_norm1.weight"), ck.f32(b + "layer_norm1.bias"), v.eps)
    q = F.linear(h, ck.f32(b + "self_attn.q_proj.weight"), ck.f32(b + "self_attn.q_proj.bias"))
    k = F.linear(h, ck.f32(b + "self_attn.k_proj.weight"), ck.f32(b + "self_attn.k_proj.bias"))
    vv = F.linear(h, ck.f32(b + "self_attn.v_proj.weight"), ck.f32(b + "self_attn.v_proj.bias"))
    q = q.view(n, t, v.heads, v.head_dim).transpose(1, 2)
    k = k.view(n, t, v.heads, v.head_dim).transpose(1, 2)
    vv = vv.view(n, t, v.heads, v.head_dim).transpose(1, 2)
    s = torch.softmax((q @ k.transpose(-1, -2)) * v.head_dim ** -0.5, dim=-1)
    a = (s @ vv).transpose(1, 2).reshape(n, t, d)
    x = x + F.linear(a, ck.f32(b + "self_attn.out_proj.weight"), ck.f32(b + "self_attn.out_proj.bias"))
    h = F.layer_norm(x, (d,), ck.f32(b + "layer_norm2.weight"), ck.f32(b + "layer_norm2.bias"), v.eps)
    h = F.gelu(F.linear(h, ck.f32(b + "mlp.fc1.weight"), ck.f32(b + "mlp.fc1.bias")), approximate="tanh")
    return x + F.linear(h, ck.f32(b + "mlp.fc2.weight"), ck.f32(b + "mlp.fc2.bias"))


def vit_forward(ck, cfg, pixel_values: torch.Tensor) -> torch.Tensor:
    """[N,3,364,364] -> last_hidden_state [N,676,1152] (post_layernorm applied; pooling head skipped — the
    reference discards it, EV/...anyres.py:273)."""
    x = vit_embed(ck, cfg, pixel_values)
    for i in range(cfg.vit.layers):
        x = vit_layer(ck, cfg, x, i)
    return F.layer_norm(x, (cfg.vit.hidden,), ck.f32(V + "post_layernorm.weight"), ck.f32(V + "post_layernorm.bias"),
                        cfg.vit.eps)


# ----------------------------------------------------------------------------------------------------------------
# pixel shuffle + projector  (EV/llava_multiimg_siglip_anyres.py:165-192; MP/model/llava/vlm_model.py:456-475)
# ----------------------------------------------------------------------------------------------------------------
def pixel_shuffle(x: torch.Tensor, scale: int = 2) -> torch.Tensor:
    """out[b, oh*W2 + ow] = cat(x[b,(2oh)*W+2ow], x[b,(2oh)*W+2ow+1], x[b,(2oh+1)*W+2ow], x[b,(2oh+1)*W+2ow+1])."""
    n, t, d = x.shape
    side = int(round(math.sqrt(t)))
    half = side // scale
    g = x.view(n, half, scale, half, scale, d)          # [b, oh, dy, ow, dx, e]
    return g.permute(0, 1, 3, 2, 4, 5).reshape(n, half * half, scale * scale * d)


def projector(ck, cfg, vit_out: torch.Tensor) -> torch.Tensor:
    """linear_2(act(linear_1(pixel_shuffle(x)))); act = exact GELU (HF) or tanh GELU (Megatron fused path, D8)."""
    h = pixel_shuffle(vit_out, cfg.shuffle)
    h = F.linear(h, ck.f32(P + "linear_1.weight"), ck.f32(P + "linear_1.bias"))
    h = F.gelu(h) if cfg.projector_act == "gelu" else F.gelu(h, approximate="tanh")
    return F.linear(h, ck.f32(P + "linear_2.weight"), ck.f32(P + "linear_2.bias"))


# ----------------------------------------------------------------------------------------------------------------
# embedding + merge
# ----------------------------------------------------------------------------------------------------------------
def merge_hf_expand(ck, cfg, input_ids: torch.Tensor, image_features: torch.Tensor):
    """HF `_merge_input_ids_with_image_features` for the only case the reference uses: B = 1, no padding, mask all
    ones (EV/...anyres.py:285-287,443-446).  Each placeholder id expands to its crop's 169 rows in
    crop order.  Returns (embeds [1,S,D], position_ids [1,S])."""
    assert input_ids.shape[0] == 1, "the reference only ever merges B=1 (see SURVEY.md App. A.1)"
    ids = input_ids[0]
    table = ck.f32(LM + "embed_tokens.weight")
    n, p, d = image_features.shape
    is_img = ids == cfg.image_token_index
    if int(is_img.sum()) != n:
        raise ValueError(
            f"The input provided to the model are wrong. The number of image tokens is {int(is_img.sum())} while "
            f"the number of image given to the model is {n}. This prevents correct indexing and breaks batch generation.")
    rows = []
    k = 0
    emb = table[ids.clamp(min=0)]
    for t in range(ids.numel()):
        if bool(is_img[t]):
            rows.append(image_features[k])
            k += 1
        else:
            rows.append(emb[t:t + 1])
    out = torch.cat(rows, 0)[None]
    pos = torch.arange(out.shape[1])[None]
    return out, pos


def merge_hf_batched(ck, cfg, input_ids: torch.Tensor, attention_mask: torch.Tensor, image_features: torch.Tensor,
                     pad_token_id: int):
    """HF `_merge_input_ids_with_image_features` for a padded batch, restated step by step from SURVEY.md App. A.1 (the
    transformers 4.38-4.4x algorithm the reference's forward calls at EV/...anyres.py:285-287; "parity unpinned": the source
    is in neither tree).  input_ids / attention_mask [B, L], image_features [N, P, D].  Returns (final_embedding [B, max_len, D],
    final_attention_mask [B, max_len], position_ids [B, max_len])."""
    table = ck.f32(LM + "embed_tokens.weight")
    n, p, d = image_features.shape
    B, L = input_ids.shape
    inputs_embeds = table[input_ids.clamp(min=0)]
    m = input_ids == cfg.image_token_index
    # 1. target positions
    max_len = int(m.sum(-1).max()) * (p - 1) + L
    new_pos = torch.cumsum(m.long() * (p - 1) + 1, -1) - 1
    nb_image_pad = max_len - 1 - new_pos[:, -1]
    left_padding = not bool((input_ids[:, -1] == pad_token_id).any())
    if left_padding:
        new_pos = new_pos + nb_image_pad[:, None]
    # 2. text tokens
    final_embedding = torch.zeros(B, max_len, d)
    final_mask = torch.zeros(B, max_len, dtype=torch.long)
    bi, ti = torch.where(~m)
    final_embedding[bi, new_pos[bi, ti]] = inputs_embeds[bi, ti]
    final_mask[bi, new_pos[bi, ti]] = attention_mask[bi, ti].long()
    # 3. image rows: the still-empty positions past the image padding, filled in row-major order
    image_rows = torch.ones(B, max_len, dtype=torch.bool)
    image_rows[bi, new_pos[bi, ti]] = False
    image_rows &= (image_rows.long().cumsum(-1) - 1) >= nb_image_pad[:, None]
    if int(image_rows.sum()) != n * p:
        raise ValueError(
            f"The input provided to the model are wrong. The number of image tokens is {int(m.sum())} while "
            f"the number of image given to the model is {n}. This prevents correct indexing and breaks batch generation.")
    final_embedding[image_rows] = image_features.reshape(-1, d)
    final_mask |= image_rows.long()
    # 4. positions; rows of pad tokens are zeroed again
    position_ids = (final_mask.cumsum(-1) - 1).masked_fill(final_mask == 0, 1)
    pb, pt = torch.where(input_ids == pad_token_id)
    final_embedding[pb, new_pos[pb, pt]] = 0
    return final_embedding, final_mask, position_ids


def merge_megatron_scatter(ck, cfg, input_ids: torch.Tensor, image_features: torch.Tensor):
    """Megatron convention (MP/model/llava/vlm_model.py:516-533): ids hold one -200 per feature ROW; embeddings of
    clamp(ids, 0) are overwritten, in row-major order, by image_features.view(-1, D)."""
    table = ck.f32(LM + "embed_tokens.weight")
    emb = table[input_ids.clamp(min=0)]  # [b, s, d]
    mask = input_ids == cfg.megatron_image_index
    flat = image_features.reshape(-1, image_features.shape[-1])
    assert flat.shape[0] == int(mask.sum()), \
        f"The number of image features doesn't match the number of image tokens. {flat.shape[0]} != {int(mask.sum())}"
    emb = emb.clone()
    emb[mask] = flat
    return emb


# ----------------------------------------------------------------------------------------------------------------
# Llama / Mistral decoder  (HF LlamaForCausalLM as called at EV/...anyres.py:322-331; Megatron twin
# MP/model/llava/transformer.py:678-885 (attention), :160-176 (MLP), :1208-1340 (layer), :1747-1860 (stack))
# ----------------------------------------------------------------------------------------------------------------
def rope_inv_freq(dec) -> torch.Tensor:
    """inv_freq = theta^(-2i/d), optionally llama3.1-rescaled
    (megatron/core/models/common/embeddings/rotary_pos_embedding.py:48-83,112-123)."""
    d = dec.head_dim
    inv = 1.0 / (dec.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    rs = dec.rope_scaling
    if rs is None:
        return inv
    low_wl = rs.original_max_position / rs.low_freq_factor
    high_wl = rs.original_max_position / rs.high_freq_factor
    out = []
    for f in inv.tolist():
        wl = 2 * math.pi / f
        if wl < high_wl:
            out.append(f)
        elif wl > low_wl:
            out.append(f / rs.factor)
        else:
            smooth = (rs.original_max_position / wl - rs.low_freq_factor) / (rs.high_freq_factor - rs.low_freq_factor)
            out.append((1 - smooth) * f / rs.factor + smooth * f)
    return torch.tensor(out, dtype=torch.float32)


def rope_cos_sin(dec, position_ids: torch.Tensor):
    """[S] positions -> cos, sin [S, head_dim] with emb = cat(freqs, freqs) (rotary_pos_embedding.py:125-157)."""
    freqs = position_ids.float()[:, None] * rope_inv_freq(dec)[None, :]
    emb = torch.cat([freqs, freqs], -1)
    return emb.cos(), emb.sin()


def apply_rope(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """t [.., S, H, D]; t*cos + rotate_half(t)*sin with rotate_half(x) = cat(-x2, x1) (rotary_pos_embedding.py:197-239)."""
    d = t.shape[-1]
    x1, x2 = t[..., : d // 2], t[..., d // 2:]
    rot = torch.cat([-x2, x1], -1)
    return t * cos[:, None, :] + rot * sin[:, None, :]


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """x * rsqrt(mean(x^2) + eps) * w in fp32 (megatron/legacy/model/rms_norm.py:26-31)."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def decoder_layer(ck, cfg, x: torch.Tensor, i: int, cos, sin, kv_out: Optional[list] = None) -> torch.Tensor:
    """x [S, D] (one sequence).  RMSNorm -> q,k,v -> RoPE(q,k) -> causal GQA SDPA (q head j uses kv head j // 4,
    transformer.py:829-836) -> o -> +res; RMSNorm -> silu(gate)*up -> down -> +res."""
    d = cfg.dec
    b = f"{LM}layers.{i}."
    s = x.shape[0]
    h = rmsnorm(x, ck.f32(b + "input_layernorm.weight"), d.eps)
    q = F.linear(h, ck.f32(b + "self_attn.q_proj.weight")).view(s, d.heads, d.head_dim)
    k = F.linear(h, ck.f32(b + "self_attn.k_proj.weight")).view(s, d.kv_heads, d.head_dim)
    v = F.linear(h, ck.f32(b + "self_attn.v_proj.weight")).view(s, d.kv_heads, d.head_dim)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    if kv_out is not None:
        kv_out.append((k.clone(), v.clone()))
    rep = d.heads // d.kv_heads
    kk = k.repeat_interleave(rep, dim=1).transpose(0, 1)  # [H, S, D]
    vv = v.repeat_interleave(rep, dim=1).transpose(0, 1)
    qq = q.transpose(0, 1)
    causal = torch.ones(s, s, dtype=torch.bool).triu(1)
    outs = []
    for h0 in range(0, d.heads, 8):  # chunk heads: bounds the [h, S, S] score tensor
        sc = (qq[h0:h0 + 8] @ kk[h0:h0 + 8].transpose(-1, -2)) * d.head_dim ** -0.5
        sc = sc.masked_fill(causal, float("-inf"))
        outs.append(torch.softmax(sc, -1) @ vv[h0:h0 + 8])
    a = torch.cat(outs, 0).transpose(0, 1).reshape(s, d.hidden)
    x = x + F.linear(a, ck.f32(b + "self_attn.o_proj.weight"))
    h = rmsnorm(x, ck.f32(b + "post_attention_layernorm.weight"), d.eps)
    g = F.linear(h, ck.f32(b + "mlp.gate_proj.weight"))
    u = F.linear(h, ck.f32(b + "mlp.up_proj.weight"))
    return x + F.linear(F.silu(g) * u, ck.f32(b + "mlp.down_proj.weight"))


def decoder_forward(ck, cfg, embeds: torch.Tensor, position_ids: torch.Tensor, last_only: bool = False,
                    kv_out: Optional[list] = None, return_hidden: bool = False) -> torch.Tensor:
    """embeds [S, D] -> logits [S, V] (all rows, as the reference computes, EV/...anyres.py:333) or [1, V]."""
    cos, sin = rope_cos_sin(cfg.dec, position_ids)
    x = embeds
    for i in range(cfg.dec.layers):
        x = decoder_layer(ck, cfg, x, i, cos, sin, kv_out)
    x = rmsnorm(x, ck.f32(LM + "norm.weight"), cfg.dec.eps)
    if return_hidden:
        return x
    if last_only:
        x = x[-1:]
    return F.linear(x, ck.f32("language_model.lm_head.weight"))


@torch.no_grad()
def prefill_logits_batched(ck, cfg, input_ids: torch.Tensor, attention_mask: torch.Tensor, pixel_values: torch.Tensor,
                           pad_token_id: int):
    """Padded batch through the HF forward (EV/...anyres.py:263-333): merge_hf_batched, then the decoder per batch row on the rows
    the attention mask keeps, at the merge's position ids — masked keys get an exactly-zero softmax weight in the fp32 reference,
    so the kept rows see the same arithmetic as in the padded [B, max_len] call.  Returns (logits [B, max_len, V] with zeros at
    masked rows, final_attention_mask)."""
    feats = projector(ck, cfg, vit_forward(ck, cfg, pixel_values))
    emb, mask, pos = merge_hf_batched(ck, cfg, input_ids, attention_mask, feats, pad_token_id)
    out = None
    for b in range(emb.shape[0]):
        keep = mask[b].bool()
        lg = decoder_forward(ck, cfg, emb[b][keep], pos[b][keep])
        if out is None:
            out = torch.zeros(emb.shape[0], emb.shape[1], lg.shape[-1])
        out[b][keep] = lg
    return out, mask


# ----------------------------------------------------------------------------------------------------------------
# end-to-end prefill  (myLlavaForConditionalGeneration.forward, EV/...anyres.py:201-361, prefill branch)
# ----------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def prefill_logits(ck, cfg, input_ids: torch.Tensor, pixel_values: torch.Tensor, last_only: bool = False,
                   stages: Optional[dict] = None) -> torch.Tensor:
    """HF-eval convention: input_ids [1, L] with one placeholder per crop, pixel_values [N,3,364,364] fp32.
    Returns logits [1, S, V] (or [1, 1, V]).  `stages`, when given, receives intermediate tensors."""
    vit_out = vit_forward(ck, cfg, pixel_values)
    feats = projector(ck, cfg, vit_out)
    embeds, pos = merge_hf_expand(ck, cfg, input_ids, feats)
    if stages is not None:
        stages.update(vit_out=vit_out, image_features=feats, embeds=embeds)
    logits = decoder_forward(ck, cfg, embeds[0], pos[0], last_only=last_only)
    return logits[None]
