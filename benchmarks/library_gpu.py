"""The "library GPU" competitor of BASELINE.md §3: the SAME prefill computed by stock libraries on the same B200 —
cuBLAS through `F.linear`, flash-attn 2 (`flash_attn_func`, what the reference's flash path calls:
Pai-Megatron-Patch/megatron_patch/model/llava/transformer.py:476-515,869-877) or torch SDPA when the wheel does not
load, and eager torch element-wise ops — in bf16, on the product's own weight tensors (kernel layout, no copy).

MEASUREMENT AID, not product code: only bench.py's `library_gpu` leg and tests import it.  It shares nothing with the
native kernels except the weight tensors, so it is also an independent check of the weight re-layout: its logits are
compared with the native path's in the bench line (`library_gpu.max_abs_diff_vs_native`) and with the CPU oracle in
tests/test_library_leg.py (fp32, reduced depth).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def _flash():
    try:
        from flash_attn import flash_attn_func  # noqa: PLC0415
        return flash_attn_func
    except Exception:  # noqa: BLE001
        return None


class LibraryPrefill:
    def __init__(self, cfg, named: dict, dtype=torch.bfloat16, attention: str = "auto"):
        self.cfg, self.dtype = cfg, dtype
        dev = named["embed"].device
        # vectors are stored fp32 by the native loader; libraries want them in the activation dtype
        self.w = {k: (v.to(dtype) if v.dtype != dtype and k != "dec.inv_freq" else v) for k, v in named.items()}
        self.inv_freq = named["dec.inv_freq"].float()
        self.flash = _flash() if (attention in ("auto", "flash") and dev.type == "cuda" and dtype != torch.float32) else None
        if attention == "flash" and self.flash is None:
            raise RuntimeError("flash_attn is not importable here")
        self.attention_impl = "flash_attn_func (flash-attn 2)" if self.flash else "torch SDPA"

    # [B, T, H, dh] in, [B, T, H, dh] out
    def _attn(self, q, k, v, causal: bool, scale: float):
        if self.flash is not None:
            return self.flash(q, k, v, softmax_scale=scale, causal=causal)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal,
                                           scale=scale, enable_gqa=q.shape[2] != k.shape[2])
        return o.transpose(1, 2)

    def vit(self, pix: torch.Tensor) -> torch.Tensor:
        """[N, 3, 364, 364] -> post-layernorm hidden states [N, 676, 1152]."""
        v, w = self.cfg.vit, self.w
        n = pix.shape[0]
        hs = w["vit.0.qkv_w"].shape[0] // (3 * v.heads) if v.layers else v.head_dim
        p = v.patch_size
        a = pix.to(self.dtype).unfold(2, p, p).unfold(3, p, p).permute(0, 2, 3, 1, 4, 5).reshape(n * v.tokens, 3 * p * p)
        a = F.pad(a, (0, w["vit.patch_w"].shape[1] - a.shape[1]))
        x = F.linear(a, w["vit.patch_w"], w["vit.patch_b"]).view(n, v.tokens, v.hidden) + w["vit.pos"][None]
        scale = v.head_dim ** -0.5
        for i in range(v.layers):
            g = f"vit.{i}."
            h = F.layer_norm(x, (v.hidden,), w[g + "ln1_g"], w[g + "ln1_b"], v.eps)
            qkv = F.linear(h, w[g + "qkv_w"], w[g + "qkv_b"]).view(n, v.tokens, 3, v.heads, hs)
            o = self._attn(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], False, scale)
            o = o[..., : v.head_dim].reshape(n, v.tokens, v.hidden)
            x = x + F.linear(o, w[g + "out_w"], w[g + "out_b"])
            h = F.layer_norm(x, (v.hidden,), w[g + "ln2_g"], w[g + "ln2_b"], v.eps)
            h = F.gelu(F.linear(h, w[g + "fc1_w"], w[g + "fc1_b"]), approximate="tanh")
            x = x + F.linear(h, w[g + "fc2_w"], w[g + "fc2_b"])
        return F.layer_norm(x, (v.hidden,), w["vit.post_g"], w["vit.post_b"], v.eps)

    def project(self, vit_out: torch.Tensor) -> torch.Tensor:
        """pixel_shuffle (the reference's view/permute chain, evaluations/models/llava_multiimg_siglip_anyres.py:165-176)
        + linear_1 / GELU / linear_2 -> [N, 169, 4096]."""
        cfg, w = self.cfg, self.w
        n, t, d = vit_out.shape
        side = int(round(math.sqrt(t)))
        half, s = side // cfg.shuffle, cfg.shuffle
        x = vit_out.view(n, half, s, half, s, d).permute(0, 1, 3, 2, 4, 5).reshape(n, half * half, s * s * d)
        h = F.linear(x, w["proj.w1"], w["proj.b1"])
        h = F.gelu(h) if cfg.projector_act == "gelu" else F.gelu(h, approximate="tanh")
        return F.linear(h, w["proj.w2"], w["proj.b2"])

    def merge(self, ids: torch.Tensor, feats: Optional[torch.Tensor]) -> torch.Tensor:
        """HF expand convention, B = 1: every placeholder id becomes its crop's 169 rows."""
        cfg = self.cfg
        ids = ids.reshape(-1)
        is_img = ids == cfg.image_token_index
        rep = torch.where(is_img, cfg.tokens_per_crop, 1)
        emb = self.w["embed"][ids.clamp(min=0)]
        out = torch.repeat_interleave(emb, rep, dim=0)
        if feats is not None:
            row_is_img = torch.repeat_interleave(is_img, rep)
            out[row_is_img] = feats.reshape(-1, feats.shape[-1]).to(out.dtype)
        return out

    def decoder(self, x: torch.Tensor, last_only: bool = True) -> torch.Tensor:
        """x [S, 4096] one sequence -> logits [1 or S, vocab_padded]."""
        d, w = self.cfg.dec, self.w
        S = x.shape[0]
        freqs = torch.arange(S, device=x.device, dtype=torch.float32)[:, None] * self.inv_freq[None]
        emb = torch.cat([freqs, freqs], -1)
        cos, sin = emb.cos().to(x.dtype)[:, None], emb.sin().to(x.dtype)[:, None]   # cast before the multiply (App. B)
        q_w, kv_w = d.heads * d.head_dim, d.kv_heads * d.head_dim
        scale = d.head_dim ** -0.5

        def rope(t):
            h = t.shape[-1] // 2
            return t * cos + torch.cat([-t[..., h:], t[..., :h]], -1) * sin

        for i in range(d.layers):
            g = f"dec.{i}."
            h = F.rms_norm(x, (d.hidden,), w[g + "ln1_g"], d.eps)
            qkv = F.linear(h, w[g + "qkv_w"])
            q = rope(qkv[:, :q_w].view(S, d.heads, d.head_dim))
            k = rope(qkv[:, q_w:q_w + kv_w].view(S, d.kv_heads, d.head_dim))
            v = qkv[:, q_w + kv_w:].view(S, d.kv_heads, d.head_dim)
            o = self._attn(q[None], k[None], v[None], True, scale)[0].reshape(S, d.hidden)
            x = x + F.linear(o, w[g + "o_w"])
            h = F.rms_norm(x, (d.hidden,), w[g + "ln2_g"], d.eps)
            gu = F.linear(h, w[g + "gateup_w"]).view(S, d.ffn // 32, 2, 32)       # 32-row gate / up granules
            x = x + F.linear((F.silu(gu[:, :, 0]) * gu[:, :, 1]).reshape(S, d.ffn), w[g + "down_w"])
        if last_only:
            x = x[-1:]
        x = F.rms_norm(x, (d.hidden,), w["dec.final_g"], d.eps)
        return F.linear(x, w["dec.lm_head"])

    @torch.no_grad()
    def prefill(self, ids: torch.Tensor, pix: Optional[torch.Tensor], last_only: bool = True) -> torch.Tensor:
        feats = self.project(self.vit(pix)) if pix is not None and pix.shape[0] else None
        return self.decoder(self.merge(ids, feats), last_only)[:, : self.cfg.dec.vocab]
