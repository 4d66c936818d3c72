"""Parity mode: the fp32-EQUIVALENT forward of Leopard-LLaVA's prefill on the native kernels.

Why: the reference's HF evaluation runs in fp32 (`torch_dtype=torch.float32`, evaluations/models/
llava_multiimg_siglip_anyres.py:373) and BASELINE.json states the tolerance as "logits within 1e-3 max-abs of the reference
forward".  The fast path keeps activations in bf16 (what the reference's Megatron training does) and lands at ~6e-2 max-abs after
59 layers; that is bf16 resolution, not a bug, but it cannot demonstrate the stated tolerance.  This module can.

How (B200-first, no fp32 SIMT GEMM): every contraction runs on the SAME tcgen05 GEMM kernel (`lp_gemm_bf16`, fp32 accumulation in
TMEM, fp32 output) with each fp32 operand split into three bf16 terms laid out along the contraction dimension
(`lp_parity_split`): x = hi + mid + lo carries 24 mantissa bits, products of bf16 terms are exact in fp32, so
    x W^T = [hi | mid | lo] [W | W | W]^T                      (weights are bf16-exact: 3x the contraction length)
    q k^T = [qh qh qh qm qm ql] [kh km kl kh km kh]^T           (both operands fp32: the six significant cross terms)
Everything between the GEMMs (LayerNorm / RMSNorm, GELU, SwiGLU, softmax, RoPE, residual adds) runs in fp32 in the small kernels
of csrc/lp_parity.cu.  torch is used for allocation and pure data movement (views, transposes, gathers) only.

Cost: 3-6x the tensor work plus per-head attention launches — a verification path (seconds per request), selected explicitly:
    logits = ParityForward(engine).prefill_logits(input_ids, pixel_values)          # fp32 [S, vocab]
    model.set_precision("fp32")   # myLlavaForConditionalGeneration: forward() then answers with these logits
Measured against the CPU oracle (tests/test_gpu_parity_mode.py): see DESIGN.md §4.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib, ops

_F32 = torch.float32


def _ceil8(n: int) -> int:
    return (n + 7) // 8 * 8


class ParityForward:
    def __init__(self, engine):
        if not torch.cuda.is_available():
            raise _lib.NativeError("leopard_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.eng, self.cfg, self.w, self.dev = engine, engine.cfg, engine.w.named, engine.device
        _lib.lib()

    # ------------------------------------------------------------------ primitives
    def _split(self, x: torch.Tensor, layout: int, rows_pad: int = 0) -> torch.Tensor:
        """x fp32 [M, K] (row stride arbitrary) -> bf16 view [M (+ zero rows), n*K] with a 16-byte aligned row stride."""
        M, K = x.shape
        n = 3 if layout == 0 else 6
        out = torch.zeros((max(M, rows_pad), _ceil8(n * K)), dtype=torch.bfloat16, device=self.dev)
        ops.parity_split(x, out, layout)
        return out[:, : n * K]

    def _weight3(self, name: str, w: torch.Tensor) -> torch.Tensor:
        """[W | W | W] along K (bf16-exact weight against a three-term activation)."""
        K = w.shape[1]
        out = torch.empty((w.shape[0], _ceil8(3 * K)), dtype=torch.bfloat16, device=self.dev)
        for j in range(3):
            out[:, j * K:(j + 1) * K] = w
        return out[:, : 3 * K]

    def linear(self, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, name: str = "") -> torch.Tensor:
        """fp32 x [M, K] @ bf16 w [N, K]^T (+ fp32 bias) -> fp32 [M, N], products exact, fp32 accumulation."""
        K = x.shape[1]
        return ops.gemm(self._split(x, 0), self._weight3(name, w[:, :K]), bias=bias, out_fp32=True)

    def norm(self, x, gamma, beta, eps, rms):
        return ops.parity_norm(x, gamma, beta, eps, rms)

    def act(self, x, kind):
        return ops.parity_act(x, kind)

    def add_(self, y, x, row_mod=0):
        return ops.parity_add_(y, x, row_mod)

    def attention(self, q, k, v, scale: float, causal: bool) -> torch.Tensor:
        """One head: q [Tq, dh], k / v [Tk, dh] fp32 (column slices allowed) -> softmax(scale q k^T) v, fp32 [Tq, dh]."""
        Tq, dh = q.shape
        Tk = k.shape[0]
        Tkp = _ceil8(Tk)
        sc = ops.gemm(self._split(q, 1), self._split(k, 2, rows_pad=Tkp), out_fp32=True)          # [Tq, Tkp]; pad keys score 0
        ops.parity_softmax_(sc, Tk, scale, causal)
        vt = v.t().contiguous()                                                                  # [dh, Tk] (data movement)
        return ops.gemm(self._split(sc[:, :Tk], 1), self._split(vt, 2, rows_pad=_ceil8(dh)), out_fp32=True)[:, :dh]

    # ------------------------------------------------------------------ stages
    def vit(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """[N, 3, 364, 364] -> post-layernorm hidden states fp32 [N * 676, 1152] (HF SiglipVisionModel.last_hidden_state)."""
        v, w = self.cfg.vit, self.w
        n = pixel_values.shape[0]
        rows = n * v.tokens
        hs = self.w["vit.0.qkv_w"].shape[0] // (3 * v.heads) if v.layers else v.head_dim      # 80: heads are zero-padded
        A = ops.im2col_patch14(pixel_values.to(self.dev).contiguous(), self.w["vit.patch_w"].shape[1])   # patches rounded to bf16 once
        x = ops.gemm(A, w["vit.patch_w"], bias=w["vit.patch_b"], out_fp32=True)
        self.add_(x, w["vit.pos"].float(), row_mod=v.tokens)
        scale = v.head_dim ** -0.5
        for i in range(v.layers):
            g = f"vit.{i}."
            h = self.norm(x, w[g + "ln1_g"], w[g + "ln1_b"], v.eps, rms=False)
            qkv = self.linear(h, w[g + "qkv_w"], w[g + "qkv_b"])                        # [rows, 3 * heads * 80], pad columns 0
            att = torch.empty((rows, v.hidden), dtype=_F32, device=self.dev)
            for c in range(n):
                r = slice(c * v.tokens, (c + 1) * v.tokens)
                for hd in range(v.heads):
                    q0, k0, v0 = hd * hs, (v.heads + hd) * hs, (2 * v.heads + hd) * hs
                    att[r, hd * v.head_dim:(hd + 1) * v.head_dim] = self.attention(
                        qkv[r, q0:q0 + v.head_dim], qkv[r, k0:k0 + v.head_dim], qkv[r, v0:v0 + v.head_dim], scale, False)
            self.add_(x, self.linear(att, w[g + "out_w"], w[g + "out_b"]))
            h = self.norm(x, w[g + "ln2_g"], w[g + "ln2_b"], v.eps, rms=False)
            h = self.act(self.linear(h, w[g + "fc1_w"], w[g + "fc1_b"]), ops.ACT_GELU_TANH)
            self.add_(x, self.linear(h, w[g + "fc2_w"], w[g + "fc2_b"]))
        return self.norm(x, w["vit.post_g"], w["vit.post_b"], v.eps, rms=False)

    def image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """multi_modal_projector(vision_tower(pixel_values).last_hidden_state) -> fp32 [N, 169, 4096]."""
        cfg, w = self.cfg, self.w
        n = pixel_values.shape[0]
        x = self.vit(pixel_values).view(n, cfg.vit.tokens, cfg.vit.hidden)
        side = int(round(math.sqrt(cfg.vit.tokens)))
        half, s = side // cfg.shuffle, cfg.shuffle                                                  # pixel_shuffle: data movement
        x = x.view(n, half, s, half, s, cfg.vit.hidden).permute(0, 1, 3, 2, 4, 5).reshape(n * half * half, -1).contiguous()
        h = self.linear(x, w["proj.w1"], w["proj.b1"])
        h = self.act(h, ops.ACT_GELU_ERF if cfg.projector_act == "gelu" else ops.ACT_GELU_TANH)
        return self.linear(h, w["proj.w2"], w["proj.b2"]).view(n, half * half, -1)

    def merge(self, ids: torch.Tensor, feats: Optional[torch.Tensor]) -> torch.Tensor:
        """HF expand convention, one sequence: each placeholder id -> its crop's 169 rows (pure gather / scatter)."""
        cfg = self.cfg
        ids = ids.reshape(-1).to(self.dev)
        is_img = ids == cfg.image_token_index
        n_img = int(is_img.sum())
        n_feat = 0 if feats is None else feats.shape[0]
        if n_img != n_feat:
            raise ValueError(f"The input provided to the model are wrong. The number of image tokens is {n_img} while the number "
                             f"of image given to the model is {n_feat}. This prevents correct indexing and breaks batch generation.")
        rep = torch.where(is_img, cfg.tokens_per_crop, 1)
        out = torch.repeat_interleave(self.w["embed"][ids.clamp(min=0)].float(), rep, dim=0)
        if feats is not None:
            out[torch.repeat_interleave(is_img, rep)] = feats.reshape(-1, feats.shape[-1])
        return out

    def decoder(self, x: torch.Tensor, last_only: bool = False, rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x fp32 [S, 4096] (one sequence, positions 0..S-1) -> fp32 logits [S, vocab] (or of the last row / of `rows` only)."""
        d, w = self.cfg.dec, self.w
        S = x.shape[0]
        cs, sn = ops.rope_table(torch.arange(S, dtype=torch.int32, device=self.dev), w["dec.inv_freq"])
        q_w, kv_w = d.heads * d.head_dim, d.kv_heads * d.head_dim
        rep = d.heads // d.kv_heads
        scale = d.head_dim ** -0.5
        for i in range(d.layers):
            g = f"dec.{i}."
            h = self.norm(x, w[g + "ln1_g"], None, d.eps, rms=True)
            qkv = self.linear(h, w[g + "qkv_w"])
            ops.parity_rope_(qkv, d.heads + d.kv_heads, d.head_dim, cs, sn)
            att = torch.empty((S, d.hidden), dtype=_F32, device=self.dev)
            for hd in range(d.heads):
                kv = hd // rep
                att[:, hd * d.head_dim:(hd + 1) * d.head_dim] = self.attention(
                    qkv[:, hd * d.head_dim:(hd + 1) * d.head_dim], qkv[:, q_w + kv * d.head_dim: q_w + (kv + 1) * d.head_dim],
                    qkv[:, q_w + kv_w + kv * d.head_dim: q_w + kv_w + (kv + 1) * d.head_dim], scale, True)
            self.add_(x, self.linear(att, w[g + "o_w"]))
            h = self.norm(x, w[g + "ln2_g"], None, d.eps, rms=True)
            m = self.act(self.linear(h, w[g + "gateup_w"]), ops.ACT_SWIGLU)
            self.add_(x, self.linear(m, w[g + "down_w"]))
        if last_only:
            x = x[-1:].contiguous()
        elif rows is not None:
            x = x[rows.to(self.dev)].contiguous()
        x = self.norm(x, w["dec.final_g"], None, d.eps, rms=True)
        return self.linear(x, w["dec.lm_head"])[:, : d.vocab]

    @torch.no_grad()
    def prefill_logits(self, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor], last_only: bool = False,
                       stages: Optional[dict] = None, rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """input_ids [L] or [1, L] (one placeholder per crop), pixel_values [N, 3, 364, 364] -> fp32 logits [S, vocab]."""
        feats = None
        if pixel_values is not None and pixel_values.shape[0] > 0:
            feats = self.image_features(pixel_values)
        x = self.merge(input_ids, feats)
        if stages is not None:
            stages.update(image_features=feats, embeds=x.clone())
        return self.decoder(x, last_only, rows)
