"""Leopard-Idefics2 prefill on the native path and its plugin surface.

Reference surface (evaluations/models/idefics2_multiimg.py:22-30,88-97): `processor, model = load_model(args)`;
`model.generate(**{input_ids, attention_mask, pixel_values[B,n,3,H,W], pixel_attention_mask[B,n,H,W]},
max_new_tokens=128)`; the decoded text is split on 'Assistant: ', so the returned ids INCLUDE the prompt.

Pipeline (all arithmetic through the C ABI):
  NaViT SigLIP tower   padded patches are REMOVED on the host (valid patches of an image are its top-left nh x nw
                       block), position ids are the bucketised fractional coordinates
                       (megatron_patch/model/idefics2/idefics_vision_tower.py:129-150) gathered by lp_embed_merge and
                       added in the patch-embedding GEMM's residual epilogue; lp_vit_layers runs the 27 layers with
                       per-image attention — identical on valid patches to the reference's key-padding mask.
  modality projection  gated MLP 1152 -> 14336 -> 4096 = SwiGLU-epilogue GEMM + GEMM
  perceiver resampler  3 x { RMSNorm(latents), RMSNorm(context), q = latents Wq, kv = concat(context, latents) Wkv,
                       lp_attention_cross (16 q / 4 kv heads x 96), o-proj + residual, RMSNorm, gated MLP + residual },
                       final RMSNorm  (perceiver_transformer.py:582-706,1043-1143,1581-1723)
  merge + decoder      lp_embed_merge (one image-token id per feature row) + lp_decoder_prefill with Mistral RoPE.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Callable, List, Optional

import torch

from . import _lib, ops
from .config import Idefics2Config, LeopardConfig
from .engine import PrefillEngine
from .native_weights import load_hf

VM, CN, TM = "model.vision_model.", "model.connector.", "model.text_model."


def _interleave_gate_up(g: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    f, h = g.shape
    return torch.cat([g.reshape(f // 32, 1, 32, h), u.reshape(f // 32, 1, 32, h)], 1).reshape(2 * f, h)


class Idefics2Weights:
    def __init__(self, cfg: Idefics2Config, get: Callable[[str], torch.Tensor], device="cuda"):
        self.cfg = cfg
        dev = torch.device(device)
        # tower + decoder reuse the LLaVA loader through a name translation (same tensor shapes, other prefixes)
        base = LeopardConfig(vit=cfg.vit, dec=cfg.dec, image_token_index=cfg.image_token_id)

        def tr(name: str) -> torch.Tensor:
            if name == "language_model.lm_head.weight":
                return get("lm_head.weight")
            return get(name.replace("vision_tower.vision_model.", VM).replace("language_model.model.", TM))

        self.base = load_hf(base, tr, dev, with_projector=False)
        bf = lambda t: t.to(dev, torch.bfloat16).contiguous()  # noqa: E731
        f32 = lambda t: t.to(dev).to(torch.bfloat16).float().contiguous()  # noqa: E731
        mp = CN + "modality_projection."
        self.mp_gateup = bf(_interleave_gate_up(get(mp + "gate_proj.weight").to(dev), get(mp + "up_proj.weight").to(dev)))
        self.mp_down = bf(get(mp + "down_proj.weight"))
        pr = CN + "perceiver_resampler."
        self.latents = bf(get(pr + "latents"))
        self.layers = []
        for i in range(cfg.perceiver.depth):
            b = f"{pr}layers.{i}."
            self.layers.append(SimpleNamespace(
                lat_norm=f32(get(b + "input_latents_norm.weight")), ctx_norm=f32(get(b + "input_context_norm.weight")),
                q_w=bf(get(b + "self_attn.q_proj.weight")),
                kv_w=bf(torch.cat([get(b + "self_attn.k_proj.weight").to(dev), get(b + "self_attn.v_proj.weight").to(dev)], 0)),
                o_w=bf(get(b + "self_attn.o_proj.weight")), post_norm=f32(get(b + "post_attention_layernorm.weight")),
                gateup=bf(_interleave_gate_up(get(b + "mlp.gate_proj.weight").to(dev), get(b + "mlp.up_proj.weight").to(dev))),
                down=bf(get(b + "mlp.down_proj.weight"))))
        self.final_norm = f32(get(pr + "norm.weight"))
        self.pos_table = bf(get(VM + "embeddings.position_embedding.weight"))


def patch_rectangles(pixel_attention_mask: torch.Tensor, patch: int = 14):
    """[n, H, W] bool -> list of (nh, nw): the valid patches of each image must form its top-left nh x nw block
    (what the Idefics2 processor's bottom/right padding produces); anything else is rejected."""
    m = pixel_attention_mask.unfold(1, patch, patch).unfold(2, patch, patch)
    pm = m.sum(dim=(-1, -2)) == patch * patch
    out = []
    for b in range(pm.shape[0]):
        nh, nw = int(pm[b, :, 0].sum()), int(pm[b, 0, :].sum())
        ref = torch.zeros_like(pm[b])
        ref[:nh, :nw] = True
        if nh == 0 or nw == 0 or not torch.equal(ref, pm[b]):
            raise ValueError("pixel_attention_mask must mark a top-left rectangle of whole patches per image")
        out.append((nh, nw))
    return out


def navit_position_ids(nh: int, nw: int, side: int) -> torch.Tensor:
    """Bucketised fractional coordinates of an nh x nw patch grid on the side x side position table
    (idefics_vision_tower.py:129-150; HF Idefics2VisionEmbeddings) — host index math."""
    boundaries = torch.arange(1 / side, 1.0, 1 / side)
    fh = torch.clamp(torch.arange(nh, dtype=torch.float32) * (1.0 / nh), max=1.0 - 1e-6)
    fw = torch.clamp(torch.arange(nw, dtype=torch.float32) * (1.0 / nw), max=1.0 - 1e-6)
    bh = torch.bucketize(fh, boundaries, right=True)
    bw = torch.bucketize(fw, boundaries, right=True)
    return (bh[:, None] * side + bw[None, :]).reshape(-1)


class Idefics2Engine:
    def __init__(self, cfg: Idefics2Config, weights: Idefics2Weights):
        self.cfg = cfg
        self.w = weights
        self.base = PrefillEngine(weights.base.cfg, weights.base)
        self.device = self.base.device

    # ------------------------------------------------------------------ vision tower (variable resolution)
    def vision(self, pixel_values: torch.Tensor, rects) -> (torch.Tensor, List[int]):
        """pixel_values [n, 3, H, W] (device), rects [(nh, nw)] -> (last_hidden_state packed [sum nh*nw, 1152], cu)."""
        v = self.cfg.vit
        L = _lib.lib()
        dims = self.w.base.vit_dims
        kpad = dims.patch_k_pad
        cu = [0]
        for nh, nw in rects:
            cu.append(cu[-1] + nh * nw)
        R = cu[-1]
        A = torch.empty((R, kpad), dtype=torch.bfloat16, device=self.device)
        pos = torch.cat([navit_position_ids(nh, nw, v.grid) for nh, nw in rects]).to(self.device, torch.long)
        for i, (nh, nw) in enumerate(rects):
            crop = pixel_values[i:i + 1, :, : nh * 14, : nw * 14].contiguous()
            rc = L.lp_im2col_patch14(_lib.ptr(crop), 1 if crop.dtype == torch.float32 else 0,
                                     _lib.ptr(A[cu[i]:cu[i + 1]]), kpad, 1, 3, nh * 14, nw * 14, _lib.stream_ptr())
            _lib.check(rc, "lp_im2col_patch14")
        pos_rows, _, _ = ops.embed_merge(pos, self.w.pos_table, None, -(2 ** 62), 1)  # gather table[pos]
        keep = self.w.base._keep
        patch_w = next(t for t in keep if t.data_ptr() == self.w.base.vit.patch_w)
        patch_b = next(t for t in keep if t.data_ptr() == self.w.base.vit.patch_b)
        x = ops.gemm(A, patch_w, bias=patch_b, residual=pos_rows[:R])
        out = torch.empty((R, v.hidden), dtype=torch.bfloat16, device=self.device)
        ws = self.base._workspace("vit_layers", L.lp_vit_layers_workspace_bytes(C.byref(dims), R))
        cu_c = (C.c_int * len(cu))(*cu)
        rc = L.lp_vit_layers(C.byref(dims), C.byref(self.w.base.vit), _lib.ptr(x), cu_c, len(rects), _lib.ptr(out),
                             out.stride(0), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "lp_vit_layers")
        return out, cu

    # ------------------------------------------------------------------ connector
    def connector(self, hidden: torch.Tensor, cu: List[int]) -> torch.Tensor:
        """packed [R, 1152] -> [n * 64, 4096]."""
        pc = self.cfg.perceiver
        w = self.w
        n = len(cu) - 1
        nl = pc.latents
        ctx = ops.gemm(ops.gemm(hidden, w.mp_gateup, act=_lib.ACT_SWIGLU), w.mp_down)  # [R, 4096]
        lat = w.latents.repeat(n, 1).contiguous()                                      # [n*64, 4096]
        R = ctx.shape[0]
        hs = torch.empty((R + n * nl, pc.hidden), dtype=torch.bfloat16, device=self.device)
        kvw = 2 * pc.kv_heads * pc.head_dim
        off = [cu[i] + i * nl for i in range(n + 1)]  # row offsets of each image's [context ; latents] block
        for L_ in w.layers:
            ln = ops.rmsnorm(lat, L_.lat_norm, pc.eps)
            for i in range(n):
                t = cu[i + 1] - cu[i]
                ops.rmsnorm(ctx[cu[i]:cu[i + 1]], L_.ctx_norm, pc.eps, out=hs[off[i]:off[i] + t])
                ops.rmsnorm(lat[i * nl:(i + 1) * nl], L_.lat_norm, pc.eps, out=hs[off[i] + t:off[i + 1]])
            q = ops.gemm(ln, L_.q_w)                       # [n*64, 1536]
            kv = ops.gemm(hs, L_.kv_w)                     # [R + n*64, 768] = K (4 x 96) | V (4 x 96)
            att = torch.empty_like(q)
            for i in range(n):
                tk = off[i + 1] - off[i]
                ops.attention_cross(q[i * nl:(i + 1) * nl], kv[off[i]:off[i + 1]], 1, nl, tk, pc.heads, pc.kv_heads,
                                    pc.head_dim, 0, pc.kv_heads * pc.head_dim, out=att[i * nl:(i + 1) * nl])
            lat = ops.gemm(att, L_.o_w, residual=lat)
            h = ops.rmsnorm(lat, L_.post_norm, pc.eps)
            lat = ops.gemm(ops.gemm(h, L_.gateup, act=_lib.ACT_SWIGLU), L_.down, residual=lat)
        return ops.rmsnorm(lat, w.final_norm, pc.eps)

    def image_features(self, pixel_values: torch.Tensor, pixel_attention_mask: Optional[torch.Tensor]):
        """[B, n, 3, H, W] (+ mask [B, n, H, W]) -> [n_real * 64, 4096]; all-zero padding images are dropped."""
        pv = pixel_values.reshape(-1, *pixel_values.shape[2:])
        real = (pv == 0).flatten(1).sum(1) != pv[0].numel()
        pv = pv[real]
        if pixel_attention_mask is None:
            pm = torch.ones((pv.shape[0], pv.shape[2], pv.shape[3]), dtype=torch.bool)
        else:
            pm = pixel_attention_mask.reshape(-1, *pixel_attention_mask.shape[2:])[real.cpu()].bool().cpu()
        rects = patch_rectangles(pm, self.cfg.vit.patch_size)
        pv = pv.to(self.device)
        if pv.dtype not in (torch.float32, torch.bfloat16):
            pv = pv.float()
        hidden, cu = self.vision(pv, rects)
        return self.connector(hidden, cu)

    @torch.no_grad()
    def prefill(self, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor],
                pixel_attention_mask: Optional[torch.Tensor], logits_mode: str = "all", want_kv: bool = False,
                kv_capacity: int = 0):
        """input_ids [1, L] with 64 image-token ids per image -> (logits, kv, S)."""
        if input_ids.shape[0] != 1:
            raise NotImplementedError("the reference eval path is B = 1 (evaluations/models/idefics2_multiimg.py:88-97)")
        ids = input_ids[0].to(self.device, torch.long)
        feats = None
        if pixel_values is not None:
            feats = self.image_features(pixel_values, pixel_attention_mask)
        n_slots = int((ids == self.cfg.image_token_id).sum())
        n_feat = 0 if feats is None else feats.shape[0]
        if n_slots != n_feat:
            raise ValueError(f"image tokens ({n_slots}) do not match image feature rows ({n_feat})")
        x, _ = self.base.merge(ids, feats, self.cfg.image_token_id, 1)
        S = ids.numel()
        pos = torch.arange(S, dtype=torch.int32, device=self.device)
        logits, kv = self.base.decoder(x, pos, [0, S], logits_mode, torch.float32, want_kv, S + kv_capacity)
        return logits, kv, S


class Idefics2ForConditionalGeneration:
    """What `AutoModelForVision2Seq.from_pretrained(...).cuda()` returns in the reference driver, natively backed."""

    def __init__(self, cfg: Idefics2Config, getter: Callable[[str], torch.Tensor], device="cuda:0"):
        self.config = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NativeError("Idefics2ForConditionalGeneration: CUDA only (no CPU fallback)")
        self.engine = Idefics2Engine(cfg, Idefics2Weights(cfg, getter, self.device))

    def eval(self):
        return self

    def cuda(self):
        return self

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, pixel_values=None, pixel_attention_mask=None, **_):
        logits, _, _ = self.engine.prefill(input_ids, pixel_values, pixel_attention_mask, "all")
        return SimpleNamespace(logits=logits[None], past_key_values=None)

    __call__ = forward

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, pixel_values=None, pixel_attention_mask=None,
                 max_new_tokens: int = 128, eos_token_id=None, **_):
        """Greedy; returns ids INCLUDING the prompt (the driver splits the decoded text on 'Assistant: ')."""
        eng = self.engine
        L = _lib.lib()
        eos = set() if eos_token_id is None else (set(eos_token_id) if isinstance(eos_token_id, (list, tuple))
                                                  else {int(eos_token_id)})
        logits, kv, S = eng.prefill(input_ids, pixel_values, pixel_attention_mask, "last", want_kv=max_new_tokens > 1,
                                    kv_capacity=max_new_tokens)
        base = eng.base
        if max_new_tokens <= 1:
            new = logits[..., : self.config.dec.vocab].argmax(-1).reshape(-1)[-1:].to(torch.long)
            return torch.cat([input_ids.to(base.device), new[None]], dim=1)
        from .decode import GreedyDecoder   # same device-resident greedy loop (CUDA-graph replay) as the LLaVA shim
        new = GreedyDecoder(base).continue_from(logits, kv, S, max_new_tokens, eos)
        return torch.cat([input_ids.to(base.device), new[None]], dim=1)
