"""Checkpoint -> kernel-native device layout (what the lp_*_weights structs of include/leopard_b200.h point at).

Accepts the two layouts the reference produces:
  * HF (transformers <= 4.4x LLaVA keys, separate q/k/v and gate/up)  — `load_hf`
  * Megatron (`language_model/{embedding,vision_tower,mm_projector,encoder,output_layer}`, fused + interleaved)
    — `megatron_to_hf` first restores HF tensors exactly as the reference converter does
      (Pai-Megatron-Patch/toolkits/model_checkpoints_convertor/llava/hf2megatron_llava.py:1100-1460 and the layer code
       megatron_patch/model/llava/transformer.py:713-738 for the GQA group interleave).
Re-layout is pure data movement (cat / pad / permute on the device); no arithmetic on the path happens here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Dict

import torch

from . import _lib
from .config import LeopardConfig

V = "vision_tower.vision_model."
P = "multi_modal_projector."
LM = "language_model.model."

HEAD_STRIDE = 80       # SigLIP head_dim 72 zero-padded to 5 x 16 (UMMA K granularity)
PATCH_K_PAD = 592      # 588 -> next multiple of 8 elements (TMA needs 16-byte row strides)


def rope_inv_freq(dec) -> torch.Tensor:
    """Host-side RoPE inverse frequencies incl. the llama3.1 rescale
    (megatron/core/models/common/embeddings/rotary_pos_embedding.py:48-83,112-123)."""
    d = dec.head_dim
    inv = 1.0 / (dec.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    rs = dec.rope_scaling
    if rs is None:
        return inv
    wavelen = 2 * math.pi / inv
    low_wl = rs.original_max_position / rs.low_freq_factor
    high_wl = rs.original_max_position / rs.high_freq_factor
    smooth = (rs.original_max_position / wavelen - rs.low_freq_factor) / (rs.high_freq_factor - rs.low_freq_factor)
    mid = (1 - smooth) * inv / rs.factor + smooth * inv
    out = torch.where(wavelen > low_wl, inv / rs.factor, torch.where(wavelen < high_wl, inv, mid))
    return out.float()


class NativeWeights:
    """Device tensors in kernel layout + the ctypes structs handed to lp_vit_forward / lp_project / lp_decoder_prefill."""

    def __init__(self, cfg: LeopardConfig, device):
        self.cfg = cfg
        self.device = torch.device(device)
        self._keep = []          # every device tensor referenced by a struct
        self.named = {}          # name -> device tensor, for the tensors the SFT step updates in place (train.py)
        self.folded = {}         # gain-folded matrices / column sums / folded biases (refresh_folded)
        self.folded_dirty = set()   # parts ("vit", "dec") whose base weights changed since the last refresh
        self.embed = None        # [vocab, hidden] bf16
        self.vocab_padded = (cfg.dec.vocab + 7) // 8 * 8
        v, d = cfg.vit, cfg.dec
        self.vit_dims = _lib.VitDims(v.layers, v.hidden, v.ffn, v.heads, v.head_dim, HEAD_STRIDE, v.grid, v.tokens,
                                     v.patch_k, PATCH_K_PAD, v.eps)
        self.dec_dims = _lib.DecDims(d.layers, d.hidden, d.ffn, d.heads, d.kv_heads, d.head_dim, self.vocab_padded,
                                     d.eps)
        self.vit_layers = (_lib.VitLayerWeights * max(v.layers, 1))()
        self.dec_layers = (_lib.DecLayerWeights * max(d.layers, 1))()
        self.vit = _lib.VitWeights()
        self.proj = _lib.ProjWeights()
        self.dec = _lib.DecWeights()

    # -- helpers
    def _bf(self, t: torch.Tensor, name: str = None) -> torch.Tensor:
        t = t.to(device=self.device, dtype=torch.bfloat16).contiguous()
        self._keep.append(t)
        if name is not None:
            self.named[name] = t
        return t

    def _f32(self, t: torch.Tensor, name: str = None) -> torch.Tensor:
        # vectors are kept as fp32 holding the checkpoint's (bf16-rounded) values
        t = t.to(device=self.device).to(torch.bfloat16).float().contiguous()
        self._keep.append(t)
        if name is not None:
            self.named[name] = t
        return t

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._keep)

    # -- gain-folded matrices for the fused-norm layer loops (lp_gemm_bf16_ex; include/leopard_b200.h)
    def _fold_set(self, key: str, value: torch.Tensor) -> torch.Tensor:
        cur = self.folded.get(key)
        if cur is None:
            cur = value.contiguous()
            self.folded[key] = cur
            self._keep.append(cur)
        else:
            cur.copy_(value)
        return cur

    @torch.no_grad()
    def refresh_folded(self, parts=("vit", "dec")):
        """(Re)derive, from the live tensors, what the fused layer loops read instead of running norm kernels:
        tower   qkv_wf = bf16(qkv_w * ln1_g), qkv_c1 = rowsum(qkv_wf), qkv_c2 = qkv_w ln1_b + qkv_b (fc1 likewise with ln2);
        decoder qkv_wf = bf16(qkv_w * ln1_g), gateup_wf = bf16(gateup_w * ln2_g).
        Pure weight preparation (the reference's TELayerNormColumnParallelLinear keeps LN and GEMM weights together the same
        way, clip_encoder.py:51-72).  First call allocates and publishes the pointers; later calls update in place, so the
        ctypes structs stay valid (call after the base weights changed, e.g. an optimizer step)."""
        n, v, d = self.named, self.cfg.vit, self.cfg.dec
        bf = torch.bfloat16
        if "vit" in parts:
            for i in range(v.layers):
                L = self.vit_layers[i]
                for mat, ln, tag in (("qkv", "ln1", "qkv"), ("fc1", "ln2", "fc1")):
                    w, b = n[f"vit.{i}.{mat}_w"], n[f"vit.{i}.{mat}_b"]
                    g, beta = n[f"vit.{i}.{ln}_g"], n[f"vit.{i}.{ln}_b"]
                    wf = self._fold_set(f"vit.{i}.{tag}_wf", (w.float() * g[None, :]).to(bf))
                    c1 = self._fold_set(f"vit.{i}.{tag}_c1", wf.float().sum(1))
                    c2 = self._fold_set(f"vit.{i}.{tag}_c2", w.float() @ beta + b)
                    setattr(L, f"{tag}_wf", self._p(wf))
                    setattr(L, f"{tag}_c1", self._p(c1))
                    setattr(L, f"{tag}_c2", self._p(c2))
            self.folded_dirty.discard("vit")
        if "dec" in parts:
            for i in range(d.layers):
                L = self.dec_layers[i]
                L.qkv_wf = self._p(self._fold_set(f"dec.{i}.qkv_wf", (n[f"dec.{i}.qkv_w"].float() * n[f"dec.{i}.ln1_g"][None, :]).to(bf)))
                L.gateup_wf = self._p(self._fold_set(f"dec.{i}.gateup_wf",
                                                    (n[f"dec.{i}.gateup_w"].float() * n[f"dec.{i}.ln2_g"][None, :]).to(bf)))
            self.folded_dirty.discard("dec")

    def sync_folded(self, part: str):
        """Called by the engine before a stage runs: refresh the folded copies of `part` if the trainer touched its weights."""
        if part in self.folded_dirty:
            self.refresh_folded((part,))


def load_hf(cfg: LeopardConfig, get: Callable[[str], torch.Tensor], device="cuda",
            with_projector: bool = True, fold: bool = True) -> NativeWeights:
    """Build NativeWeights from HF-layout tensors; `get(name)` returns one tensor (any float dtype, any device)."""
    nw = NativeWeights(cfg, device)
    v, d = cfg.vit, cfg.dec
    dev = nw.device

    # ---- SigLIP tower
    pw = get(V + "embeddings.patch_embedding.weight").to(dev).reshape(v.hidden, v.patch_k)
    pw = torch.nn.functional.pad(pw, (0, PATCH_K_PAD - v.patch_k))
    nw.vit.patch_w = nw._p(nw._bf(pw, "vit.patch_w"))
    nw.vit.patch_b = nw._p(nw._f32(get(V + "embeddings.patch_embedding.bias"), "vit.patch_b"))
    nw.vit.pos = nw._p(nw._bf(get(V + "embeddings.position_embedding.weight"), "vit.pos"))
    pad_rows = HEAD_STRIDE - v.head_dim

    def pad_heads(w):  # [heads*72, ...] -> [heads*80, ...] with zero rows after each head
        w = w.to(dev)
        shp = w.shape
        w = w.reshape(v.heads, v.head_dim, -1)
        w = torch.nn.functional.pad(w, (0, 0, 0, pad_rows))
        return w.reshape(v.heads * HEAD_STRIDE, *shp[1:])

    for i in range(v.layers):
        b = f"{V}encoder.layers.{i}."
        L = nw.vit_layers[i]
        L.ln1_g = nw._p(nw._f32(get(b + "layer_norm1.weight"), f"vit.{i}.ln1_g"))
        L.ln1_b = nw._p(nw._f32(get(b + "layer_norm1.bias"), f"vit.{i}.ln1_b"))
        L.ln2_g = nw._p(nw._f32(get(b + "layer_norm2.weight"), f"vit.{i}.ln2_g"))
        L.ln2_b = nw._p(nw._f32(get(b + "layer_norm2.bias"), f"vit.{i}.ln2_b"))
        qkv_w = torch.cat([pad_heads(get(b + f"self_attn.{n}_proj.weight")) for n in "qkv"], 0)
        qkv_b = torch.cat([pad_heads(get(b + f"self_attn.{n}_proj.bias")) for n in "qkv"], 0)
        L.qkv_w = nw._p(nw._bf(qkv_w, f"vit.{i}.qkv_w"))
        L.qkv_b = nw._p(nw._f32(qkv_b, f"vit.{i}.qkv_b"))
        L.out_w = nw._p(nw._bf(get(b + "self_attn.out_proj.weight"), f"vit.{i}.out_w"))
        L.out_b = nw._p(nw._f32(get(b + "self_attn.out_proj.bias"), f"vit.{i}.out_b"))
        L.fc1_w = nw._p(nw._bf(get(b + "mlp.fc1.weight"), f"vit.{i}.fc1_w"))
        L.fc1_b = nw._p(nw._f32(get(b + "mlp.fc1.bias"), f"vit.{i}.fc1_b"))
        L.fc2_w = nw._p(nw._bf(get(b + "mlp.fc2.weight"), f"vit.{i}.fc2_w"))
        L.fc2_b = nw._p(nw._f32(get(b + "mlp.fc2.bias"), f"vit.{i}.fc2_b"))
    nw.vit.layers = C.cast(nw.vit_layers, C.POINTER(_lib.VitLayerWeights))
    nw.vit.post_g = nw._p(nw._f32(get(V + "post_layernorm.weight"), "vit.post_g"))
    nw.vit.post_b = nw._p(nw._f32(get(V + "post_layernorm.bias"), "vit.post_b"))

    # ---- projector (Leopard-LLaVA only; Idefics2 has its own connector, see idefics2.py)
    if with_projector:
        nw.proj.w1 = nw._p(nw._bf(get(P + "linear_1.weight"), "proj.w1"))
        nw.proj.b1 = nw._p(nw._f32(get(P + "linear_1.bias"), "proj.b1"))
        nw.proj.w2 = nw._p(nw._bf(get(P + "linear_2.weight"), "proj.w2"))
        nw.proj.b2 = nw._p(nw._f32(get(P + "linear_2.bias"), "proj.b2"))
        nw.proj.act = _lib.ACT_GELU_ERF if cfg.projector_act == "gelu" else _lib.ACT_GELU_TANH

    # ---- decoder
    nw.embed = nw._bf(get(LM + "embed_tokens.weight"), "embed")
    for i in range(d.layers):
        b = f"{LM}layers.{i}."
        L = nw.dec_layers[i]
        L.ln1_g = nw._p(nw._f32(get(b + "input_layernorm.weight"), f"dec.{i}.ln1_g"))
        L.ln2_g = nw._p(nw._f32(get(b + "post_attention_layernorm.weight"), f"dec.{i}.ln2_g"))
        qkv = torch.cat([get(b + f"self_attn.{n}_proj.weight").to(dev) for n in "qkv"], 0)
        L.qkv_w = nw._p(nw._bf(qkv, f"dec.{i}.qkv_w"))
        del qkv
        L.o_w = nw._p(nw._bf(get(b + "self_attn.o_proj.weight"), f"dec.{i}.o_w"))
        g = get(b + "mlp.gate_proj.weight").to(dev).reshape(d.ffn // 32, 1, 32, d.hidden)
        u = get(b + "mlp.up_proj.weight").to(dev).reshape(d.ffn // 32, 1, 32, d.hidden)
        L.gateup_w = nw._p(nw._bf(torch.cat([g, u], 1).reshape(2 * d.ffn, d.hidden), f"dec.{i}.gateup_w"))
        del g, u
        L.down_w = nw._p(nw._bf(get(b + "mlp.down_proj.weight"), f"dec.{i}.down_w"))
    nw.dec.layers = C.cast(nw.dec_layers, C.POINTER(_lib.DecLayerWeights))
    nw.dec.final_g = nw._p(nw._f32(get(LM + "norm.weight"), "dec.final_g"))
    head = get("language_model.lm_head.weight").to(dev)
    if nw.vocab_padded != d.vocab:
        head = torch.nn.functional.pad(head, (0, 0, 0, nw.vocab_padded - d.vocab))
    nw.dec.lm_head = nw._p(nw._bf(head, "dec.lm_head"))
    inv = rope_inv_freq(d).to(dev).contiguous()
    nw._keep.append(inv)
    nw.named["dec.inv_freq"] = inv
    nw.dec.inv_freq = nw._p(inv)
    if fold:
        nw.refresh_folded()
    return nw


def named_to_hf(cfg: LeopardConfig, named: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of load_hf for the tensors the SFT step updates (NativeWeights.named): kernel layout -> HF-named tensors
    (split the fused qkv rows, undo the 32-row gate/up interleave, drop the vocabulary padding).  Pure data movement;
    returns views / reshapes of the live device tensors, so a checkpoint written from them holds the trained values."""
    d, v = cfg.dec, cfg.vit
    out: Dict[str, torch.Tensor] = {}
    q_w, kv_w = d.heads * d.head_dim, d.kv_heads * d.head_dim
    for name, t in named.items():
        if name.startswith("vit."):
            parts = name.split(".")
            if len(parts) == 2:
                key = {"patch_w": "embeddings.patch_embedding.weight", "patch_b": "embeddings.patch_embedding.bias",
                       "pos": "embeddings.position_embedding.weight", "post_g": "post_layernorm.weight",
                       "post_b": "post_layernorm.bias"}[parts[1]]
                if parts[1] == "patch_w":       # drop the K padding, back to the conv kernel shape
                    t = t[:, : v.patch_k].reshape(v.hidden, 3, v.patch_size, v.patch_size)
                out[V + key] = t
                continue
            b = f"{V}encoder.layers.{parts[1]}."
            kind = parts[2]
            if kind in ("qkv_w", "qkv_b"):      # [3, heads, 80 (72 + zero pad), ...] -> three [heads * 72, ...] tensors
                g = t.reshape(3, v.heads, HEAD_STRIDE, *t.shape[1:])[:, :, : v.head_dim]
                for j, n in enumerate("qkv"):
                    out[b + f"self_attn.{n}_proj." + ("weight" if kind == "qkv_w" else "bias")] = g[j].reshape(
                        v.hidden, *t.shape[1:])
            else:
                out[b + {"ln1_g": "layer_norm1.weight", "ln1_b": "layer_norm1.bias", "ln2_g": "layer_norm2.weight",
                         "ln2_b": "layer_norm2.bias", "out_w": "self_attn.out_proj.weight", "out_b": "self_attn.out_proj.bias",
                         "fc1_w": "mlp.fc1.weight", "fc1_b": "mlp.fc1.bias", "fc2_w": "mlp.fc2.weight",
                         "fc2_b": "mlp.fc2.bias"}[kind]] = t
        elif name == "embed":
            out[LM + "embed_tokens.weight"] = t
        elif name == "dec.inv_freq":
            continue
        elif name == "dec.final_g":
            out[LM + "norm.weight"] = t
        elif name == "dec.lm_head":
            out["language_model.lm_head.weight"] = t[: d.vocab]
        elif name.startswith("proj."):
            out[P + {"w1": "linear_1.weight", "b1": "linear_1.bias", "w2": "linear_2.weight", "b2": "linear_2.bias"}[
                name[5:]]] = t
        elif name.startswith("dec.") and name.split(".")[1].isdigit():
            _, i, kind = name.split(".")
            b = f"{LM}layers.{i}."
            if kind == "qkv_w":
                out[b + "self_attn.q_proj.weight"] = t[:q_w]
                out[b + "self_attn.k_proj.weight"] = t[q_w: q_w + kv_w]
                out[b + "self_attn.v_proj.weight"] = t[q_w + kv_w:]
            elif kind == "gateup_w":
                g = t.reshape(d.ffn // 32, 2, 32, d.hidden)
                out[b + "mlp.gate_proj.weight"] = g[:, 0].reshape(d.ffn, d.hidden)
                out[b + "mlp.up_proj.weight"] = g[:, 1].reshape(d.ffn, d.hidden)
            else:
                out[b + {"o_w": "self_attn.o_proj.weight", "down_w": "mlp.down_proj.weight",
                         "ln1_g": "input_layernorm.weight", "ln2_g": "post_attention_layernorm.weight"}[kind]] = t
    return out


def random_native(cfg: LeopardConfig, device="cuda", seed: int = 1234) -> NativeWeights:
    """Benchmark-only initialisation: same distributions as weights.synth_tensor but drawn ON the device (fast);
    values differ from the CPU-seeded checkpoint, so this is for throughput runs, never for parity."""
    from .weights import tensor_specs
    g = torch.Generator(device=device).manual_seed(seed)
    specs = {n: (s, k) for n, s, k in tensor_specs(cfg)}

    def get(name):
        shape, kind = specs[name]
        if kind in ("linear", "conv"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            return (torch.randn(shape, device=device, generator=g) * fan_in ** -0.5).to(torch.bfloat16)
        if kind == "norm":
            return (1.0 + torch.randn(shape, device=device, generator=g) * 0.02).to(torch.bfloat16)
        scale = 0.05 if kind == "embed" else 0.02
        return (torch.randn(shape, device=device, generator=g) * scale).to(torch.bfloat16)

    return load_hf(cfg, get, device)


# ----------------------------------------------------------------------------------------------------------------
# Megatron checkpoint layout -> HF layout
# ----------------------------------------------------------------------------------------------------------------
def megatron_to_hf(cfg: LeopardConfig, lm: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`lm` = flattened `state['model']['language_model']` of a TP=1 Megatron checkpoint with keys such as
    'encoder.layers.0.self_attention.query_key_value.weight'.  Returns HF-named tensors."""
    v, d = cfg.vit, cfg.dec
    out: Dict[str, torch.Tensor] = {}
    out[LM + "embed_tokens.weight"] = lm["embedding.word_embeddings.weight"][: d.vocab]
    out["language_model.lm_head.weight"] = lm["output_layer.weight"][: d.vocab]
    fin = "encoder.final_norm.weight" if "encoder.final_norm.weight" in lm else "encoder.final_layernorm.weight"
    out[LM + "norm.weight"] = lm[fin]
    rep = d.heads // d.kv_heads
    hd = d.head_dim
    for i in range(d.layers):
        s, t = f"encoder.layers.{i}.", f"{LM}layers.{i}."
        out[t + "input_layernorm.weight"] = lm[s + "input_norm.weight"]
        out[t + "post_attention_layernorm.weight"] = lm[s + "post_attention_norm.weight"]
        # fused rows per kv group: [rep q heads | k | v] x kv_heads (megatron_patch/model/llava/transformer.py:713-738)
        qkv = lm[s + "self_attention.query_key_value.weight"].reshape(d.kv_heads, (rep + 2) * hd, d.hidden)
        out[t + "self_attn.q_proj.weight"] = qkv[:, : rep * hd].reshape(d.heads * hd, d.hidden)
        out[t + "self_attn.k_proj.weight"] = qkv[:, rep * hd: (rep + 1) * hd].reshape(d.kv_heads * hd, d.hidden)
        out[t + "self_attn.v_proj.weight"] = qkv[:, (rep + 1) * hd:].reshape(d.kv_heads * hd, d.hidden)
        out[t + "self_attn.o_proj.weight"] = lm[s + "self_attention.dense.weight"]
        h4 = lm[s + "mlp.dense_h_to_4h.weight"]  # [gate ; up] (hf2megatron_llava.py:1287-1295)
        out[t + "mlp.gate_proj.weight"], out[t + "mlp.up_proj.weight"] = h4[: d.ffn], h4[d.ffn:]
        out[t + "mlp.down_proj.weight"] = lm[s + "mlp.dense_4h_to_h.weight"]
    out[P + "linear_1.weight"] = lm["mm_projector.encoder.linear_fc1.weight"]
    out[P + "linear_1.bias"] = lm["mm_projector.encoder.linear_fc1.bias"]
    out[P + "linear_2.weight"] = lm["mm_projector.encoder.linear_fc2.weight"]
    out[P + "linear_2.bias"] = lm["mm_projector.encoder.linear_fc2.bias"]
    out[V + "embeddings.patch_embedding.weight"] = lm["vision_tower.conv1.weight"]
    out[V + "embeddings.patch_embedding.bias"] = lm["vision_tower.conv1.bias"]
    out[V + "embeddings.position_embedding.weight"] = lm["vision_tower.position_embeddings.weight"]
    out[V + "post_layernorm.weight"] = lm["vision_tower.ln_post.weight"]
    out[V + "post_layernorm.bias"] = lm["vision_tower.ln_post.bias"]
    for i in range(v.layers):
        s, t = f"vision_tower.transformer.layers.{i}.", f"{V}encoder.layers.{i}."
        # rows interleaved [head, (q,k,v), 72] (hf2megatron_llava.py:1355-1374)
        w = lm[s + "self_attention.linear_qkv.weight"].reshape(v.heads, 3, v.head_dim, v.hidden)
        bb = lm[s + "self_attention.linear_qkv.bias"].reshape(v.heads, 3, v.head_dim)
        for j, n in enumerate("qkv"):
            out[t + f"self_attn.{n}_proj.weight"] = w[:, j].reshape(v.hidden, v.hidden)
            out[t + f"self_attn.{n}_proj.bias"] = bb[:, j].reshape(v.hidden)
        out[t + "layer_norm1.weight"] = lm[s + "self_attention.linear_qkv.layer_norm_weight"]
        out[t + "layer_norm1.bias"] = lm[s + "self_attention.linear_qkv.layer_norm_bias"]
        out[t + "self_attn.out_proj.weight"] = lm[s + "self_attention.linear_proj.weight"]
        out[t + "self_attn.out_proj.bias"] = lm[s + "self_attention.linear_proj.bias"]
        out[t + "mlp.fc1.weight"] = lm[s + "mlp.linear_fc1.weight"]
        out[t + "mlp.fc1.bias"] = lm[s + "mlp.linear_fc1.bias"]
        out[t + "layer_norm2.weight"] = lm[s + "mlp.linear_fc1.layer_norm_weight"]
        out[t + "layer_norm2.bias"] = lm[s + "mlp.linear_fc1.layer_norm_bias"]
        out[t + "mlp.fc2.weight"] = lm[s + "mlp.linear_fc2.weight"]
        out[t + "mlp.fc2.bias"] = lm[s + "mlp.linear_fc2.bias"]
    return out


def hf_to_megatron(cfg: LeopardConfig, get: Callable[[str], torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of megatron_to_hf (used by tests and by the Megatron shim's synthetic constructor)."""
    v, d = cfg.vit, cfg.dec
    lm: Dict[str, torch.Tensor] = {}
    lm["embedding.word_embeddings.weight"] = get(LM + "embed_tokens.weight")
    lm["output_layer.weight"] = get("language_model.lm_head.weight")
    lm["encoder.final_norm.weight"] = get(LM + "norm.weight")
    rep, hd = d.heads // d.kv_heads, d.head_dim
    for i in range(d.layers):
        s, t = f"encoder.layers.{i}.", f"{LM}layers.{i}."
        lm[s + "input_norm.weight"] = get(t + "input_layernorm.weight")
        lm[s + "post_attention_norm.weight"] = get(t + "post_attention_layernorm.weight")
        q = get(t + "self_attn.q_proj.weight").reshape(d.kv_heads, rep * hd, d.hidden)
        k = get(t + "self_attn.k_proj.weight").reshape(d.kv_heads, hd, d.hidden)
        vv = get(t + "self_attn.v_proj.weight").reshape(d.kv_heads, hd, d.hidden)
        lm[s + "self_attention.query_key_value.weight"] = torch.cat([q, k, vv], 1).reshape(-1, d.hidden)
        lm[s + "self_attention.dense.weight"] = get(t + "self_attn.o_proj.weight")
        lm[s + "mlp.dense_h_to_4h.weight"] = torch.cat([get(t + "mlp.gate_proj.weight"), get(t + "mlp.up_proj.weight")], 0)
        lm[s + "mlp.dense_4h_to_h.weight"] = get(t + "mlp.down_proj.weight")
    lm["mm_projector.encoder.linear_fc1.weight"] = get(P + "linear_1.weight")
    lm["mm_projector.encoder.linear_fc1.bias"] = get(P + "linear_1.bias")
    lm["mm_projector.encoder.linear_fc2.weight"] = get(P + "linear_2.weight")
    lm["mm_projector.encoder.linear_fc2.bias"] = get(P + "linear_2.bias")
    lm["vision_tower.conv1.weight"] = get(V + "embeddings.patch_embedding.weight")
    lm["vision_tower.conv1.bias"] = get(V + "embeddings.patch_embedding.bias")
    lm["vision_tower.position_embeddings.weight"] = get(V + "embeddings.position_embedding.weight")
    lm["vision_tower.ln_post.weight"] = get(V + "post_layernorm.weight")
    lm["vision_tower.ln_post.bias"] = get(V + "post_layernorm.bias")
    for i in range(v.layers):
        s, t = f"vision_tower.transformer.layers.{i}.", f"{V}encoder.layers.{i}."
        ws = [get(t + f"self_attn.{n}_proj.weight").reshape(v.heads, 1, v.head_dim, v.hidden) for n in "qkv"]
        bs = [get(t + f"self_attn.{n}_proj.bias").reshape(v.heads, 1, v.head_dim) for n in "qkv"]
        lm[s + "self_attention.linear_qkv.weight"] = torch.cat(ws, 1).reshape(3 * v.hidden, v.hidden)
        lm[s + "self_attention.linear_qkv.bias"] = torch.cat(bs, 1).reshape(3 * v.hidden)
        lm[s + "self_attention.linear_qkv.layer_norm_weight"] = get(t + "layer_norm1.weight")
        lm[s + "self_attention.linear_qkv.layer_norm_bias"] = get(t + "layer_norm1.bias")
        lm[s + "self_attention.linear_proj.weight"] = get(t + "self_attn.out_proj.weight")
        lm[s + "self_attention.linear_proj.bias"] = get(t + "self_attn.out_proj.bias")
        lm[s + "mlp.linear_fc1.weight"] = get(t + "mlp.fc1.weight")
        lm[s + "mlp.linear_fc1.bias"] = get(t + "mlp.fc1.bias")
        lm[s + "mlp.linear_fc1.layer_norm_weight"] = get(t + "layer_norm2.weight")
        lm[s + "mlp.linear_fc1.layer_norm_bias"] = get(t + "layer_norm2.bias")
        lm[s + "mlp.linear_fc2.weight"] = get(t + "mlp.fc2.weight")
        lm[s + "mlp.linear_fc2.bias"] = get(t + "mlp.fc2.bias")
    return lm
