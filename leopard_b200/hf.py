"""Drop-in for the HF-eval plugin surface of the reference (first row).

Same names and call conventions as evaluations/models/llava_multiimg_siglip_anyres.py:
    pixel_shuffle(x, scale_factor=2)                                   (:165-176)
    myLlavaMultiModalProjector(config).forward(image_features)         (:179-192)
    myLlavaForConditionalGeneration.from_pretrained(...).eval().to(dev)
        .forward(input_ids, pixel_values, attention_mask, ...) -> LlavaCausalLMOutputWithPast-like   (:201-361)
        .generate(input_ids, pixel_values=, attention_mask=, pad_token_id=, eos_token_id=, max_new_tokens=, use_cache=)
                                                                        (call site :448-452)
All arithmetic runs in the native library; tensors may arrive on the host (they are copied to the model's device, as
`.to(llava.device)` does in the reference driver) but there is no CPU compute path: without CUDA every call raises.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import _lib, ops
from .config import DecoderConfig, LeopardConfig, RopeScaling, VitConfig, leopard_llava
from .engine import PrefillEngine
from .native_weights import NativeWeights, load_hf


# ----------------------------------------------------------------------------------------------------------------
def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.NativeError(f"{what}: tensor is on {t.device}; leopard_b200 computes on CUDA only (no CPU fallback)")


def pixel_shuffle(x: torch.Tensor, scale_factor: int = 2) -> torch.Tensor:
    """[bsz, seq, dim] -> [bsz, seq/4, 4*dim]; 2x2 spatial neighbourhoods concatenated along channels
    (evaluations/models/llava_multiimg_siglip_anyres.py:165-176).  Native kernel: lp_pixel_shuffle."""
    if scale_factor != 2:
        raise NotImplementedError("Leopard uses scale_factor=2 only")
    _require_cuda(x, "pixel_shuffle")
    bsz, seq, dim = x.shape
    side = int(seq ** 0.5)
    y = ops.pixel_shuffle(x.to(torch.bfloat16).reshape(bsz * seq, dim).contiguous(), side)
    return y.view(bsz, seq // 4, dim * 4).to(x.dtype)


class myLlavaMultiModalProjector(nn.Module):
    """linear_2(act(linear_1(pixel_shuffle(x)))) with the reference's attribute names (…anyres.py:179-192).
    `config` needs .vision_config.hidden_size, .text_config.hidden_size and .projector_hidden_act."""

    def __init__(self, config):
        super().__init__()
        vh, th = config.vision_config.hidden_size, config.text_config.hidden_size
        self.linear_1 = nn.Linear(vh * 4, th, bias=True)
        self.act_name = getattr(config, "projector_hidden_act", "gelu")
        self.act = nn.GELU() if self.act_name == "gelu" else nn.GELU(approximate="tanh")  # attribute parity only
        self.linear_2 = nn.Linear(th, th, bias=True)
        self._native = None

    def _sync(self, device):
        ver = tuple(p._version for p in self.parameters()) + (str(device),)
        if self._native is None or self._native[0] != ver:
            bf = lambda t: t.detach().to(device, torch.bfloat16).contiguous()  # noqa: E731
            f32 = lambda t: t.detach().to(device).to(torch.bfloat16).float().contiguous()  # noqa: E731
            self._native = (ver, bf(self.linear_1.weight), f32(self.linear_1.bias), bf(self.linear_2.weight),
                            f32(self.linear_2.bias))
        return self._native[1:]

    def forward(self, image_features: torch.Tensor) -> torch.Tensor:
        _require_cuda(image_features, "myLlavaMultiModalProjector")
        w1, b1, w2, b2 = self._sync(image_features.device)
        n, seq, dim = image_features.shape
        x = ops.pixel_shuffle(image_features.to(torch.bfloat16).reshape(n * seq, dim).contiguous(), int(seq ** 0.5))
        act = _lib.ACT_GELU_ERF if self.act_name == "gelu" else _lib.ACT_GELU_TANH
        h = ops.gemm(x, w1, bias=b1, act=act)
        y = ops.gemm(h, w2, bias=b2)
        return y.view(n, seq // 4, -1).to(image_features.dtype)


@dataclass
class LlavaCausalLMOutputWithPast:
    """Field-compatible with transformers' LlavaCausalLMOutputWithPast (…anyres.py:351-361)."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[tuple] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None

    def __getitem__(self, i):
        return tuple(v for v in (self.loss, self.logits, self.past_key_values) if v is not None)[i]


class NativeKVCache(tuple):
    """`past_key_values` as the reference reads it — a tuple over layers of (key, value), each `[1, kv_heads, S,
    head_dim]` (…anyres.py:294 indexes `[0][0][:, :, :, 0]`) — whose entries are VIEWS of the native cache buffer
    `[layers, 2, capacity, kv_w]`.  Passing it back to `forward(input_ids=[1, 1], past_key_values=...)` lets the cached
    single-token branch (…anyres.py:291-320) append in place instead of re-importing the tensors."""

    def __new__(cls, kv: torch.Tensor, length: int, kv_heads: int, head_dim: int):
        views = tuple((kv[l, 0, :length].view(length, kv_heads, head_dim).permute(1, 0, 2)[None],
                       kv[l, 1, :length].view(length, kv_heads, head_dim).permute(1, 0, 2)[None])
                      for l in range(kv.shape[0]))
        self = super().__new__(cls, views)
        self.kv, self.length = kv, length
        return self

    def get_seq_length(self):
        return self.length


def _config_from_hf_json(d: dict) -> LeopardConfig:
    t, v = d.get("text_config", {}), d.get("vision_config", {})
    rs = t.get("rope_scaling")
    dec = DecoderConfig(hidden=t.get("hidden_size", 4096), ffn=t.get("intermediate_size", 14336),
                        heads=t.get("num_attention_heads", 32), kv_heads=t.get("num_key_value_heads", 8),
                        layers=t.get("num_hidden_layers", 32), vocab=t.get("vocab_size", 128256),
                        rope_theta=float(t.get("rope_theta", 500000.0)),
                        rope_scaling=None if not rs else RopeScaling(
                            factor=rs.get("factor", 8.0), low_freq_factor=rs.get("low_freq_factor", 1.0),
                            high_freq_factor=rs.get("high_freq_factor", 4.0),
                            original_max_position=rs.get("original_max_position_embeddings", 8192)),
                        eps=t.get("rms_norm_eps", 1e-5))
    vit = VitConfig(image_size=v.get("image_size", 364), patch_size=v.get("patch_size", 14),
                    hidden=v.get("hidden_size", 1152), ffn=v.get("intermediate_size", 4304),
                    heads=v.get("num_attention_heads", 16), layers=v.get("num_hidden_layers", 27),
                    eps=v.get("layer_norm_eps", 1e-6))
    return LeopardConfig(vit=vit, dec=dec, image_token_index=d.get("image_token_index", 128200),
                         projector_act=d.get("projector_hidden_act", "gelu"))


class _VisionTower:
    """`model.vision_tower(pixel_values).last_hidden_state` (…anyres.py:268,273)."""

    def __init__(self, owner):
        self._o = owner

    def __call__(self, pixel_values, **_):
        eng = self._o.engine
        out = eng.vit(pixel_values.to(eng.device), shuffle=False)
        n = pixel_values.shape[0]
        return SimpleNamespace(last_hidden_state=out.view(n, eng.cfg.vit.tokens, -1).to(pixel_values.dtype))


class _Projector:
    """`model.multi_modal_projector(selected_image_feature)` (…anyres.py:283)."""

    def __init__(self, owner):
        self._o = owner

    def __call__(self, image_features):
        eng = self._o.engine
        n, seq, dim = image_features.shape
        x = ops.pixel_shuffle(image_features.to(eng.device, torch.bfloat16).reshape(n * seq, dim).contiguous(),
                              int(seq ** 0.5))
        return eng.project(x).view(n, seq // 4, -1).to(image_features.dtype)


class _LanguageModel:
    """`model.language_model(inputs_embeds=..., position_ids=..., use_cache=...)` -> (logits, past) (…anyres.py:322-333)."""

    def __init__(self, owner):
        self._o = owner

    def __call__(self, inputs_embeds=None, position_ids=None, attention_mask=None, past_key_values=None,
                 use_cache=None, **_):
        if past_key_values is not None:
            raise NotImplementedError("cached decode goes through myLlavaForConditionalGeneration.generate")
        eng = self._o.engine
        if inputs_embeds.shape[0] != 1:
            raise NotImplementedError("the reference path is B = 1 (SURVEY.md App. A.1)")
        x = inputs_embeds[0].to(eng.device, torch.bfloat16).contiguous().clone()
        S = x.shape[0]
        pos = torch.arange(S, device=eng.device) if position_ids is None else position_ids.reshape(-1)
        logits, kv = eng.decoder(x, pos, [0, S], "all", torch.float32, want_kv=bool(use_cache))
        out = SimpleNamespace(logits=logits[None], past_key_values=self._o._legacy_cache(kv, S) if kv is not None else None,
                              hidden_states=None, attentions=None)
        return out


class _NativeParameters(nn.Module):
    """Holds the kernel-layout weight tensors as (frozen) nn.Parameters that SHARE storage with the pointers the native library
    reads, so `model.parameters()`, `named_parameters()`, parameter counting and hooks see the real weights."""

    def __init__(self, named: dict):
        super().__init__()
        self.names = {}
        for k, t in named.items():
            if not t.is_floating_point():
                continue
            attr = k.replace(".", "__")
            self.names[attr] = k
            self.register_parameter(attr, nn.Parameter(t, requires_grad=False))


def plan_padded_batch(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], image_id: int, rows_per_image: int):
    """Host plan of HF's `_merge_input_ids_with_image_features` for a padded batch (transformers 4.38-4.4x algorithm, SURVEY.md
    App. A.1 steps 1-3; the reference's forward calls it at …anyres.py:285-287).  Padding is removed instead of masked: the rows the
    attention mask keeps of batch row b form one packed sequence of S_b = kept ids + (rows_per_image - 1) per image placeholder,
    with positions 0 .. S_b - 1 (= cumsum(mask) - 1 of step 4), and land in the merged [B, max_len] layout at offset off_b:
    left padding -> right-aligned (new_pos += nb_image_pad), right padding -> left-aligned.
    Returns (max_len, left_padding, [(t0, t1, n_images, S_b, off_b), ...])."""
    B, L = input_ids.shape
    ids = input_ids.cpu()
    mask = torch.ones_like(ids, dtype=torch.bool) if attention_mask is None else attention_mask.cpu().bool()
    if mask.shape != ids.shape:
        raise ValueError(f"attention_mask {tuple(mask.shape)} does not match input_ids {tuple(ids.shape)}")
    n_img = (ids == image_id).sum(-1)
    max_len = int(n_img.max()) * (rows_per_image - 1) + L
    left_padding = bool(mask[:, -1].all())          # HF: not any(input_ids[:, -1] == pad_token_id)
    rows = []
    for b in range(B):
        keep = torch.nonzero(mask[b]).flatten()
        if keep.numel() == 0:
            raise ValueError(f"batch row {b} is all padding")
        t0, t1 = int(keep[0]), int(keep[-1]) + 1
        if t1 - t0 != keep.numel():
            raise NotImplementedError(f"batch row {b}: padding inside the prompt is outside what HF's merge supports")
        if bool(((ids[b] == image_id) & ~mask[b]).any()):
            raise ValueError(f"batch row {b}: image placeholder under the padding mask")
        if left_padding and t1 != L:
            raise NotImplementedError("mixed left / right padding in one batch")
        if not left_padding and t0 != 0:
            raise NotImplementedError("mixed left / right padding in one batch")
        nb = int(n_img[b])
        if not left_padding and nb != int(n_img.max()):
            # HF's step 3 drops the FIRST nb_image_pad empty positions of a row, which under right padding are real image rows:
            # its own output is misplaced there ("breaks batch generation"); refuse instead of reproducing it
            raise NotImplementedError("right-padded batches need the same number of images in every row (use left padding, as "
                                      "HF's batched generation does)")
        S_b = (t1 - t0) + nb * (rows_per_image - 1)
        rows.append((t0, t1, nb, S_b, max_len - S_b if left_padding else 0))
    return max_len, left_padding, rows


class myLlavaForConditionalGeneration(nn.Module):
    """Native-backed stand-in for the reference class of the same name (…anyres.py:195-361).  An `nn.Module` like the
    `LlavaForConditionalGeneration` it replaces: `parameters()` / `named_parameters()` enumerate the kernel-layout weights,
    `state_dict()` / `load_state_dict()` speak the reference's HF key layout (re-laid-out on the fly, loads are IN PLACE),
    `eval()` / `train()` / hooks / `__call__` behave as usual; `.to()` accepts CUDA devices only (no CPU fallback)."""

    def __init__(self, config: Union[LeopardConfig, dict, None] = None):
        super().__init__()
        if config is None:
            config = leopard_llava()
        if isinstance(config, dict):
            config = _config_from_hf_json(config)
        self.config = config
        self.engine: Optional[PrefillEngine] = None
        self._native: Optional[NativeWeights] = None
        self._pending = None  # HF-layout getter waiting for a CUDA device
        self.device = torch.device("cpu")
        self.dtype = torch.float32
        self.training = False
        self._precision = "bf16"
        self.native = None       # _NativeParameters once the weights are on a device
        self.vision_tower = _VisionTower(self)
        self.multi_modal_projector = _Projector(self)
        self.language_model = _LanguageModel(self)

    def _bind(self, native: NativeWeights):
        self._native, self.device = native, native.device
        self.engine = PrefillEngine(self.config, native)
        self.native = _NativeParameters(native.named)

    # ------------------------------------------------------------------ nn.Module surface over the native weights
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """HF-layout tensors (the reference checkpoint's keys; views / re-layouts of the live device tensors)."""
        if self._native is None:
            return {} if destination is None else destination
        from .native_weights import named_to_hf
        out = {prefix + k: v for k, v in named_to_hf(self.config, self._native.named).items()}
        if destination is not None:
            destination.update(out)
            return destination
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Load HF-layout tensors IN PLACE (device pointers, ctypes structs and captured graphs stay valid)."""
        get = lambda name: state_dict[name]  # noqa: E731
        if self._native is None:
            self._pending = get
            return nn.modules.module._IncompatibleKeys([], [])
        fresh = load_hf(self.config, get, self._native.device)
        for dst, src in zip(self._native._keep, fresh._keep):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise _lib.NativeError(f"load_state_dict: tensor mismatch {tuple(dst.shape)} vs {tuple(src.shape)}")
            dst.copy_(src)
        return nn.modules.module._IncompatibleKeys([], [])

    def half(self):
        return self      # the native path computes in bf16 with fp32 accumulation whatever the caller's dtype request

    float = bfloat16 = half

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_native(cls, cfg: LeopardConfig, native: NativeWeights):
        m = cls(cfg)
        m._bind(native)
        return m

    @classmethod
    def from_state_dict(cls, cfg: LeopardConfig, getter, device=None):
        """`getter(name)` returns HF-layout tensors (a dict's __getitem__, a SyntheticCheckpoint, ...)."""
        m = cls(cfg)
        m._pending = getter
        if device is not None:
            m.to(device)
        return m

    @classmethod
    def from_pretrained(cls, path, torch_dtype=None, **_):
        """Load an HF-layout checkpoint directory (config.json + *.safetensors / pytorch_model*.bin), the layout
        written by the reference converter (hf2megatron_llava.py:1026-1052,1453).  Weights stay on the host until
        `.to(device)`; `torch_dtype` is accepted for call compatibility (the native path computes in bf16)."""
        path = Path(path)
        cfg = _config_from_hf_json(json.loads((path / "config.json").read_text()))
        tensors = {}
        st = sorted(path.glob("*.safetensors"))
        if st:
            from safetensors.torch import load_file
            for f in st:
                tensors.update(load_file(str(f)))
        else:
            for f in sorted(path.glob("pytorch_model*.bin")):
                tensors.update(torch.load(f, map_location="cpu", weights_only=True))
        if not tensors:
            raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {path}")
        # transformers >= 4.5x nests the towers under `model.`; accept both spellings
        def get(name):
            for cand in (name, "model." + name, name.replace("language_model.model.", "model.language_model.")):
                if cand in tensors:
                    return tensors[cand]
            raise KeyError(name)
        m = cls.from_state_dict(cfg, get)
        if torch_dtype is not None:
            m.dtype = torch_dtype
        return m

    def set_precision(self, precision: str):
        """"bf16" (default): the fast path — bf16 activations, fp32 accumulation.  "fp32": PARITY MODE for forward() without
        a cache — the fp32-equivalent arithmetic of the reference's `torch_dtype=torch.float32` evaluation (…anyres.py:373) on
        the same tensor-core GEMM with three-term operand splitting (leopard_b200/parity.py); seconds per request, meant for
        checking logits against the reference, not for throughput.  generate() always decodes on the fast path."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self._precision = precision
        return self

    def to(self, device=None, *args, **kwargs):
        if device is None or isinstance(device, torch.dtype):      # .to(dtype): the native path's arithmetic type is fixed
            return self
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.NativeError("myLlavaForConditionalGeneration.to(): only CUDA devices are supported (no CPU fallback)")
        if self._native is None or self._native.device != device:
            if self._pending is None:
                raise _lib.NativeError("no weights to move: construct with from_pretrained/from_state_dict/from_native")
            self._bind(load_hf(self.config, self._pending, device))
        self.device = device
        return self

    def cuda(self, index=0):
        return self.to(f"cuda:{index}")

    def get_input_embeddings(self):
        eng = self._eng()

        def embed(input_ids):
            flat = input_ids.reshape(-1).to(eng.device)
            out, _ = eng.merge(flat, None, image_id=-(2 ** 62), rows_per_slot=1)
            return out.view(*input_ids.shape, -1)
        return embed

    def _eng(self) -> PrefillEngine:
        if self.engine is None:
            raise _lib.NativeError("model is not on a CUDA device yet: call .to('cuda:0') (no CPU fallback)")
        return self.engine

    def _legacy_cache(self, kv, S):
        """[layers, 2, cap, kv_w] -> tuple-of-tuples [layer][k|v] shaped [1, kv_heads, S, head_dim] (views), the legacy
        format the reference indexes directly (…anyres.py:294)."""
        d = self.config.dec
        return NativeKVCache(kv, int(S), d.kv_heads, d.head_dim)

    def _native_cache(self, past, room: int) -> "NativeKVCache":
        """Any legacy `past_key_values` -> a native cache with room for `room` more rows (copying when it is a foreign
        tuple of `[1, kv_heads, S, head_dim]` tensors, e.g. produced by transformers, or when capacity is exhausted)."""
        d = self.config.dec
        if isinstance(past, NativeKVCache) and past.kv.shape[2] >= past.length + room:
            return past
        S = int(past[0][0].shape[2])
        cap = -(-(S + max(room, 64)) // 256) * 256
        kv = torch.empty((d.layers, 2, cap, d.kv_heads * d.head_dim), dtype=torch.bfloat16, device=self.engine.device)
        for l in range(d.layers):
            for j in (0, 1):
                t = past[l][j].to(self.engine.device, torch.bfloat16)
                kv[l, j, :S] = t[0].permute(1, 0, 2).reshape(S, -1)
        return NativeKVCache(kv, S, d.kv_heads, d.head_dim)

    def _cached_step(self, input_ids, past):
        """The cached branch of the reference forward (…anyres.py:291-320) for B = 1 without padding: the new token's
        position is the number of cached rows; one native decode step over the cache."""
        import ctypes as C
        eng = self.engine
        L = _lib.lib()
        cache = self._native_cache(past, 1)
        tok = input_ids.reshape(-1)[-1:].to(eng.device, torch.long).contiguous()
        ws = eng._workspace("decode", L.lp_decode_workspace_bytes(C.byref(eng.w.dec_dims)))
        logits = torch.empty((eng.w.vocab_padded,), dtype=torch.float32, device=eng.device)
        rc = L.lp_decode_step(C.byref(eng.w.dec_dims), C.byref(eng.w.dec), _lib.ptr(eng.w.embed), eng.w.embed.stride(0),
                              _lib.ptr(tok), cache.length, _lib.ptr(cache.kv), cache.kv.shape[2], _lib.ptr(logits),
                              _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "lp_decode_step")
        d = self.config.dec
        return logits[: d.vocab].view(1, 1, -1), NativeKVCache(cache.kv, cache.length + 1, d.kv_heads, d.head_dim)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, input_ids=None, pixel_values=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, vision_feature_layer=None, vision_feature_select_strategy=None, labels=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        eng = self._eng()
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are not materialised by the fused path")
        if past_key_values is not None:
            if input_ids is None or input_ids.shape[0] != 1 or input_ids.shape[1] != 1:
                raise NotImplementedError("with past_key_values the reference feeds ONE new token (…anyres.py:291)")
            logits, past = self._cached_step(input_ids, past_key_values)
            if return_dict is False:
                return (logits, past)
            return LlavaCausalLMOutputWithPast(logits=logits, past_key_values=past)
        if inputs_embeds is not None:
            out = self.language_model(inputs_embeds=inputs_embeds, position_ids=position_ids, use_cache=use_cache)
            logits, past = out.logits, out.past_key_values
        else:
            if input_ids.dim() != 2:
                raise ValueError("input_ids must be [B, L]")
            if input_ids.shape[0] != 1 or (attention_mask is not None and not bool(attention_mask.all())):
                return self._forward_padded_batch(input_ids, pixel_values, attention_mask, labels, use_cache, return_dict)
            pix = None if pixel_values is None else pixel_values.to(eng.device, non_blocking=True)
            ids = input_ids.to(eng.device, non_blocking=True)
            if self._precision == "fp32" and not use_cache:
                from .parity import ParityForward
                logits = ParityForward(eng).prefill_logits(ids[0], None if pix is None else pix.float())[None]
                if return_dict is False:
                    return (logits,)
                return LlavaCausalLMOutputWithPast(logits=logits)
            # room for the cached single-token steps a caller may drive through forward(past_key_values=...)
            logits, kv, cu = eng.prefill([(input_ids[0], pix)], "all", torch.float32, want_kv=bool(use_cache), kv_capacity=128)
            logits = logits[None]
            past = self._legacy_cache(kv, cu[-1]) if kv is not None else None
        loss = None
        if labels is not None:
            # shift-by-one CE over the merged sequence; not on the prefill hot path (eval never passes labels)
            lab = labels.to(logits.device)
            if lab.shape[1] != logits.shape[1]:
                raise ValueError("labels must already be expanded to the merged length")
            per_tok = ops.cross_entropy(logits[0, :-1].contiguous(), lab[0, 1:].clamp(min=-1))   # lp_cross_entropy
            n_valid = (lab[0, 1:] >= 0).sum().clamp(min=1)
            loss = per_tok.sum() / n_valid
        if return_dict is False:
            return tuple(v for v in (loss, logits, past) if v is not None)
        return LlavaCausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=past)

    def _forward_padded_batch(self, input_ids, pixel_values, attention_mask, labels, use_cache, return_dict):
        """B > 1 and / or padded prompts: HF's batched merge (plan_padded_batch) on the packed engine — every batch row is one
        packed decoder sequence without its padding, all crops go through the tower in one batch, and the logits are scattered
        back into the merged [B, max_len, V] layout (zeros at padded rows, which HF fills with values nobody reads)."""
        eng = self._eng()
        if use_cache:
            raise NotImplementedError("use_cache with a padded batch: generate() drives batched decoding row by row")
        if labels is not None:
            raise NotImplementedError("labels with a padded batch")
        if self._precision == "fp32":
            raise NotImplementedError("parity mode is B = 1")
        cfg = self.config
        max_len, _, rows = plan_padded_batch(input_ids, attention_mask, cfg.image_token_index, cfg.tokens_per_crop)
        n_total = sum(r[2] for r in rows)
        n_given = 0 if pixel_values is None else int(pixel_values.shape[0])
        if n_total != n_given:
            raise ValueError(
                f"The input provided to the model are wrong. The number of image tokens is {n_total} while "
                f"the number of image given to the model is {n_given}. This prevents correct indexing and breaks batch generation.")
        pix = None if pixel_values is None else pixel_values.to(eng.device, non_blocking=True)
        samples, c0 = [], 0
        for b, (t0, t1, nb, _, _) in enumerate(rows):
            samples.append((input_ids[b, t0:t1], pix[c0:c0 + nb] if nb else None))
            c0 += nb
        packed, _, cu = eng.prefill(samples, "all", torch.float32, slots=[r[2] for r in rows])
        logits = torch.zeros((len(rows), max_len, packed.shape[-1]), dtype=packed.dtype, device=packed.device)
        for b, (_, _, _, S_b, off) in enumerate(rows):
            logits[b, off:off + S_b] = packed[cu[b]:cu[b + 1]]
        if return_dict is False:
            return (logits,)
        return LlavaCausalLMOutputWithPast(logits=logits)

    # ------------------------------------------------------------------ generate (greedy, as the reference calls it)
    @torch.no_grad()
    def generate(self, input_ids, pixel_values=None, attention_mask=None, pad_token_id=None, eos_token_id=None,
                 max_new_tokens: int = 128, use_cache: bool = True, **_):
        """Greedy decoding; returns LongTensor [1, L_in + n_new] (prompt included, un-expanded — the caller slices
        `[:, L_in:]`, …anyres.py:453-454).  Prefill emits only the last row's logits (lp_decoder_prefill LAST mode)."""
        eng = self._eng()
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, L]")
        eos = set()
        if eos_token_id is not None:
            eos = set(eos_token_id) if isinstance(eos_token_id, (list, tuple)) else {int(eos_token_id)}
        if input_ids.shape[0] != 1 or (attention_mask is not None and not bool(attention_mask.all())):
            return self._generate_padded_batch(input_ids, pixel_values, attention_mask, pad_token_id, eos, max_new_tokens)
        ids = input_ids.to(eng.device, non_blocking=True)
        pix = None if pixel_values is None else pixel_values.to(eng.device, non_blocking=True)
        from .decode import GreedyDecoder
        dec = GreedyDecoder(eng)
        # the ids the CALLER holds (normally host memory) plan the merge: counting placeholders there costs no device sync
        new_tokens = dec.run(input_ids[0], pix, max_new_tokens, eos)
        return torch.cat([ids, new_tokens[None]], dim=1)

    def _generate_padded_batch(self, input_ids, pixel_values, attention_mask, pad_token_id, eos, max_new_tokens):
        """Batched greedy generation the way GenerationMixin returns it: [B, L + longest continuation], rows that stopped early are
        filled with pad_token_id.  Every row is prefilled without its padding and decoded over its own KV cache (the decode kernels
        are single-sequence: rows run one after another)."""
        eng = self._eng()
        cfg = self.config
        _, _, rows = plan_padded_batch(input_ids, attention_mask, cfg.image_token_index, cfg.tokens_per_crop)
        if sum(r[2] for r in rows) != (0 if pixel_values is None else int(pixel_values.shape[0])):
            raise ValueError("The input provided to the model are wrong: image placeholders and pixel_values disagree")
        if pad_token_id is None:
            if len(eos) == 0:
                raise ValueError("batched generate needs pad_token_id (or an eos id to pad with)")
            pad_token_id = min(eos)
        pix = None if pixel_values is None else pixel_values.to(eng.device, non_blocking=True)
        from .decode import GreedyDecoder
        dec = GreedyDecoder(eng)
        outs, c0 = [], 0
        for b, (t0, t1, nb, _, _) in enumerate(rows):
            outs.append(dec.run(input_ids[b, t0:t1], pix[c0:c0 + nb] if nb else None, max_new_tokens, eos))
            c0 += nb
        n_new = max(o.numel() for o in outs)
        full = torch.full((len(rows), input_ids.shape[1] + n_new), int(pad_token_id), dtype=torch.long, device=eng.device)
        full[:, : input_ids.shape[1]] = input_ids.to(eng.device)
        for b, o in enumerate(outs):
            full[b, input_ids.shape[1]: input_ids.shape[1] + o.numel()] = o
        return full
