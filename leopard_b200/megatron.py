"""Drop-in for the Megatron plugin surface of the reference (second row): forward / prefill, and the training
iteration's entry points (mm_collate_fn, loss_func, train_step over leopard_b200.train.SFTTrainer).

    model_provider(pre_process=True, post_process=True) -> model         (Pai-Megatron-Patch/examples/llava/
                                                                          pretrain_megatron_llava.py:37-50)
    model(tokens, position_ids, attention_mask, images=images) -> logits [b, s, V]
                                                                         (forward_step :157-158; GPTModel.forward
                                                                          megatron_patch/model/llava/gpt_model.py:90-113)
Conventions kept from the reference: image slots are ids == -200, one per feature ROW (169 per crop), crops ordered
tiles + [global]; ids -100 (padding) read embedding row 0 (vlm_model.py:226); `images` with last dim <= 100 is the
"no image" dummy (vlm_model.py:518; dummy = zeros(1, 10), pretrain_megatron_llava.py:188); a feature/slot count
mismatch raises AssertionError with the reference's message (vlm_model.py:528-529).  The attention mask is ignored
exactly like the flash-attention path of the reference (causal only).  forward(labels=...) returns the per-token loss
[b, s] (lp_cross_entropy).  There is no autograd graph: the training iteration (§8f-1, first implementation — DESIGN.md §6b)
is `train_step`, which hands the batch to SFTTrainer.step (forward, loss_func, backward, reduce, clip, AdamW on native
kernels).  Checkpoint surface: state_dict_for_save_checkpoint() / load_state_dict() speak the reference's nested Megatron
layout and export the live (trained) tensors.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib
from .config import LeopardConfig, leopard_llava
from .engine import PrefillEngine
from .native_weights import NativeWeights, hf_to_megatron, load_hf, megatron_to_hf, named_to_hf

IMAGE_TOKEN_INDEX = -200
IGNORE_INDEX = -100

_provider_state = {"cfg": None, "getter": None, "device": None}

_LM_KEY = "language_model"   # GPTModel._language_model_key (megatron_patch/model/llava/gpt_model.py:115-123)


def nest_language_model_state(flat: dict) -> dict:
    """Flat Megatron keys ('encoder.layers.0.mlp.dense_h_to_4h.weight', ...) -> the nested dict the reference saves under
    state['model']['language_model'] (TransformerLanguageModel.state_dict_for_save_checkpoint,
    megatron_patch/model/llava/vlm_model.py:684-711): one sub-dict per component with that component's own state_dict
    keys — `embedding` is nested once more ({'word_embeddings': {'weight': ...}}, as Embedding saves itself), the SigLIP
    tower's keys carry no 'vision_tower.' prefix (the converter re-adds it, hf2megatron_llava.py:1310-1315)."""
    out: dict = {}
    for k, v in flat.items():
        top, rest = k.split(".", 1)
        if top == "embedding":
            mod, leaf = rest.split(".", 1)
            out.setdefault(top, {}).setdefault(mod, {})[leaf] = v
        else:
            out.setdefault(top, {})[rest] = v
    return out


def flatten_language_model_state(state: dict) -> dict:
    """Inverse of nest_language_model_state; accepts the full {'language_model': {...}} checkpoint entry, the inner
    nested dict, or an already flat dict.  TransformerEngine `_extra_state` entries are dropped (vlm_model.py:729-731)."""
    if _LM_KEY in state and isinstance(state[_LM_KEY], dict):
        state = state[_LM_KEY]
    flat: dict = {}

    def walk(prefix, node):
        for k, v in node.items():
            name = f"{prefix}.{k}" if prefix else k
            if isinstance(v, dict):
                walk(name, v)
            elif not name.endswith("_extra_state"):
                flat[name] = v
    walk("", state)
    # older checkpoints: '.attention.' -> '.self_attention.' (vlm_model.py:752-760) and `layernorm` -> `norm` in every
    # decoder component name (ParallelTransformer.load_state_dict, megatron_patch/model/llava/transformer.py:1862-1873)
    return {k.replace("layernorm", "norm").replace(".attention.", ".self_attention.")
            if k.startswith("encoder.") else k: v for k, v in flat.items()}


def configure_provider(cfg: Optional[LeopardConfig] = None, hf_getter: Optional[Callable] = None,
                       megatron_state: Optional[dict] = None, device="cuda"):
    """What Megatron's `args` + checkpoint loading supply in the reference: the configuration and the weights
    `model_provider()` should build from.  `megatron_state` is `state['model']['language_model']` flattened to
    'encoder.layers.0....' keys (TP = 1)."""
    cfg = cfg or leopard_llava()
    if megatron_state is not None:
        hf = megatron_to_hf(cfg, megatron_state)
        hf_getter = hf.__getitem__
    _provider_state.update(cfg=cfg, getter=hf_getter, device=device)


class GPTModel(torch.nn.Module):
    """Native stand-in for megatron_patch.model.llava.gpt_model.GPTModel (a MegatronModule, gpt_model.py:54): forward (logits /
    per-token loss) here; the training iteration runs through `train_step` + leopard_b200.train.SFTTrainer on this model's engine
    (no autograd graph).  MegatronModule surface kept: `config`, `pre_process` / `post_process`,
    `share_embeddings_and_output_weights` (False: Leopard unties the lm_head), `set_input_tensor`,
    `state_dict_for_save_checkpoint`, `load_state_dict`, and — being an nn.Module — `parameters()` over the kernel-layout
    weights, `.cuda()`, `.eval()` / `.train()`, so Megatron's wrappers can hold it."""

    def __init__(self, cfg: LeopardConfig, native: NativeWeights, pre_process=True, post_process=True,
                 hf_getter: Optional[Callable] = None):
        super().__init__()
        if not (pre_process and post_process):
            raise NotImplementedError("pipeline parallel stages are out of scope (the reference runs PP = 1)")
        self.cfg = self.config = cfg
        self.pre_process, self.post_process = pre_process, post_process
        self.share_embeddings_and_output_weights = False
        self.engine = PrefillEngine(cfg, native)
        from .hf import _NativeParameters
        self.native = _NativeParameters(native.named)
        self.input_tensor = None
        self._hf_getter = hf_getter   # source tensors (HF layout) for state_dict_for_save_checkpoint

    # ---- checkpoint surface (gpt_model.py:115-137, vlm_model.py:684-791); TP = 1
    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        """{'language_model': {'embedding', 'vision_tower', 'mm_projector', 'encoder', 'output_layer'}} in Megatron's own
        layouts (group-interleaved QKV, [gate; up], [head, (q,k,v), 72]) — what Megatron's checkpointing writes under
        state['model'] and what hf2megatron_llava.py reads back."""
        if self._hf_getter is None:
            raise _lib.NativeError("this model was built from native weights only; no source tensors to export")
        # tensors the SFT step updates in place are read back from the device (kernel layout -> HF names); the frozen
        # rest (vision tower) comes from the source checkpoint
        live = named_to_hf(self.cfg, self.engine.w.named)
        src = self._hf_getter

        def get(name):
            t = live.get(name)
            return src(name) if t is None else t.detach().to("cpu", torch.bfloat16)

        return {_LM_KEY: nest_language_model_state(hf_to_megatron(self.cfg, get))}

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Customized load like the reference's (accepts the nested checkpoint entry with or without the
        'language_model' level): converts the Megatron layouts back and rebuilds the native weights on the device."""
        hf = megatron_to_hf(self.cfg, flatten_language_model_state(state_dict))
        device = self.engine.device
        self._hf_getter = hf.__getitem__
        # IN PLACE: an SFTTrainer built before the load (Megatron's order is setup_model_and_optimizer, then
        # load_checkpoint) keeps pointing at this engine's tensors, the ctypes weight structs and any captured CUDA graph
        # stay valid; the trainer's fp32 main copies are re-seeded from the loaded weights
        fresh = load_hf(self.cfg, self._hf_getter, device)
        old = self.engine.w
        if len(fresh._keep) != len(old._keep):
            raise _lib.NativeError("load_state_dict: checkpoint does not match this model's tensor inventory")
        for dst, src in zip(old._keep, fresh._keep):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise _lib.NativeError(f"load_state_dict: tensor mismatch {tuple(dst.shape)} vs {tuple(src.shape)}")
            dst.copy_(src)
        del fresh
        for ref in getattr(self.engine, "_trainers", []):
            tr = ref()
            if tr is not None:
                tr.resync_master()

    def set_input_tensor(self, t):  # gpt_model.py:86-88
        self.input_tensor = t

    @torch.no_grad()
    def forward(self, input_ids, position_ids=None, attention_mask=None, retriever_input_ids=None,
                retriever_position_ids=None, retriever_attn_mask=None, labels=None, tokentype_ids=None,
                inference_params=None, images=None):
        eng = self.engine
        b, s = input_ids.shape
        has_img = images is not None and images.shape[-1] > 100  # vlm_model.py:518
        feats = None
        if has_img:
            feats = eng.image_features(images.to(eng.device)).reshape(-1, self.cfg.dec.hidden)
        n_slots = int((input_ids == IMAGE_TOKEN_INDEX).sum())
        n_feat = 0 if feats is None else feats.shape[0]
        if has_img:
            assert n_feat == n_slots, \
                f"The number of image features doesn't match the number of image tokens. {n_feat} != {n_slots}"
        ids = input_ids.to(eng.device)
        x = torch.empty((b * s, self.cfg.dec.hidden), dtype=torch.bfloat16, device=eng.device)
        f0 = 0
        for i in range(b):  # features are consumed in row-major [b, s] order (vlm_model.py:531-533)
            k = int((ids[i] == IMAGE_TOKEN_INDEX).sum()) if has_img else 0
            row_ids = ids[i] if has_img else ids[i].clamp(min=0)
            eng.merge(row_ids, feats[f0:f0 + k] if k else None, IMAGE_TOKEN_INDEX, 1, out=x[i * s:(i + 1) * s])
            f0 += k
        if position_ids is None:
            pos = torch.arange(s, device=eng.device, dtype=torch.int32).repeat(b)
        else:
            pos = position_ids.reshape(-1)
        logits, _ = eng.decoder(x, pos, [i * s for i in range(b + 1)], "all", torch.bfloat16)
        if labels is None:
            return logits.view(b, s, -1)
        # post_language_model_processing (gpt_model.py:27-51): per-token CE of the fp32 logits against `labels`
        # (already shifted by get_batch, pretrain_megatron_llava.py:86-88); masking / averaging stays in loss_func
        from . import ops
        return ops.cross_entropy(logits, labels.reshape(-1).clamp(min=-1)).view(b, s)


def loss_func(loss_mask: torch.Tensor, labels: torch.Tensor, output_tensor: torch.Tensor):
    """Same signature and arithmetic as the script's `loss_func` (examples/llava/pretrain_megatron_llava.py:101-121, bound
    with `partial(loss_func, loss_mask, labels)` in forward_step :160): logits of position t are scored against the
    label at t + 1 (the shift happens HERE, not in get_batch), per-token cross entropy of the fp32 logits
    (`lp_cross_entropy` replaces vocab_parallel_cross_entropy, TP = 1), weighted by `loss_mask[:, 1:]`, divided by the sum
    of the weights.  Returns (loss, {'lm loss': loss}) — the data-parallel average of the logging copy is the caller's
    (Megatron's) business.  Labels equal to the image slot id (-200) or to ignore ids never carry weight in the
    reference's batches; negative labels score 0 here instead of indexing out of range."""
    from . import ops
    b, s, v = output_tensor.shape
    logits = output_tensor[:, :-1, :].reshape(b * (s - 1), v)
    if logits.dtype not in (torch.float32, torch.bfloat16):
        logits = logits.float()
    tgt = labels[:, 1:].reshape(-1).to(logits.device).clamp(min=-1)
    w = loss_mask[:, 1:].reshape(-1).to(logits.device).float()
    per_token = ops.cross_entropy(logits.contiguous(), tgt.contiguous())
    loss = torch.sum(per_token.view(-1) * w) / w.sum()
    return loss, {"lm loss": loss}


IGNORE_INDEX = -100


def mm_collate_fn(batch):
    """The script's collate function (pretrain_megatron_llava.py:163-191), same output dict: drops None samples (None for
    an empty batch), right-pads input_ids / labels with IGNORE_INDEX and weights / preferences with 0 up to the longest
    sample rounded UP to a multiple of 8, concatenates the per-sample crop tensors along dim 0, and substitutes the
    `[1, 10]` bf16 dummy (the `images.shape[-1] > 100` sentinel of vlm_model.py:518) when no sample has an image."""
    batch = [item for item in batch if item is not None]
    if len(batch) == 0:
        return None
    max_len = -(-max(len(item["input_ids"]) for item in batch) // 8) * 8
    n = len(batch)
    out = {"input_ids": torch.full((n, max_len), IGNORE_INDEX, dtype=torch.long),
           "labels": torch.full((n, max_len), IGNORE_INDEX, dtype=torch.long),
           "weights": torch.zeros(n, max_len, dtype=torch.float),
           "preferences": torch.zeros(n, max_len, dtype=torch.float)}
    for i, item in enumerate(batch):
        for key in ("input_ids", "labels", "weights", "preferences"):
            out[key][i, : len(item[key])] = item[key]
    images = [item["image"] for item in batch if "image" in item and item["image"] is not None]
    out["image"] = torch.cat(images, dim=0) if images else torch.zeros(1, 10, dtype=torch.bfloat16)
    return out


class _BatchRecorder:
    """What `forward_step_func` sees as the model: records the arguments of `model(tokens, position_ids, attention_mask,
    images=images)` (pretrain_megatron_llava.py:157-158) instead of computing — the fused native step does the arithmetic."""

    def __init__(self, real):
        self._real, self.call = real, None

    def __call__(self, tokens, position_ids=None, attention_mask=None, **kw):
        self.call = dict(tokens=tokens, position_ids=position_ids, attention_mask=attention_mask, **kw)
        return None

    def __getattr__(self, name):
        return getattr(self._real, name)


def _batch_from_forward_step(forward_step_func, data_iterator, model):
    """Run the script's own `forward_step(data_iterator, model)` (get_batch incl. its broadcast / masks, :124-160) against a
    recorder and read the micro-batch back: tokens and images from the model call, loss mask and labels from the returned
    `partial(loss_func, loss_mask, labels)` (:160)."""
    rec = _BatchRecorder(model[0] if isinstance(model, (list, tuple)) else model)
    out = forward_step_func(data_iterator, rec)
    loss_closure = out[1] if isinstance(out, tuple) else None
    if rec.call is None or loss_closure is None or not hasattr(loss_closure, "args") or len(loss_closure.args) < 2:
        raise _lib.NativeError("train_step: forward_step_func must call model(tokens, position_ids, attention_mask, images=...) and "
                               "return (output, partial(loss_func, loss_mask, labels)) as examples/llava/pretrain_megatron_llava.py does")
    loss_mask, labels = loss_closure.args[0], loss_closure.args[1]
    return rec.call["tokens"].long(), labels.long(), loss_mask.float(), rec.call.get("images")


def train_step(forward_step_func, data_iterator, model, optimizer, opt_param_scheduler=None, config=None,
               num_microbatches: int = 1):
    """Megatron's `train_step` (Megatron-LM-240424/megatron/training/training.py:544-607) for this path: same signature,
    same return tuple (loss_reduced, skipped_iter, grad_norm, num_zeros_in_grad).  `optimizer` is a
    leopard_b200.train.SFTTrainer built on `model[0].engine`: zero-grad, forward, loss_func, backward, the data-parallel
    reduce, clipping and the Adam update all happen inside its `step`, on native kernels.  `forward_step_func` — the script's
    own forward_step — IS called, once per micro-batch, against a recorder in place of the model: its get_batch supplies the
    batch (tokens / images from the model call, loss mask and labels from the returned loss closure); the arithmetic it would
    have triggered through autograd runs in the fused step instead.  With forward_step_func=None the batch is read directly
    with get_batch's keys (:52-99): 'input_ids', 'labels', 'weights' (the loss mask), 'image'.
    The learning-rate schedule lives in the trainer (cosine with warm-up, as the launch script configures);
    `num_microbatches` = Megatron's get_num_microbatches() (global batch / (micro batch x data-parallel size))."""
    micro = []
    for _ in range(num_microbatches):        # get_num_microbatches() batches per iteration, gradients accumulated
        if forward_step_func is not None:
            micro.append(_batch_from_forward_step(forward_step_func, data_iterator, model))
            continue
        data = next(data_iterator)
        while data is None:                  # get_batch :59-62 skips empty collates
            data = next(data_iterator)
        micro.append((data["input_ids"].long(), data["labels"].long(), data["weights"].float(), data["image"]))
    if num_microbatches == 1:
        loss, grad_norm = optimizer.step(*micro[0])
    else:
        loss, grad_norm = optimizer.step(*[list(col) for col in zip(*micro)])
    return {"lm loss": loss}, int(getattr(optimizer, "skipped_last", False)), grad_norm, None


def model_provider(pre_process=True, post_process=True):
    """Same signature as examples/llava/pretrain_megatron_llava.py:37-50."""
    st = _provider_state
    if st["getter"] is None:
        raise _lib.NativeError("model_provider(): call configure_provider(...) first (Megatron's args/checkpoint "
                               "plumbing is outside the hot path)")
    native = load_hf(st["cfg"], st["getter"], st["device"])
    return GPTModel(st["cfg"], native, pre_process, post_process, hf_getter=st["getter"])


# ----------------------------------------------------------------------------------------------------------------
# Idefics2 provider (Pai-Megatron-Patch/examples/idefics2/pretrain_megatron_idefics2.py:36-54)
# ----------------------------------------------------------------------------------------------------------------
class Idefics2GPTModel:
    """Forward-only stand-in for megatron_patch.model.idefics2.gpt_model.GPTModel (:95-122): same positional /
    keyword signature, `pixel_values [b, n, 3, H, W]` + `pixel_attention_mask [b, n, H, W]`, logits `[b, s, V]` or the
    per-token loss `[b, s]` when `labels` is given.  b = 1 (the shipped script runs micro-batch 1)."""

    def __init__(self, cfg, getter, device):
        from .idefics2 import Idefics2Engine, Idefics2Weights
        self.cfg = cfg
        self.engine = Idefics2Engine(cfg, Idefics2Weights(cfg, getter, device))

    def set_input_tensor(self, t):
        self.input_tensor = t

    @torch.no_grad()
    def forward(self, input_ids, position_ids=None, attention_mask=None, retriever_input_ids=None,
                retriever_position_ids=None, retriever_attn_mask=None, labels=None, images=None,
                inference_params=None, pixel_values=None, pixel_attention_mask=None):
        logits, _, _ = self.engine.prefill(input_ids, pixel_values, pixel_attention_mask, "all")
        logits = logits[None]
        if labels is None:
            return logits
        from . import ops
        b, s, _ = logits.shape
        return ops.cross_entropy(logits[0].contiguous(), labels.reshape(-1).clamp(min=-1)).view(b, s)

    __call__ = forward


def idefics2_model_provider(cfg, hf_getter, device="cuda", pre_process=True, post_process=True):
    """`model_provider` twin for Idefics2: `cfg` = leopard_b200.config.Idefics2Config, `hf_getter(name)` returns HF
    Idefics2-layout tensors (the layout `AutoConfig/…from_pretrained(args.load)` refers to in the reference)."""
    if not (pre_process and post_process):
        raise NotImplementedError("pipeline parallel stages are out of scope (the reference runs PP = 1)")
    return Idefics2GPTModel(cfg, hf_getter, device)
