"""CPU ORACLE (test infrastructure — never imported by the product path): the Megatron SFT step of Leopard-LLaVA.

What one training iteration of Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py does to the model, restated
in plain torch fp32 on the CPU:
  forward_step :124-160      logits = model(tokens, position_ids, attention_mask, images=images)
  loss_func    :101-121      shift, per-token CE on fp32 logits, weights = loss_mask[:, 1:], divide by the weight sum
  backward                   torch autograd in the reference; here (a) autograd over oracle.model — the end-to-end
                             checker — and (b) the closed-form backward of every block, which is the specification the
                             CUDA kernels are written against; tests/test_train_oracle.py pins (b) to (a)
  optimizer.step             Megatron-LM-240424/megatron/core/optimizer/optimizer.py:300-370 (Float16Optimizer: fp32 main
                             params and grads), clip_grads.py:103-141 (global L2 norm, coeff = max / (norm + 1e-6)),
                             apex FusedAdam in adam_w_mode (un-vendored; restated from the published AdamW update and
                             pinned against torch.optim.AdamW in the tests)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Tuple

import torch
import torch.nn.functional as F

from . import model as M


# ----------------------------------------------------------------------------------------------------------------
# closed-form backward of the blocks (specification of the kernels in leopard_b200/csrc/lp_train.cu)
# ----------------------------------------------------------------------------------------------------------------
def linear_bwd(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, has_bias: bool = False):
    """y = x w^T (+ b):  dx = dy w (dgrad), dw = dy^T x (wgrad), db = column sums of dy."""
    return dy @ w, dy.t() @ x, (dy.sum(0) if has_bias else None)


def rmsnorm_bwd(x: torch.Tensor, w: torch.Tensor, eps: float, dy: torch.Tensor):
    """y = x r w with r = rsqrt(mean(x^2) + eps):  dx = r (dy w) - x r^3 mean(dy w x);  dw = sum_rows dy x r."""
    r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    g = dy * w
    dx = r * g - x * r.pow(3) * (g * x).mean(-1, keepdim=True)
    return dx, (dy * x * r).sum(0)


def layernorm_bwd(x: torch.Tensor, w: torch.Tensor, eps: float, dy: torch.Tensor):
    """y = (x - mu) r w + b:  dx = r (g - mean(g) - xh mean(g xh)), g = dy w, xh = (x - mu) r."""
    mu = x.mean(-1, keepdim=True)
    r = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
    xh = (x - mu) * r
    g = dy * w
    dx = r * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    return dx, (dy * xh).sum(0), dy.sum(0)


def swiglu_bwd(g: torch.Tensor, u: torch.Tensor, dm: torch.Tensor):
    """m = silu(g) u:  dg = dm u s (1 + g (1 - s)), du = dm g s, s = sigmoid(g)."""
    s = torch.sigmoid(g)
    return dm * u * s * (1 + g * (1 - s)), dm * g * s


def gelu_bwd(x: torch.Tensor, dy: torch.Tensor, approximate: str = "none"):
    """exact (erf) GELU: dy (Phi(x) + x phi(x));  tanh GELU: dy (0.5 (1 + t) + 0.5 x (1 - t^2) c (1 + 3 a x^2))."""
    if approximate == "tanh":
        c, a = math.sqrt(2.0 / math.pi), 0.044715
        t = torch.tanh(c * (x + a * x ** 3))
        return dy * (0.5 * (1 + t) + 0.5 * x * (1 - t * t) * c * (1 + 3 * a * x * x))
    phi = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    return dy * (0.5 * (1 + torch.erf(x / math.sqrt(2.0))) + x * phi)


def rope_bwd(dt: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """RoPE is a rotation per (position, frequency) pair, so its adjoint is the rotation by the opposite angle."""
    return M.apply_rope(dt, cos, -sin)


def attention_fwd(q, k, v, causal: bool, scale: float):
    """q [S, H, D], k/v [S, Hkv, D] (q head j reads kv head j // (H / Hkv)).  Returns (o [S, H, D], lse [H, S]) with
    lse the natural-log sum-exp of the scaled, masked scores."""
    S, H, D = q.shape
    rep = H // k.shape[1]
    kk = k.repeat_interleave(rep, 1).transpose(0, 1)
    vv = v.repeat_interleave(rep, 1).transpose(0, 1)
    sc = (q.transpose(0, 1) @ kk.transpose(-1, -2)) * scale
    if causal:
        sc = sc.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
    lse = torch.logsumexp(sc, -1)
    o = torch.exp(sc - lse[..., None]) @ vv
    return o.transpose(0, 1), lse


def attention_bwd(q, k, v, o, lse, do, causal: bool, scale: float):
    """P = exp(scale q k^T - lse);  delta = rowsum(do o);  dV = P^T dO;  dS = P (dO V^T - delta);
    dQ = scale dS K;  dK = scale dS^T Q;  dK / dV of a kv head sum over the q heads of its group."""
    S, H, D = q.shape
    Hkv = k.shape[1]
    rep = H // Hkv
    qq, oo, dd = q.transpose(0, 1), o.transpose(0, 1), do.transpose(0, 1)          # [H, S, D]
    kk = k.repeat_interleave(rep, 1).transpose(0, 1)
    vv = v.repeat_interleave(rep, 1).transpose(0, 1)
    sc = (qq @ kk.transpose(-1, -2)) * scale
    if causal:
        sc = sc.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
    p = torch.exp(sc - lse[..., None])
    delta = (dd * oo).sum(-1, keepdim=True)
    dv = p.transpose(-1, -2) @ dd
    ds = p * (dd @ vv.transpose(-1, -2) - delta)
    dq = (ds @ kk) * scale
    dk = (ds.transpose(-1, -2) @ qq) * scale
    dk = dk.view(Hkv, rep, S, D).sum(1)
    dv = dv.view(Hkv, rep, S, D).sum(1)
    return dq.transpose(0, 1), dk.transpose(0, 1), dv.transpose(0, 1)


def masked_lm_loss(logits: torch.Tensor, labels: torch.Tensor, loss_mask: torch.Tensor):
    """loss_func (pretrain_megatron_llava.py:101-121) and its gradient w.r.t. the logits.
    logits [b, s, V], labels [b, s] (any id at unweighted positions), loss_mask [b, s] float.
    loss = sum_t w[t+1] CE(logits[t], labels[t+1]) / sum_t w[t+1];  dlogits[t] = (softmax(logits[t]) - onehot) w[t+1] / sum w,
    the last position of every sequence gets no gradient."""
    b, s, v = logits.shape
    x = logits[:, :-1].float()
    tgt = labels[:, 1:]
    w = loss_mask[:, 1:].float()
    valid = (tgt >= 0) & (tgt < v)
    lse = torch.logsumexp(x, -1)
    picked = x.gather(-1, tgt.clamp(0, v - 1).unsqueeze(-1)).squeeze(-1)
    per_tok = torch.where(valid, lse - picked, lse)     # Megatron's vocab-parallel CE scores an out-of-range label as lse
    denom = w.sum()
    loss = (per_tok * w).sum() / denom
    p = torch.softmax(x, -1)
    onehot = torch.zeros_like(p).scatter_(-1, tgt.clamp(0, v - 1).unsqueeze(-1), valid.float().unsqueeze(-1))
    d = torch.zeros_like(logits, dtype=torch.float32)
    d[:, :-1] = (p - onehot) * (w / denom).unsqueeze(-1)
    return loss, d


# ----------------------------------------------------------------------------------------------------------------
# optimizer (fp32 main params; Megatron Float16OptimizerWithFloat16Params + apex FusedAdam, adam_w_mode)
# ----------------------------------------------------------------------------------------------------------------
def global_grad_norm(grads: Iterable[torch.Tensor]) -> float:
    """clip_grads.py:103-130, norm_type 2: sqrt of the sum of squared entries over every gradient."""
    return float(math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads)))


def clip_coefficient(total_norm: float, max_norm: float) -> float:
    """clip_grads.py:133-139: grads are multiplied by max / (norm + 1e-6) only when that is below 1."""
    c = max_norm / (total_norm + 1.0e-6)
    return c if c < 1.0 else 1.0


def adamw_update(p, g, m, v, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                 weight_decay: float = 0.0):
    """One FusedAdam (adam_w_mode) update in fp32, in place on p, m, v; `step` counts from 1."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    mh = m / (1 - beta1 ** step)
    vh = v / (1 - beta2 ** step)
    p.sub_(lr * (mh / (vh.sqrt() + eps) + weight_decay * p))
    return p


def cosine_lr(it: int, lr: float, min_lr: float, warmup_iters: int, decay_iters: int) -> float:
    """Megatron OptimizerParamScheduler.get_lr, 'cosine' style
    (Megatron-LM-240424/megatron/training/optimizer_param_scheduler.py:79-126, init_lr = 0): linear warm-up from 0, then
    min_lr + (lr - min_lr) * 0.5 (1 + cos(pi * progress)), min_lr after decay_iters."""
    if warmup_iters > 0 and it <= warmup_iters:
        return lr * it / warmup_iters
    if it > decay_iters:
        return min_lr
    ratio = (it - warmup_iters) / (decay_iters - warmup_iters)
    return min_lr + (lr - min_lr) * 0.5 * (math.cos(math.pi * ratio) + 1.0)


# ----------------------------------------------------------------------------------------------------------------
# end-to-end SFT step by autograd over the oracle model (the checker for the product's hand-written backward)
# ----------------------------------------------------------------------------------------------------------------
class TrainableCheckpoint:
    """Wraps a checkpoint (`f32(name)`) so that every tensor is a persistent fp32 autograd leaf."""

    def __init__(self, ck, trainable=lambda name: True):
        self._ck = ck
        self._trainable = trainable
        self.params: Dict[str, torch.Tensor] = {}

    def f32(self, name: str) -> torch.Tensor:
        t = self.params.get(name)
        if t is None:
            t = self._ck.f32(name).detach().clone().float()
            t.requires_grad_(bool(self._trainable(name)))
            self.params[name] = t
        return t

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: p.grad for k, p in self.params.items() if p.requires_grad and p.grad is not None}


def forward_logits(ck, cfg, tokens: torch.Tensor, images: Optional[torch.Tensor]) -> torch.Tensor:
    """TransformerLanguageModel.forward for a [b, s] batch in the Megatron convention (one -200 per feature row,
    vlm_model.py:505-533): ViT -> projector -> scatter -> decoder -> logits [b, s, V].  Position ids are 0..s-1 per
    sequence (get_ltor_masks_and_position_ids without resets)."""
    feats = None
    if images is not None and images.shape[-1] > 100:
        feats = M.projector(ck, cfg, M.vit_forward(ck, cfg, images))
    if feats is None:
        emb = ck.f32(M.LM + "embed_tokens.weight")[tokens.clamp(min=0)]
    else:
        emb = M.merge_megatron_scatter(ck, cfg, tokens, feats)
    pos = torch.arange(tokens.shape[1])
    return torch.stack([M.decoder_forward(ck, cfg, emb[i], pos) for i in range(tokens.shape[0])], 0)


def sft_loss_and_grads(ck: TrainableCheckpoint, cfg, tokens, labels, loss_mask, images) -> Tuple[float, Dict[str, torch.Tensor]]:
    for p in ck.params.values():
        p.grad = None
    logits = forward_logits(ck, cfg, tokens, images)
    loss, _ = masked_lm_loss(logits, labels, loss_mask)
    loss.backward()
    return float(loss.detach()), ck.grads()


def sft_step(ck: TrainableCheckpoint, cfg, state: dict, tokens, labels, loss_mask, images, lr: float,
             clip_grad: float = 1.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
             weight_decay: float = 0.0):
    """One full iteration: loss, grads, global-norm clip, AdamW on every trainable tensor.  `state` carries
    {'step': int, 'm': {...}, 'v': {...}} between calls.  Returns (loss, grad_norm)."""
    loss, grads = sft_loss_and_grads(ck, cfg, tokens, labels, loss_mask, images)
    norm = global_grad_norm(grads.values())
    coeff = clip_coefficient(norm, clip_grad) if clip_grad > 0 else 1.0
    state["step"] = state.get("step", 0) + 1
    with torch.no_grad():
        for name, g in grads.items():
            p = ck.params[name]
            m = state.setdefault("m", {}).setdefault(name, torch.zeros_like(p))
            v = state.setdefault("v", {}).setdefault(name, torch.zeros_like(p))
            adamw_update(p, g * coeff, m, v, state["step"], lr, beta1, beta2, eps, weight_decay)
    return loss, norm
