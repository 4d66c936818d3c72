"""SFT step on the native kernels — FIRST IMPLEMENTATION, not yet measured on a B200.

What Megatron's `train_step` does for Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py, sequenced on the host
over C-ABI calls (allocation, shapes and ordering only; every contraction, reduction and update is a native kernel):

  forward_step :124-160   ViT (frozen: --freeze-clip-vision-tower) -> pixel shuffle -> projector -> scatter into the token
                          embeddings -> 32 decoder layers -> final norm -> fp32 logits            (same kernels as prefill)
  loss_func    :101-121   lp_ce_fwd_bwd on rows shifted by one (weights = loss_mask / sum(loss_mask) / dp_world)
  backward                layer by layer with full recompute from the saved layer inputs (--recompute-granularity full
                          --recompute-method uniform --recompute-num-layers 1, run_pretrain_megatron_llava.sh:96-100;
                          recompute="none" keeps the activations instead — they fit a B200 — and skips that quarter of the work):
                          dgrad = lp_gemm_bf16(dY, W^T), wgrad = lp_gemm_bf16(dY^T, X^T) -> fp32 main grads
                          (lp_transpose_bf16 supplies the re-laid-out operands), norm / SwiGLU / GELU / RoPE / merge
                          backward in lp_train.cu; attention backward over materialised [T, T] score tiles per head
                          (lp_softmax_bwd) — a flash-style tcgen05 backward replaces it next
  reduce                  data-parallel all-reduce of the flat fp32 gradient buffer, one bucket per decoder layer issued
                          as soon as that layer's backward is enqueued (NCCL on its own stream overlaps the next layer)
  optimizer.step          clip_grads.py:103-141 (global L2 norm) + FusedAdam adam_w_mode on fp32 main params
                          (optimizer.py:300-370), bf16 model weights refreshed in the same kernel

Specification and checker: oracle/train.py (closed forms pinned to autograd; end-to-end step by autograd over the prefill
oracle).  tools/gpu_check_train.py compares this module with it on a B200.  There is no CPU fallback.
The vision tower is frozen by default (--freeze-clip-vision-tower); freeze_vision_tower=False back-propagates through it
with the materialised attention backward per (crop, head) — functional, launch-bound.  Not covered yet: TP > 1.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops

ALIGN = 64  # fp32 elements: every tensor's slice of the flat buffers starts on a 256-byte boundary


# ----------------------------------------------------------------------------------------------------------------
# host helpers (pure; tested on the CPU)
# ----------------------------------------------------------------------------------------------------------------
def plan_flat(sizes: Sequence[Tuple[str, int]]) -> Tuple[Dict[str, Tuple[int, int]], int]:
    """[(name, numel)] in bucket order -> ({name: (offset, numel)}, total) with ALIGN-element alignment."""
    off, out = 0, {}
    for name, n in sizes:
        out[name] = (off, n)
        off += -(-n // ALIGN) * ALIGN
    return out, off


def shifted_targets(labels: torch.Tensor, loss_mask: torch.Tensor, dp_world: int = 1):
    """loss_func's shift (pretrain_megatron_llava.py:103-105) as per-row targets and weights for lp_ce_fwd_bwd:
    row (b, t) scores labels[b, t + 1] with weight loss_mask[b, t + 1] / sum(loss_mask[:, 1:]) / dp_world; the last row of
    every sequence has weight 0.  Returns (targets int64 [b * s], weights fp32 [b * s])."""
    b, s = labels.shape
    tgt = torch.full((b, s), -1, dtype=torch.long, device=labels.device)
    tgt[:, :-1] = labels[:, 1:]
    w = torch.zeros((b, s), dtype=torch.float32, device=labels.device)
    w[:, :-1] = loss_mask[:, 1:].float()
    w = w / (w.sum() * dp_world)
    return tgt.reshape(-1).contiguous(), w.reshape(-1).contiguous()


def layer_buckets(names: Sequence[str], layout: Dict[str, Tuple[int, int]], total: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) ranges of the flat gradient buffer, one per name prefix ('dec.31', ..., 'proj', ...) in
    the order the backward pass completes them (the order of `names`)."""
    out: List[Tuple[str, int, int]] = []
    for n in names:
        key = ".".join(n.split(".")[:2]) if n.startswith("dec.") and n.split(".")[1].isdigit() else n.split(".")[0]
        off, cnt = layout[n]
        end = off + -(-cnt // ALIGN) * ALIGN
        if out and out[-1][0] == key and out[-1][2] == off:
            out[-1] = (key, out[-1][1], end)
        else:
            out.append((key, off, end))
    assert out[-1][2] == total
    return [(a, b) for _, a, b in out]


def cosine_lr(it: int, lr: float, min_lr: float, warmup_iters: int, decay_iters: int) -> float:
    """OptimizerParamScheduler.get_lr, 'cosine' (Megatron-LM-240424/megatron/training/optimizer_param_scheduler.py:79-126)."""
    if warmup_iters > 0 and it <= warmup_iters:
        return lr * it / warmup_iters
    if it > decay_iters or decay_iters <= warmup_iters:
        return min_lr
    ratio = (it - warmup_iters) / (decay_iters - warmup_iters)
    return min_lr + (lr - min_lr) * 0.5 * (math.cos(math.pi * ratio) + 1.0)


def trainable_names(n_layers: int, freeze_llm: bool = False, with_projector: bool = True, vit_layers: int = 0) -> List[str]:
    """Names (NativeWeights.named) in the order their gradients are completed by the backward pass; vit_layers > 0 adds the
    vision tower (one bucket: it is 5 % of the parameters)."""
    names: List[str] = []
    if not freeze_llm:
        names += ["dec.lm_head", "dec.final_g"]
        for i in reversed(range(n_layers)):
            names += [f"dec.{i}.down_w", f"dec.{i}.gateup_w", f"dec.{i}.ln2_g", f"dec.{i}.o_w", f"dec.{i}.qkv_w",
                      f"dec.{i}.ln1_g"]
        names += ["embed"]
    if with_projector:
        names += ["proj.w2", "proj.b2", "proj.w1", "proj.b1"]
    if vit_layers > 0:
        names += ["vit.post_g", "vit.post_b"]
        for i in reversed(range(vit_layers)):
            names += [f"vit.{i}.{k}" for k in ("fc2_w", "fc2_b", "fc1_w", "fc1_b", "ln2_g", "ln2_b", "out_w", "out_b", "qkv_w",
                                               "qkv_b", "ln1_g", "ln1_b")]
        names += ["vit.pos", "vit.patch_b", "vit.patch_w"]
    return names


# ----------------------------------------------------------------------------------------------------------------
class SFTTrainer:
    """`tr = SFTTrainer(engine, lr=1e-5); loss, grad_norm = tr.step(tokens, labels, loss_mask, images)`

    tokens / labels / loss_mask: [b, s] in the Megatron convention of mm_collate_fn (one -200 per image feature row,
    labels not shifted, loss_mask = the dataset's 'weights'); every sequence of the micro-batch has the same length s,
    s % 8 == 0.  images: [N, 3, 364, 364] crops of the whole micro-batch in slot order."""

    def __init__(self, engine, lr: float = 1e-5, min_lr: float = 0.0, warmup_iters: int = 0, decay_iters: int = 0,
                 beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.0,
                 clip_grad: float = 1.0, freeze_llm: bool = False, freeze_vision_tower: bool = True, process_group=None,
                 attn_bwd: Optional[str] = None, recompute: str = "full"):
        if not torch.cuda.is_available():
            raise _lib.NativeError("leopard_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _lib.lib()
        self.eng, self.cfg, self.w = engine, engine.cfg, engine.w
        self.dev = engine.device
        self.hp = dict(lr=lr, min_lr=min_lr, warmup=warmup_iters, decay=decay_iters, beta1=beta1, beta2=beta2, eps=eps,
                       wd=weight_decay, clip=clip_grad)
        self.freeze_llm = freeze_llm
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.iteration = 0
        d = self.cfg.dec
        self.freeze_vit = freeze_vision_tower
        self.names = trainable_names(d.layers, freeze_llm, vit_layers=0 if freeze_vision_tower else self.cfg.vit.layers)
        named = self.w.named
        self.layout, total = plan_flat([(n, named[n].numel()) for n in self.names])
        self.buckets = layer_buckets(self.names, self.layout, total)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.flat_grad = torch.zeros(total, **f32)
        self.flat_m = torch.zeros(total, **f32)
        self.flat_v = torch.zeros(total, **f32)
        # fp32 main copies of the bf16 matrices; fp32 vectors (norm weights, biases) are their own main copy
        mat = [n for n in self.names if named[n].dtype == torch.bfloat16]
        self.master_layout, mtotal = plan_flat([(n, named[n].numel()) for n in mat])
        self.flat_master = torch.empty(mtotal, **f32)
        for n in mat:
            self._view(self.flat_master, self.master_layout, n).copy_(named[n].float())
        self._norm_ws = torch.empty(_lib.lib().lp_grad_norm_workspace_bytes() // 4, **f32)
        self._norm_out = torch.zeros(2, **f32)
        self.q_w = d.heads * d.head_dim
        self.kv_w = d.kv_heads * d.head_dim
        self.qkv_w = self.q_w + 2 * self.kv_w
        self._scratch_g = None
        # "full": keep only each layer's input and recompute the layer in the backward pass (the reference's setting,
        # --recompute-granularity full).  "none": keep every layer's activations — ~0.53 GB per decoder layer at S = 4096,
        # 17 GB for the 32 layers, which the 180 GB of a B200 has room for next to the 144 GB of weights and optimizer state
        # — and skip the recompute, a quarter of the step's GEMM work.
        if recompute not in ("full", "none"):
            raise ValueError(f"recompute must be 'full' or 'none', got {recompute!r}")
        self.recompute = recompute
        self._accum = False        # True while a micro-batch after the first is being back-propagated
        self.skipped_last = False  # the last step() found non-finite gradients and left the weights untouched
        self._acc_scratch = {}
        # attention backward: 'materialized' (verified composition of GEMMs + lp_softmax_bwd) or 'flash' (lp_attention_bwd)
        # the model shim re-seeds the fp32 main copies when a checkpoint is loaded into the engine after this trainer was built
        # (Megatron's order: setup_model_and_optimizer, then load_checkpoint)
        if not hasattr(engine, "_trainers"):
            engine._trainers = []
        engine._trainers.append(weakref.ref(self))
        # dgrad / wgrad operands: "mn" = the GEMM consumes the row-major tensors as MN-major operands (lp_gemm_bf16_ex a_t / b_t), no
        # transposed copies; "transpose" = explicit lp_transpose_bf16 copies (any shape).  Shapes the MN launch cannot take (fewer
        # 256 x 256 tiles than SM pairs: the vision tower's 1152 x 1152 matrices) use the copies either way.
        self.operands = os.environ.get("LEOPARD_SFT_OPERANDS", "mn")
        try:
            self._pairs = torch.cuda.get_device_properties(self.dev).multi_processor_count // 2
        except Exception:  # noqa: BLE001  (CPU sequencing tests run this file against the op-contract restatement)
            self._pairs = 74
        self.attn_bwd = attn_bwd or os.environ.get("LEOPARD_ATTN_BWD", "flash")   # flash: 73 vs 102 ms per 4-layer step on a B200
        if self.attn_bwd not in ("materialized", "flash"):
            raise ValueError(f"attn_bwd must be 'materialized' or 'flash', got {self.attn_bwd!r}")

    # ------------------------------------------------------------------ views
    def _view(self, flat: torch.Tensor, layout, name: str) -> torch.Tensor:
        off, n = layout[name]
        return flat[off:off + n].view(self.w.named[name].shape)

    def grad(self, name: str) -> torch.Tensor:
        return self._view(self.flat_grad, self.layout, name)

    def resync_master(self):
        """Re-seed the fp32 main copies from the engine's (re-loaded) bf16 weights; Adam moments are kept."""
        for n in self.master_layout:
            self._view(self.flat_master, self.master_layout, n).copy_(self.w.named[n].float())

    # ------------------------------------------------------------------ optimizer state (checkpoint / resume)
    def state_dict(self) -> dict:
        """What Megatron keeps under state['optimizer'] for this path: the step count, Adam moments and the fp32 main
        params, per tensor name (kernel layout; model weights themselves are exported by the Megatron shim's
        state_dict_for_save_checkpoint).  fp32 vectors are their own main copy and are saved as such."""
        named = self.w.named
        out = {"iteration": self.iteration, "hyper": dict(self.hp), "exp_avg": {}, "exp_avg_sq": {}, "main": {}}
        for n in self.names:
            out["exp_avg"][n] = self._view(self.flat_m, self.layout, n).detach().cpu().clone()
            out["exp_avg_sq"][n] = self._view(self.flat_v, self.layout, n).detach().cpu().clone()
            main = self._view(self.flat_master, self.master_layout, n) if n in self.master_layout else named[n]
            out["main"][n] = main.detach().cpu().clone()
        return out

    def load_state_dict(self, sd: dict):
        """Resume: restores step count, moments and main params, and refreshes the bf16 model weights from the latter."""
        named = self.w.named
        missing = [n for n in self.names if n not in sd["main"]]
        if missing:
            raise KeyError(f"optimizer state lacks {missing[:3]}{'...' if len(missing) > 3 else ''}")
        self.iteration = int(sd["iteration"])
        for n in self.names:
            self._view(self.flat_m, self.layout, n).copy_(sd["exp_avg"][n])
            self._view(self.flat_v, self.layout, n).copy_(sd["exp_avg_sq"][n])
            if n in self.master_layout:
                self._view(self.flat_master, self.master_layout, n).copy_(sd["main"][n])
                named[n].copy_(sd["main"][n])          # fp32 -> bf16, round to nearest even like the update kernel
            else:
                named[n].copy_(sd["main"][n])

    # ------------------------------------------------------------------ GEMM forms
    def _dgrad(self, dy: torch.Tensor, w: torch.Tensor, out=None) -> torch.Tensor:
        """dX[M, K] = dY[M, N] W[N, K]: W is the contraction-major ("transposed") operand — consumed MN-major in place when the
        shape allows, else through an explicit W^T [K, N] copy."""
        if self._mn_ok(dy.shape[0], w.shape[1]):
            return ops.gemm_ex(dy, w, out=out, b_t=True)
        return ops.gemm(dy, ops.transpose(w), out=out)

    def _mn_ok(self, m: int, n: int) -> bool:
        """Can lp_gemm_bf16_ex take this [m, n] output with transposed operands (CTA-pair mode, 256-wide tiles)?"""
        return (self.operands == "mn" and n % 8 == 0 and m % 8 == 0 and m > 128 and -(-m // 256) * -(-n // 256) >= self._pairs)

    def _sink(self, name: str) -> Optional[torch.Tensor]:
        """Where a kernel that OVERWRITES its output should write the gradient of `name`: the tensor's slice of the flat
        buffer on the first micro-batch, a scratch tensor of the same shape afterwards (added in by `_commit`), None when
        the tensor is frozen."""
        if name not in self.layout:
            return None
        if not self._accum:
            return self.grad(name)
        shape = self.w.named[name].shape
        buf = self._acc_scratch.get(shape)
        if buf is None:
            buf = torch.empty(shape, dtype=torch.float32, device=self.dev)
            self._acc_scratch[shape] = buf
        return buf

    def _commit(self, name: str, sink: Optional[torch.Tensor]):
        if sink is not None and self._accum:
            ops.accumulate(self.grad(name).reshape(-1), sink.reshape(-1))

    def _wgrad(self, name: str, dy: torch.Tensor, x: torch.Tensor):
        """grad[name][N, K] (fp32) (+)= dY^T[N, M] X[M, K], fed dY^T and X^T [K, M]; skipped when frozen."""
        sink = self._sink(name)
        if sink is not None:
            if self._mn_ok(dy.shape[1], x.shape[1]):      # dY [tokens, N] = At, X [tokens, K] = Wt: both consumed in place
                ops.gemm_ex(dy, x, out=sink, a_t=True, b_t=True)
            else:
                ops.gemm(ops.transpose(dy), ops.transpose(x), out=sink)
            self._commit(name, sink)

    def _rms_bwd(self, name: str, x, gamma, dy, add=None) -> torch.Tensor:
        """RMSNorm backward; the weight gradient goes to `name` (scratch when that weight is frozen)."""
        sink = self._sink(name)
        if sink is None:
            if self._scratch_g is None:
                self._scratch_g = torch.empty(self.cfg.dec.hidden, dtype=torch.float32, device=self.dev)
            return ops.rmsnorm_bwd(x, gamma, dy, self.cfg.dec.eps, self._scratch_g, add=add)
        dx = ops.rmsnorm_bwd(x, gamma, dy, self.cfg.dec.eps, sink, add=add)
        self._commit(name, sink)
        return dx

    def _colsum(self, name: str, dy: torch.Tensor):
        sink = self._sink(name)
        if sink is not None:
            ops.colsum(dy, sink)
            self._commit(name, sink)

    # ------------------------------------------------------------------ decoder layer
    def _layer_weights(self, i: int):
        n = self.w.named
        return (n[f"dec.{i}.ln1_g"], n[f"dec.{i}.qkv_w"], n[f"dec.{i}.o_w"], n[f"dec.{i}.ln2_g"], n[f"dec.{i}.gateup_w"],
                n[f"dec.{i}.down_w"])

    def _attention(self, qkv: torch.Tensor, b: int, s: int) -> torch.Tensor:
        d = self.cfg.dec
        return ops.attention(qkv, b, s, d.heads, d.kv_heads, d.head_dim, d.head_dim, 0, self.q_w, self.q_w + self.kv_w,
                             True, variant=self._attn_variant)

    def _layer_forward(self, i: int, x: torch.Tensor, b: int, s: int, keep: bool):
        """Same kernel sequence as lp_decoder_prefill's layer body (csrc/lp_model.cu), out of place.  keep = True returns
        the intermediates the backward needs."""
        d = self.cfg.dec
        ln1, wqkv, wo, ln2, wgu, wd = self._layer_weights(i)
        h1 = ops.rmsnorm(x, ln1, d.eps)
        qkv = ops.gemm(h1, wqkv)
        ops.rope_inplace(qkv, 0, d.heads + d.kv_heads, d.head_dim, self._cos, self._sin)
        att = self._attention(qkv, b, s)
        x2 = ops.gemm(att, wo, residual=x)
        h2 = ops.rmsnorm(x2, ln2, d.eps)
        if not keep:
            m = ops.gemm(h2, wgu, act=ops.ACT_SWIGLU)
            return ops.gemm(m, wd, residual=x2)
        gu = ops.gemm(h2, wgu)
        m = ops.swiglu_fwd(gu)
        return h1, qkv, att, x2, h2, gu, m

    def _attention_backward(self, qkv: torch.Tensor, datt: torch.Tensor, b: int, s: int, heads: Optional[int] = None,
                            kv_heads: Optional[int] = None, hd: Optional[int] = None, stride: Optional[int] = None,
                            causal: bool = True) -> torch.Tensor:
        """dqkv (layout of qkv) from datt [R, heads * hd]: per sequence and kv group, score tiles materialised per q head.
        dQ_h = dS_h K_g;  dK_g = sum_h dS_h^T Q_h;  dV_g = sum_h P_h^T dO_h (the sums run inside one GEMM over the
        concatenated contraction dimension).  qkv holds [q heads | k heads | v heads] with `stride` columns per head of
        which the first `hd` are used (decoder: 128 / 128; vision tower: 80 / 72, the pad columns keep a zero gradient)."""
        d = self.cfg.dec
        heads = d.heads if heads is None else heads
        kv_heads = d.kv_heads if kv_heads is None else kv_heads
        hd = d.head_dim if hd is None else hd
        stride = hd if stride is None else stride
        rep = heads // kv_heads
        k0, v0 = heads * stride, (heads + kv_heads) * stride
        scale = hd ** -0.5
        sp = -(-s // 8) * 8                 # leading dimensions padded to 8 elements (TMA row-stride rule)
        dqkv = torch.zeros_like(qkv) if (stride != hd) else torch.empty_like(qkv)
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        sc_full = torch.empty((s, sp), dtype=torch.float32, device=self.dev)
        dp_full = torch.empty((s, sp), dtype=torch.float32, device=self.dev)
        sc, dp = sc_full[:, :s], dp_full[:, :s]
        # the GEMM stores 16-byte groups (N % 8 == 0): a sequence length like 676 (vision tower) gets key / value
        # operands copied into sp zero-padded rows; the extra score columns [s, sp) are never read
        pad_kv = sp != s
        if pad_kv:
            kpad, vpad = torch.zeros((sp, hd), **bf), torch.zeros((sp, hd), **bf)
        P = torch.empty((s, sp), **bf)[:, :s]
        dS = torch.empty((s, sp), **bf)[:, :s]
        # concatenated-contraction operands: the gap columns [s, sp) of every block stay zero on both sides
        pt_cat = torch.zeros((s, rep * sp), **bf)
        dst_cat = torch.zeros((s, rep * sp), **bf)
        dot_cat = torch.zeros((hd, rep * sp), **bf)
        qt_cat = torch.zeros((hd, rep * sp), **bf)
        kt = torch.empty((hd, sp), **bf)[:, :s]
        for bi in range(b):
            rows = slice(bi * s, (bi + 1) * s)
            for g in range(kv_heads):
                k_g = qkv[rows, k0 + g * stride: k0 + g * stride + hd]
                v_g = qkv[rows, v0 + g * stride: v0 + g * stride + hd]
                ops.transpose(k_g, out=kt)
                if pad_kv:
                    kpad[:s].copy_(k_g)
                    vpad[:s].copy_(v_g)
                for j in range(rep):
                    h = g * rep + j
                    q_h = qkv[rows, h * stride: h * stride + hd]
                    do_h = datt[rows, h * hd:(h + 1) * hd]
                    ops.gemm(q_h, kpad if pad_kv else k_g, out=sc_full if pad_kv else sc, out_fp32=True)
                    ops.gemm(do_h, vpad if pad_kv else v_g, out=dp_full if pad_kv else dp, out_fp32=True)
                    ops.softmax_bwd(sc, dp, causal, scale, P=P, dS=dS)
                    ops.gemm(dS, kt, out=dqkv[rows, h * stride: h * stride + hd])
                    ops.transpose(P, out=pt_cat[:, j * sp: j * sp + s])
                    ops.transpose(dS, out=dst_cat[:, j * sp: j * sp + s])
                    ops.transpose(do_h, out=dot_cat[:, j * sp: j * sp + s])
                    ops.transpose(q_h, out=qt_cat[:, j * sp: j * sp + s])
                ops.gemm(dst_cat, qt_cat, out=dqkv[rows, k0 + g * stride: k0 + g * stride + hd])
                ops.gemm(pt_cat, dot_cat, out=dqkv[rows, v0 + g * stride: v0 + g * stride + hd])
        return dqkv

    # ------------------------------------------------------------------ vision tower (trainable: freeze_vision_tower=False)
    def _vit_weights(self, i: int):
        n = self.w.named
        return {k: n[f"vit.{i}.{k}"] for k in ("ln1_g", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_g", "ln2_b", "fc1_w",
                                               "fc1_b", "fc2_w", "fc2_b")}

    def _vit_layer_forward(self, i: int, x: torch.Tensor, n_img: int, keep: bool):
        """Same kernel sequence as vit_layers_impl (csrc/lp_model.cu), out of place."""
        v, L = self.cfg.vit, self._vit_weights(i)
        hs = self.w.vit_dims.head_stride
        h = ops.layernorm(x, L["ln1_g"], L["ln1_b"], v.eps)
        qkv = ops.gemm(h, L["qkv_w"], bias=L["qkv_b"])
        att = ops.attention(qkv, n_img, v.tokens, v.heads, v.heads, v.head_dim, hs, 0, v.heads * hs, 2 * v.heads * hs, False,
                            variant=self._attn_variant)
        x2 = ops.gemm(att, L["out_w"], bias=L["out_b"], residual=x)
        h2 = ops.layernorm(x2, L["ln2_g"], L["ln2_b"], v.eps)
        a = ops.gemm(h2, L["fc1_w"], bias=L["fc1_b"], act=ops.ACT_GELU_TANH)
        if not keep:
            return ops.gemm(a, L["fc2_w"], bias=L["fc2_b"], residual=x2)
        z = ops.gemm(h2, L["fc1_w"], bias=L["fc1_b"])         # pre-activation, for gelu'
        return h, qkv, att, x2, h2, z, a

    def _ln_bwd(self, gname: str, bname: str, x, gamma, dy, add=None) -> torch.Tensor:
        v = self.cfg.vit
        sg, sb = self._sink(gname), self._sink(bname)
        if sb is sg and sg is not None:      # the two scratch tensors have the same shape while accumulating
            sb = torch.empty_like(sg)
        dx = ops.layernorm_bwd(x, gamma, dy, v.eps, sg, sb, add=add)
        self._commit(gname, sg)
        self._commit(bname, sb)
        return dx

    def _vit_layer_backward(self, i: int, x: torch.Tensor, dy: torch.Tensor, n_img: int, acts=None) -> torch.Tensor:
        v, L = self.cfg.vit, self._vit_weights(i)
        hs = self.w.vit_dims.head_stride
        p = f"vit.{i}."
        h, qkv, att, x2, h2, z, a = acts if acts is not None else self._vit_layer_forward(i, x, n_img, keep=True)
        self._wgrad(p + "fc2_w", dy, a)
        self._colsum(p + "fc2_b", dy)
        da = self._dgrad(dy, L["fc2_w"])
        dz = ops.gelu_bwd(z, da, ops.ACT_GELU_TANH)
        del a, z, da
        self._wgrad(p + "fc1_w", dz, h2)
        self._colsum(p + "fc1_b", dz)
        dh2 = self._dgrad(dz, L["fc1_w"])
        del dz, h2
        dx2 = self._ln_bwd(p + "ln2_g", p + "ln2_b", x2, L["ln2_g"], dh2, add=dy)
        del dh2, x2
        self._wgrad(p + "out_w", dx2, att)
        self._colsum(p + "out_b", dx2)
        datt = self._dgrad(dx2, L["out_w"])
        if self.attn_bwd == "flash":
            dqkv = ops.attention_bwd(qkv, att, datt, n_img, v.tokens, v.heads, v.heads, v.head_dim, 0, v.heads * hs,
                                     2 * v.heads * hs, False, head_stride=hs)
        else:
            dqkv = self._attention_backward(qkv, datt, n_img, v.tokens, heads=v.heads, kv_heads=v.heads, hd=v.head_dim,
                                            stride=hs, causal=False)
        del att, datt, qkv
        self._wgrad(p + "qkv_w", dqkv, h)
        self._colsum(p + "qkv_b", dqkv)
        dh = self._dgrad(dqkv, L["qkv_w"])
        del dqkv, h
        return self._ln_bwd(p + "ln1_g", p + "ln1_b", x, L["ln1_g"], dh, add=dx2)

    def _vit_forward_train(self, images: torch.Tensor):
        """lp_vit_forward's sequence (im2col -> patch GEMM + bias + position rows -> layers -> ln_post with the shuffled
        store), keeping the im2col matrix, the layer inputs and the final residual stream."""
        v, w = self.cfg.vit, self.w.named
        n_img = images.shape[0]
        A = ops.im2col_patch14(images.to(self.dev).contiguous(), self.w.vit_dims.patch_k_pad)
        x = ops.gemm(A, w["vit.patch_w"], bias=w["vit.patch_b"], residual=w["vit.pos"], res_row_mod=v.tokens)
        xs = []
        for i in range(v.layers):
            if self.recompute == "none":
                acts = self._vit_layer_forward(i, x, n_img, keep=True)
                xs.append((x, acts))
                x = ops.gemm(acts[6], self._vit_weights(i)["fc2_w"], bias=self._vit_weights(i)["fc2_b"], residual=acts[3])
            else:
                xs.append((x, None))
                x = self._vit_layer_forward(i, x, n_img, keep=False)
        shuf = ops.layernorm(x, w["vit.post_g"], w["vit.post_b"], v.eps, shuffle_side=v.grid)
        return shuf, (A, xs, x, n_img)

    def _vit_backward(self, saved, dshuf: torch.Tensor, reduce):
        v, w = self.cfg.vit, self.w.named
        A, xs, x_last, n_img = saved
        dx = self._ln_bwd("vit.post_g", "vit.post_b", x_last, w["vit.post_g"], ops.pixel_unshuffle(dshuf, v.grid))
        for i in reversed(range(v.layers)):
            x_in, acts = xs.pop()
            dx = self._vit_layer_backward(i, x_in, dx, n_img, acts)
        sink = self._sink("vit.pos")                    # the forward adds pos[row % tokens] to every crop
        ops.rowmod_sum(dx, v.tokens, sink)
        self._commit("vit.pos", sink)
        self._colsum("vit.patch_b", dx)
        self._wgrad("vit.patch_w", dx, A)               # [hidden, patch_k_pad]; the K-padding columns of A are zero

    def _layer_backward(self, i: int, x: torch.Tensor, dy: torch.Tensor, b: int, s: int, acts=None) -> torch.Tensor:
        d = self.cfg.dec
        ln1, wqkv, wo, ln2, wgu, wd = self._layer_weights(i)
        h1, qkv, att, x2, h2, gu, m = acts if acts is not None else self._layer_forward(i, x, b, s, keep=True)
        # MLP
        self._wgrad(f"dec.{i}.down_w", dy, m)
        dm = self._dgrad(dy, wd)
        del m
        dgu = ops.swiglu_bwd(gu, dm)
        del gu, dm
        self._wgrad(f"dec.{i}.gateup_w", dgu, h2)
        dh2 = self._dgrad(dgu, wgu)
        del dgu, h2
        dx2 = self._rms_bwd(f"dec.{i}.ln2_g", x2, ln2, dh2, add=dy)
        del dh2, x2
        # attention
        self._wgrad(f"dec.{i}.o_w", dx2, att)
        datt = self._dgrad(dx2, wo)
        if self.attn_bwd == "flash":      # tcgen05 flash-style backward (lp_attn_bwd.cu)
            dqkv = ops.attention_bwd(qkv, att, datt, b, s, d.heads, d.kv_heads, d.head_dim, 0, self.q_w,
                                     self.q_w + self.kv_w, True)
        else:                             # materialised score tiles per head (lp_softmax_bwd)
            dqkv = self._attention_backward(qkv, datt, b, s)
        del att, datt, qkv
        ops.rope_inplace(dqkv, 0, d.heads + d.kv_heads, d.head_dim, self._cos, self._neg_sin)   # adjoint rotation
        self._wgrad(f"dec.{i}.qkv_w", dqkv, h1)
        dh1 = self._dgrad(dqkv, wqkv)
        del dqkv, h1
        return self._rms_bwd(f"dec.{i}.ln1_g", x, ln1, dh1, add=dx2)

    # ------------------------------------------------------------------ gradient reduction
    def _reduce_bucket(self, k: int, handles: list):
        if self.world > 1:
            a, e = self.buckets[k]
            handles.append(torch.distributed.all_reduce(self.flat_grad[a:e], group=self.pg, async_op=True))

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, tokens, labels, loss_mask, images):
        """One iteration.  Each argument is a tensor (one micro-batch) or a list of them (gradient accumulation over
        get_num_microbatches() micro-batches, as Megatron's forward_backward_func does: every micro-batch loss is its own
        masked mean, divided by the number of micro-batches).  Returns (loss, grad_norm) as Python floats — the mean of the
        micro-batch losses of this rank, and the global gradient norm before clipping; one device -> host read."""
        if isinstance(tokens, (list, tuple)):
            micro = list(zip(tokens, labels, loss_mask, images))
        else:
            micro = [(tokens, labels, loss_mask, images)]
        self.iteration += 1
        self.flat_grad.zero_()
        handles: list = []
        losses = []
        for k, mb in enumerate(micro):
            self._accum = k > 0
            losses.append(self._forward_backward(*mb, n_micro=len(micro), last=k == len(micro) - 1, handles=handles))
        self._accum = False
        for h in handles:
            h.wait()

        # ---- clip + AdamW
        w, hp = self.w.named, self.hp
        # Megatron sets the rate from the number of COMPLETED steps (OptimizerParamScheduler.step runs after optimizer.step,
        # training.py:590-596), so iteration k uses lr(k - 1) and the first warm-up step has lr = 0; Adam's bias correction
        # keeps the 1-based count
        lr = hp["lr"] if hp["decay"] <= 0 else cosine_lr(self.iteration - 1, hp["lr"], hp["min_lr"], hp["warmup"], hp["decay"])
        ops.grad_norm([self.flat_grad], hp["clip"], self._norm_ws, self._norm_out)
        for n in self.names:
            model_t = w[n]
            g = self.grad(n).reshape(-1)
            m = self._view(self.flat_m, self.layout, n).reshape(-1)
            v = self._view(self.flat_v, self.layout, n).reshape(-1)
            if model_t.dtype == torch.bfloat16:
                p = self._view(self.flat_master, self.master_layout, n).reshape(-1)
                ops.adamw_step(p, g, m, v, model_t, self.iteration, lr, hp["beta1"], hp["beta2"], hp["eps"], hp["wd"],
                               self._norm_out)
            else:
                ops.adamw_step(model_t.reshape(-1), g, m, v, None, self.iteration, lr, hp["beta1"], hp["beta2"], hp["eps"],
                               hp["wd"], self._norm_out)
        # the engine's gain-folded copies of these weights (fused-norm prefill) are stale now; re-derived on next use
        if not self.freeze_llm:
            self.w.folded_dirty.add("dec")
        if not self.freeze_vit:
            self.w.folded_dirty.add("vit")
        loss = torch.stack(losses).mean()
        if self.world > 1:           # the logged loss is the data-parallel average (average_losses_across_data_parallel_group,
            loss = loss.clone()      # pretrain_megatron_llava.py:119)
            torch.distributed.all_reduce(loss, group=self.pg)
            loss /= self.world
        out = torch.cat([loss.reshape(1), self._norm_out[:1]]).tolist()
        self.skipped_last = not math.isfinite(out[1])
        if self.skipped_last:        # inf / NaN gradients: the update kernels returned without touching anything
            self.iteration -= 1      # (Megatron's found-inf path: the step does not count, optimizer.py:313-340)
        return out[0], out[1]

    def _forward_backward(self, tokens, labels, loss_mask, images, n_micro: int, last: bool, handles: list):
        """Forward, loss and backward of one micro-batch; gradients are written (first micro-batch) or added (later ones)
        into the flat buffer; on the last micro-batch every bucket's all-reduce is issued as soon as it is complete."""
        cfg, d, w = self.cfg, self.cfg.dec, self.w.named
        self._attn_variant = 336     # v3 attention (+ its persistent form for the vision tower's short sequences)
        b, s = tokens.shape
        if s % 8:
            raise ValueError(f"sequence length {s} must be a multiple of 8")
        R = b * s
        n_img_rows = int((tokens == cfg.megatron_image_index).sum())      # counted where the batch lives (the host): no device sync
        ids = tokens.reshape(-1).to(self.dev, torch.long).contiguous()

        # ---- vision tower (frozen) + projector, keeping what the projector's backward needs
        have_images = images is not None and images.shape[-1] > 100 and n_img_rows > 0
        feats = shuf = z1 = a1 = None
        vit_saved = None
        if not have_images:
            # dummy [1, 10] image (pretrain_megatron_llava.py:188): the reference skips the scatter (vlm_model.py:518), so any
            # -200 slot keeps the embedding of clamp(id, 0) = row 0 and its gradient flows there
            ids = ids.clamp(min=0)
        if have_images:
            if self.freeze_vit:
                shuf = self.eng.vit(images, shuffle=True)
            else:
                shuf, vit_saved = self._vit_forward_train(images)
            z1 = ops.gemm(shuf, w["proj.w1"], bias=w["proj.b1"])
            a1 = ops.gemm(shuf, w["proj.w1"], bias=w["proj.b1"], act=self.w.proj.act)
            feats = ops.gemm(a1, w["proj.w2"], bias=w["proj.b2"])
            if feats.shape[0] != n_img_rows:
                raise AssertionError("The number of image features doesn't match the number of image tokens. "
                                     f"{feats.shape[0]} != {n_img_rows}")
        # ---- scatter into the embeddings (vlm_model.py:516-533), all sequences packed [R, hidden]
        x = torch.empty((R, d.hidden), dtype=torch.bfloat16, device=self.dev)
        self.eng.merge(ids, feats, cfg.megatron_image_index, 1, out=x)
        pos = torch.arange(s, dtype=torch.int32, device=self.dev).repeat(b)
        self._cos, self._sin = ops.rope_table(pos, w["dec.inv_freq"])
        self._neg_sin = -self._sin

        # ---- decoder forward; kept: every layer's input (recompute = 'full') or all of its activations ('none')
        xs = []
        for i in range(d.layers):
            if self.recompute == "none":
                acts = self._layer_forward(i, x, b, s, keep=True)
                xs.append((x, acts))
                x = ops.gemm(acts[6], w[f"dec.{i}.down_w"], residual=acts[3])       # m W_down^T + x2
            else:
                xs.append((x, None))
                x = self._layer_forward(i, x, b, s, keep=False)
        hN = ops.rmsnorm(x, w["dec.final_g"], d.eps)
        logits = ops.gemm(hN, w["dec.lm_head"], out_fp32=True)                    # [R, vocab_padded] fp32

        # ---- loss and its gradient
        tgt, wt = shifted_targets(labels.to(self.dev), loss_mask.to(self.dev), self.world * n_micro)
        loss_rows, dlogits = ops.ce_fwd_bwd(logits, tgt, wt, d.vocab)
        del logits
        loss = torch.dot(loss_rows, wt) * (self.world * n_micro)     # this micro-batch's masked mean (what is reported)

        # ---- backward (buckets of the flat gradient buffer are reduced in the order they complete)
        def reduce(k):
            if last:
                self._reduce_bucket(k, handles)

        self._wgrad("dec.lm_head", dlogits, hN)
        dh = self._dgrad(dlogits, w["dec.lm_head"])
        del dlogits
        dx = self._rms_bwd("dec.final_g", x, w["dec.final_g"], dh)
        del dh, hN
        bucket = 0
        if not self.freeze_llm:
            reduce(bucket)      # 'dec.lm_head' + 'dec.final_g'
            bucket += 1
        for i in reversed(range(d.layers)):
            x_in, acts = xs.pop()
            dx = self._layer_backward(i, x_in, dx, b, s, acts)
            del acts
            if not self.freeze_llm:
                reduce(bucket)
                bucket += 1
        # ---- merge backward: text rows feed the embedding table, image rows feed the projector
        if not self.freeze_llm:
            ops.embed_bwd(dx, ids, cfg.megatron_image_index, self.grad("embed"))     # adds: accumulates by construction
            reduce(bucket)
            bucket += 1
        if have_images:
            idx = torch.nonzero(ids == cfg.megatron_image_index).reshape(-1).to(torch.int32)
            dfe = ops.gather_rows(dx, idx)
            self._wgrad("proj.w2", dfe, a1)
            self._colsum("proj.b2", dfe)
            da1 = self._dgrad(dfe, w["proj.w2"])
            dz1 = ops.gelu_bwd(z1, da1, self.w.proj.act)
            self._wgrad("proj.w1", dz1, shuf)
            self._colsum("proj.b1", dz1)
        reduce(bucket)
        if not self.freeze_vit:
            bucket += 1
            if have_images:
                self._vit_backward(vit_saved, self._dgrad(dz1, w["proj.w1"]), reduce)
            reduce(bucket)
        assert bucket == len(self.buckets) - 1
        return loss
