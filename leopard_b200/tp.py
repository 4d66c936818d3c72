"""Tensor parallel + Megatron sequence parallel decoder: the reference's real training / inference sharding.

Reference: `--tensor-model-parallel-size 8 --sequence-parallel` (Pai-Megatron-Patch/examples/llava/train_multiimg_llava_siglip.sh:
57-71,151-153); ColumnParallelLinear / RowParallelLinear with sequence parallelism
(Megatron-LM-240603/megatron/core/tensor_parallel/layers.py:387-454: all-gather along the sequence before the QKV / fc1 GEMM,
reduce-scatter along the sequence after the O / fc2 GEMM; mappings.py:119-199), and the TP-sharded checkpoint layout written /
read by toolkits/model_checkpoints_convertor/llava/hf2megatron_llava.py (cat-0 for column-parallel, cat-1 for row-parallel
:1176-1184; QKV rows grouped per kv group :1226-1258 and megatron_patch/model/llava/transformer.py:713-738; dense_h_to_4h of every TP
shard = [gate chunk ; up chunk] :1287-1295).

What is here
  * shard_megatron_state / merge_megatron_shards   TP = 1 <-> TP = n checkpoint dictionaries (decoder, embedding, output layer,
    projector, tower), exact inverses, same per-key rules as the converter;
  * shard_decoder_weights   one rank's decoder shard in the kernel-native layout (fused [q | k | v] rows of ITS heads, 32-row
    gate / up granules of ITS ffn slice, K-sliced o / down), straight from HF-named tensors;
  * TPDecoder   the layer loop for one rank: activations are row-sharded between the TP regions (sequence parallel),
        RMSNorm(local rows) -> ALL-GATHER rows -> QKV GEMM (column shard) -> RoPE -> attention over the rank's heads ->
        O GEMM (K shard, partial sums) -> REDUCE-SCATTER rows (+ residual) -> RMSNorm -> ALL-GATHER -> gate/up GEMM + SwiGLU ->
        down GEMM (partial) -> REDUCE-SCATTER (+ residual); final norm; vocabulary-parallel lm_head + all-gather of the logits.
    `comm="nccl"`: torch.distributed collectives (the baseline, also what the gloo CPU tests run);
    `comm="fused"`: the all-gather is done INSIDE the consuming GEMM kernel over NVLink peer memory (lp_gemm_allgather: the
    epilogue warps pull the peers' rows while the first tiles — the rank's own rows — are being multiplied) and the
    reduce-scatter leaves the GEMM's epilogue as TMA stores of partial tiles into the owner rank's memory, tile by tile
    (lp_gemm_reduce_scatter), followed by a fixed-order reduction on the owner.  See csrc/lp_tp.cu and DESIGN.md §9.
Every column-parallel output is bit-identical to the same columns of the TP = 1 GEMM (same kernel, same K order); row-parallel
sums differ from TP = 1 only by the bf16 rounding of the partial sums (as in the reference, whose reduce-scatter sums bf16).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .config import LeopardConfig

LM = "language_model.model."

# ---------------------------------------------------------------------------------------------------------------------------
# checkpoint sharding (Megatron layout, flattened `language_model` keys as in native_weights.megatron_to_hf)
# ---------------------------------------------------------------------------------------------------------------------------
_ROW_PARALLEL = ("self_attention.dense.weight", "mlp.dense_4h_to_h.weight", "mm_projector.encoder.linear_fc2.weight",
                 "self_attention.linear_proj.weight", "mlp.linear_fc2.weight")
_COL_PARALLEL = ("self_attention.query_key_value.weight", "embedding.word_embeddings.weight", "output_layer.weight",
                 "mm_projector.encoder.linear_fc1.weight", "mm_projector.encoder.linear_fc1.bias",
                 "self_attention.linear_qkv.weight", "self_attention.linear_qkv.bias", "mlp.linear_fc1.weight",
                 "mlp.linear_fc1.bias")


def _kind(key: str) -> str:
    if key.endswith("mlp.dense_h_to_4h.weight"):
        return "gateup"
    if any(key.endswith(s) for s in _ROW_PARALLEL):
        return "row"
    if any(key.endswith(s) for s in _COL_PARALLEL):
        return "col"
    return "replicated"


def shard_megatron_state(lm: Dict[str, torch.Tensor], tp: int) -> List[Dict[str, torch.Tensor]]:
    """TP = 1 state -> `tp` per-rank states (what `mp_rank_0r/model_optim_rng.pt` holds under model/language_model)."""
    out: List[Dict[str, torch.Tensor]] = [dict() for _ in range(tp)]
    for k, t in lm.items():
        kind = _kind(k)
        if kind == "replicated":
            parts = [t] * tp
        elif kind == "gateup":      # TP = 1 tensor is [gate ; up]; shard r is [gate chunk r ; up chunk r]
            g, u = t.chunk(2, dim=0)
            parts = [torch.cat([gc, uc], 0) for gc, uc in zip(g.chunk(tp, 0), u.chunk(tp, 0))]
        else:
            parts = list(t.chunk(tp, dim=0 if kind == "col" else 1))
        if len(parts) != tp:
            raise ValueError(f"{k}: shape {tuple(t.shape)} does not split {tp} ways")
        for r in range(tp):
            out[r][k] = parts[r]
    return out


def merge_megatron_shards(shards: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Inverse of shard_megatron_state (the converter's merge: hf2megatron_llava.py:1176-1184,1287-1295)."""
    tp = len(shards)
    out: Dict[str, torch.Tensor] = {}
    for k in shards[0]:
        kind = _kind(k)
        ts = [s[k] for s in shards]
        if kind == "replicated":
            out[k] = ts[0]
        elif kind == "gateup":
            halves = [t.chunk(2, dim=0) for t in ts]
            out[k] = torch.cat([h[0] for h in halves] + [h[1] for h in halves], 0)
        else:
            out[k] = torch.cat(ts, dim=0 if kind == "col" else 1)
    assert all(len(s) == len(shards[0]) for s in shards) and tp >= 1
    return out


# ---------------------------------------------------------------------------------------------------------------------------
# one rank's decoder weights in kernel layout
# ---------------------------------------------------------------------------------------------------------------------------
def _interleave_gate_up(g: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    f, h = g.shape
    return torch.cat([g.reshape(f // 32, 1, 32, h), u.reshape(f // 32, 1, 32, h)], 1).reshape(2 * f, h)


def shard_decoder_weights(cfg: LeopardConfig, get: Callable[[str], torch.Tensor], tp: int, rank: int, device) -> dict:
    """HF-named tensors -> this rank's decoder shard.  q head j lives on rank j // (heads / tp) together with its kv group
    (heads are contiguous per kv group in both layouts), ffn channel c on rank c // (ffn / tp)."""
    d = cfg.dec
    if d.kv_heads % tp or d.ffn % (32 * tp) or d.vocab % tp:
        raise ValueError(f"tp={tp} must divide kv_heads={d.kv_heads}, ffn/32 and the vocabulary")
    hq, hk, hd = d.heads // tp, d.kv_heads // tp, d.head_dim
    fl = d.ffn // tp
    bf = lambda t: t.to(device, torch.bfloat16).contiguous()  # noqa: E731
    f32 = lambda t: t.to(device).to(torch.bfloat16).float().contiguous()  # noqa: E731
    W = {"layers": [], "tp": tp, "rank": rank}
    for i in range(d.layers):
        b = f"{LM}layers.{i}."
        q = get(b + "self_attn.q_proj.weight")[rank * hq * hd:(rank + 1) * hq * hd]
        k = get(b + "self_attn.k_proj.weight")[rank * hk * hd:(rank + 1) * hk * hd]
        v = get(b + "self_attn.v_proj.weight")[rank * hk * hd:(rank + 1) * hk * hd]
        g = get(b + "mlp.gate_proj.weight")[rank * fl:(rank + 1) * fl]
        u = get(b + "mlp.up_proj.weight")[rank * fl:(rank + 1) * fl]
        W["layers"].append({
            "ln1_g": f32(get(b + "input_layernorm.weight")), "ln2_g": f32(get(b + "post_attention_layernorm.weight")),
            "qkv_w": bf(torch.cat([q, k, v], 0)),
            "o_w": bf(get(b + "self_attn.o_proj.weight")[:, rank * hq * hd:(rank + 1) * hq * hd]),
            "gateup_w": bf(_interleave_gate_up(g.to(device), u.to(device))),
            "down_w": bf(get(b + "mlp.down_proj.weight")[:, rank * fl:(rank + 1) * fl]),
        })
    W["final_g"] = f32(get(LM + "norm.weight"))
    vl = d.vocab // tp
    W["lm_head"] = bf(get("language_model.lm_head.weight")[rank * vl:(rank + 1) * vl])          # vocabulary parallel
    from .native_weights import rope_inv_freq
    W["inv_freq"] = rope_inv_freq(d).to(device).contiguous()
    return W


# ---------------------------------------------------------------------------------------------------------------------------
# the layer loop of one rank
# ---------------------------------------------------------------------------------------------------------------------------
class TPDecoder:
    def __init__(self, cfg: LeopardConfig, weights: dict, group=None, comm: str = "nccl", fused=None):
        self.cfg, self.w, self.group, self.comm = cfg, weights, group, comm
        self.tp, self.rank = weights["tp"], weights["rank"]
        if comm not in ("nccl", "fused"):
            raise ValueError("comm must be 'nccl' or 'fused'")
        self.fused = fused          # leopard_b200.tp_fused.FusedComm (symmetric buffers + kernels) when comm == "fused"
        if comm == "fused" and fused is None:
            raise _lib.NativeError("comm='fused' needs a FusedComm (NVLink peer memory); there is no fallback")

    # -- collectives (baseline path) ----------------------------------------------------------------------------------------
    def _all_gather_rows(self, local: torch.Tensor) -> torch.Tensor:
        full = torch.empty((local.shape[0] * self.tp, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous(), group=self.group)
        return full

    def _reduce_scatter_rows(self, partial: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        """sum over ranks of partial [S, H] (bf16, like the reference's reduce-scatter) -> own rows + residual."""
        sl = partial.shape[0] // self.tp
        if dist.get_backend(self.group) == "gloo":       # gloo has no reduce_scatter: all-reduce in fp32 of bf16 values, slice
            tmp = partial.float()
            dist.all_reduce(tmp, group=self.group)
            mine = tmp[self.rank * sl:(self.rank + 1) * sl]
        else:
            out = torch.empty((sl, partial.shape[1]), dtype=partial.dtype, device=partial.device)
            dist.reduce_scatter_tensor(out, partial.contiguous(), group=self.group)
            mine = out.float()
        return (mine + residual.float()).to(torch.bfloat16)

    # -- forward --------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x_local: torch.Tensor, S: int, last_only: bool = True) -> torch.Tensor:
        """x_local bf16 [S / tp, H]: this rank's rows of the merged embeddings (one sequence, positions 0 .. S-1).
        Returns fp32 logits [1 or S, vocab] (gathered over the vocabulary shards) on every rank."""
        d, w, tp = self.cfg.dec, self.w, self.tp
        if S % tp or x_local.shape[0] != S // tp:
            raise ValueError(f"sequence parallelism needs S % tp == 0 and S / tp local rows (S={S}, tp={tp})")
        hq, hk, hd = d.heads // tp, d.kv_heads // tp, d.head_dim
        q_w, kv_w = hq * hd, hk * hd
        dev = x_local.device
        cs, sn = ops.rope_table(torch.arange(S, dtype=torch.int32, device=dev), w["inv_freq"])
        x = x_local
        for L in w["layers"]:
            # sequence-parallel region: RMSNorm on the local rows only
            if self.comm == "fused":
                ops.rmsnorm(x, L["ln1_g"], d.eps, out=self.fused.local_rows())
                qkv = self.fused.allgather_gemm(L["qkv_w"])                          # all-gather inside the GEMM kernel
            else:
                h = ops.rmsnorm(x, L["ln1_g"], d.eps)
                qkv = ops.gemm(self._all_gather_rows(h), L["qkv_w"])                # [S, (hq + 2 hk) * hd]
            ops.rope_inplace(qkv, 0, hq + hk, hd, cs, sn)
            att = ops.attention(qkv, 1, S, hq, hk, hd, hd, 0, q_w, q_w + kv_w, True)           # [S, hq * hd]
            if self.comm == "fused":
                x = self.fused.gemm_reduce_scatter(att, L["o_w"], x)
            else:
                x = self._reduce_scatter_rows(ops.gemm(att, L["o_w"]), x)
            if self.comm == "fused":
                ops.rmsnorm(x, L["ln2_g"], d.eps, out=self.fused.local_rows())
                m = self.fused.allgather_gemm(L["gateup_w"], act=ops.ACT_SWIGLU)
                x = self.fused.gemm_reduce_scatter(m, L["down_w"], x)
            else:
                h = ops.rmsnorm(x, L["ln2_g"], d.eps)
                m = ops.gemm(self._all_gather_rows(h), L["gateup_w"], act=ops.ACT_SWIGLU)       # [S, ffn / tp]
                x = self._reduce_scatter_rows(ops.gemm(m, L["down_w"]), x)
        # final norm on local rows; the rows needed for logits are gathered, the lm_head is vocabulary-parallel
        h = ops.rmsnorm(x, w["final_g"], d.eps)
        full = self._all_gather_rows(h)
        rows = full[-1:].contiguous() if last_only else full
        logits_local = ops.gemm(rows, w["lm_head"], out_fp32=True)                # [rows, vocab / tp]
        parts = [torch.empty_like(logits_local) for _ in range(tp)]
        dist.all_gather(parts, logits_local.contiguous(), group=self.group)
        return torch.cat(parts, 1)
