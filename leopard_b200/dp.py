"""Data-parallel sharding of independent prefill requests over the GPUs of one node.

The path has no exchange step: each request (its crops, projector rows and decoder sequence) depends on nothing else,
and the reference itself runs one process per GPU on contiguous index shards and merges result files
(evaluations/models/run_eval_llava_siglip_multiimg.sh:9-13, evaluations/models/eval_utils.py:84-89).  Here:
  * weights are replicated (17 GB bf16 per GPU);
  * requests are assigned by COST (crop counts vary 1..50 per request, so contiguous sharding load-imbalances);
  * the only collectives are the batch scatter (rank 0 -> owners, size-prefixed point-to-point) and the gather of the
    per-request results — NCCL over NVLink when the tensors are CUDA, gloo on CPU (tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def request_cost(n_crops: int, seq_len: int) -> float:
    """GFLOP model of one request: 625.6 per crop + 13.96 per token + 2.62e-4 * S^2."""
    return 625.6 * n_crops + 13.96 * seq_len + 2.62144e-4 * seq_len * seq_len


def assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time greedy bin packing; deterministic (ties -> lower index / lower rank)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += costs[i]
    for b in bins:
        b.sort()
    return bins


def contiguous(n: int, world: int) -> List[List[int]]:
    """The reference's own split (eval_utils.py:84-89 `split_shard`): contiguous index ranges."""
    per = (n + world - 1) // world
    return [list(range(r * per, min(n, (r + 1) * per))) for r in range(world)]


def scatter_requests(requests: Optional[List[Tuple[torch.Tensor, Optional[torch.Tensor]]]], device, src: int = 0):
    """Rank `src` holds `requests` = [(input_ids [L] int64, pixel_values [N,3,H,W] or None)]; every rank returns
    (its own requests on `device`, their global indices).  Shapes travel first (int64 header), then payloads."""
    rank, world = dist.get_rank(), dist.get_world_size()
    plan_box = [None]
    if rank == src:
        costs = [request_cost(0 if p is None else p.shape[0], i.numel() + (0 if p is None else p.shape[0] * 168))
                 for i, p in requests]
        bins = assign(costs, world)
        meta = [(int(i.numel()), None if p is None else tuple(p.shape), None if p is None else str(p.dtype))
                for i, p in requests]
        plan_box = [(bins, meta)]
    dist.broadcast_object_list(plan_box, src=src)
    bins, meta = plan_box[0]
    mine = bins[rank]
    out = []
    if rank == src:
        pending = []
        for r, idxs in enumerate(bins):
            for k in idxs:
                ids, pix = requests[k]
                if r == src:
                    continue
                pending.append(dist.isend(ids.to(device).contiguous(), dst=r))
                if pix is not None:
                    pending.append(dist.isend(pix.to(device).contiguous(), dst=r))
        for k in mine:
            ids, pix = requests[k]
            out.append((ids.to(device), None if pix is None else pix.to(device)))
        for p in pending:
            p.wait()
    else:
        for k in mine:
            L, pshape, pdtype = meta[k]
            ids = torch.empty(L, dtype=torch.long, device=device)
            dist.recv(ids, src=src)
            pix = None
            if pshape is not None:
                pix = torch.empty(pshape, dtype=getattr(torch, pdtype.split(".")[-1]), device=device)
                dist.recv(pix, src=src)
            out.append((ids, pix))
    return out, mine


def gather_results(local: torch.Tensor, mine: Sequence[int], total: int) -> torch.Tensor:
    """local: [len(mine), W] results of this rank's requests (e.g. last-token logits or token ids) -> [total, W] on
    every rank, rows in global request order."""
    world = dist.get_world_size()
    counts = [None] * world
    dist.all_gather_object(counts, list(mine))
    cap = max(len(c) for c in counts)
    W = local.shape[1]
    pad = torch.zeros((cap, W), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.zeros((total, W), dtype=local.dtype, device=local.device)
    for r, idxs in enumerate(counts):
        for j, k in enumerate(idxs):
            out[k] = bufs[r][j]
    return out
