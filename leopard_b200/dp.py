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


_HDR_COLS = 8       # owner, ids length, n_crops, C, H, W, pixel dtype code, reserved
_DTYPES = [torch.float32, torch.bfloat16, torch.float16]


def _header(requests, bins) -> torch.Tensor:
    owner = {k: r for r, idxs in enumerate(bins) for k in idxs}
    rows = []
    for k, (ids, pix) in enumerate(requests):
        shp = (0, 0, 0, 0) if pix is None else tuple(pix.shape)
        rows.append([owner[k], int(ids.numel()), *shp, 0 if pix is None else _DTYPES.index(pix.dtype), 0])
    return torch.tensor(rows, dtype=torch.int64).reshape(-1, _HDR_COLS)


def scatter_requests(requests: Optional[List[Tuple[torch.Tensor, Optional[torch.Tensor]]]], device, src: int = 0,
                     n_requests: Optional[int] = None, bins: Optional[List[List[int]]] = None):
    """Rank `src` holds `requests` = [(input_ids [L] int64, pixel_values [N,3,H,W] or None)] (host or device tensors);
    every rank returns (its own requests on `device`, their global indices).  One small int64 header broadcast carries
    the cost-balanced assignment and the shapes (N and L vary per request), then the payloads travel as ONE batch of
    point-to-point transfers (NCCL send/recv over NVLink; gloo on CPU).  `n_requests` lets non-source ranks skip the
    count broadcast when the queue length is known; `bins` overrides the assignment (e.g. one request per rank)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.device(device)
    if n_requests is None:
        cnt = torch.tensor([len(requests) if rank == src else 0], dtype=torch.int64, device=device)
        dist.broadcast(cnt, src=src)
        n_requests = int(cnt.item())
    if rank == src:
        if bins is None:
            costs = [request_cost(0 if p is None else p.shape[0], i.numel() + (0 if p is None else p.shape[0] * 168))
                     for i, p in requests]
            bins = assign(costs, world)
        hdr = _header(requests, bins).to(device)
    else:
        hdr = torch.empty((n_requests, _HDR_COLS), dtype=torch.int64, device=device)
    dist.broadcast(hdr, src=src)
    hdr = hdr.cpu().tolist()         # shapes are needed on the host to size the receive buffers (one small D2H)
    mine = [k for k, row in enumerate(hdr) if row[0] == rank]
    ops, out = [], []
    if rank == src:
        for k, row in enumerate(hdr):
            ids, pix = requests[k]
            ids = ids.reshape(-1).to(device, non_blocking=True)
            pix = None if pix is None else pix.to(device, non_blocking=True)
            if row[0] == src:
                out.append((ids, pix))
                continue
            ops.append(dist.P2POp(dist.isend, ids.contiguous(), row[0]))
            if pix is not None:
                ops.append(dist.P2POp(dist.isend, pix.contiguous(), row[0]))
    else:
        for k in mine:
            _, L, n, c, h, w, dt, _ = hdr[k]
            ids = torch.empty(L, dtype=torch.long, device=device)
            ops.append(dist.P2POp(dist.irecv, ids, src))
            pix = None
            if n > 0:
                pix = torch.empty((n, c, h, w), dtype=_DTYPES[dt], device=device)
                ops.append(dist.P2POp(dist.irecv, pix, src))
            out.append((ids, pix))
    if ops:
        for w_ in dist.batch_isend_irecv(ops):
            w_.wait()
    return out, mine


def gather_results(local: torch.Tensor, mine: Sequence[int], total: int, cap: Optional[int] = None) -> torch.Tensor:
    """local: [len(mine), W] results of this rank's requests (e.g. last-token logits or token ids) -> [total, W] on
    every rank, rows in global request order.  Two all-gathers (padded index lists, padded rows); no pickling.
    `cap` = the largest per-rank request count when the caller knows it (saves one scalar all-reduce + host sync)."""
    world = dist.get_world_size()
    dev = local.device
    if cap is None:
        capt = torch.tensor([len(mine)], dtype=torch.int64, device=dev)
        dist.all_reduce(capt, op=dist.ReduceOp.MAX)
        cap = int(capt.item())
    cap = max(cap, 1)
    W = local.shape[1]
    idx = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    idx[: len(mine)] = torch.tensor(list(mine), dtype=torch.int64, device=dev)
    pad = torch.zeros((cap, W), dtype=local.dtype, device=dev)
    pad[: local.shape[0]] = local
    flat_idx = torch.empty((world * cap,), dtype=torch.int64, device=dev)        # rank-major concatenation along dim 0
    flat_rows = torch.empty((world * cap, W), dtype=local.dtype, device=dev)
    dist.all_gather_into_tensor(flat_idx, idx)
    dist.all_gather_into_tensor(flat_rows, pad)
    out = torch.zeros((total, W), dtype=local.dtype, device=dev)
    keep = flat_idx >= 0
    out[flat_idx[keep]] = flat_rows[keep]
    return out
