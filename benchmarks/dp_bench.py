"""bench.py's multi-GPU leg: a queue of independent prefill requests resident on rank 0 is scattered over NCCL (cost-balanced,
leopard_b200/dp.py), every rank prefills its share as ONE varlen batch on the native path, and the greedy tokens are gathered
— scatter and gather INSIDE the timed step, per-phase times from CUDA events (max over ranks)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from leopard_b200 import dp


class DpHarness:
    def __init__(self, eng, dev, world: int, rank: int):
        self.eng, self.dev, self.world, self.rank = eng, dev, world, rank
        self.queue, self.n, self.bins, self.events, self.last = None, 0, None, [], None

    def set_queue(self, requests: Optional[list], n: int, bins: Optional[List[List[int]]] = None):
        """requests (rank 0 only): [(ids [L] int64, pixel_values [N,3,364,364] bf16)] already on rank 0's device."""
        self.queue, self.n, self.bins, self.events = requests, n, bins, []
        self.cap = None if bins is None else max(len(b) for b in bins)

    def step(self):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        if self.world > 1:
            local, mine = dp.scatter_requests(self.queue, self.dev, 0, n_requests=self.n, bins=self.bins)
        else:
            local, mine = self.queue, list(range(self.n))
        ev[1].record()
        if local:
            # HF convention: one placeholder per crop, so the slot counts are known from the shapes (no device sync per request)
            logits, _, _ = self.eng.prefill(local, "last", slots=[0 if p is None else p.shape[0] for _, p in local])
            tok = logits.argmax(-1, keepdim=True)
        else:
            tok = torch.zeros((0, 1), dtype=torch.int64, device=self.dev)
        ev[2].record()
        self.last = dp.gather_results(tok, mine, self.n, cap=self.cap) if self.world > 1 else tok
        ev[3].record()
        self.events.append(ev)
        return mine

    def phase_ms(self):
        """(scatter, compute, gather) ms per step: mean over the recorded steps on this rank, MAX over ranks."""
        torch.cuda.synchronize()
        k = max(len(self.events), 1)
        t = torch.tensor([sum(e[i].elapsed_time(e[i + 1]) for e in self.events) / k for i in range(3)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        self.events = []
        return [float(x) for x in t.tolist()]


# a mixed queue for the strong-scaling leg: 8 requests, 1 .. 28 crops ("8/4/2/1 samples per GPU as a varlen batch")
STRONG_QUEUE = [[(896, 1152)] * 4, [(1344, 1344)], [(896, 1152)] * 2, [(364, 364)], [(1344, 1344)] * 2, [(896, 1152)] * 4,
                [(728, 1092)], [(1092, 1456)] * 3]


def make_strong_queue(cfg, dev, n_text: int = 64):
    from leopard_b200.preprocess import plan_crops
    reqs, meta = [], []
    for k, sizes in enumerate(STRONG_QUEUE):
        n = sum(plan_crops(sizes))
        g = torch.Generator().manual_seed(100 + k)
        pix = (torch.rand((n, 3, cfg.vit.image_size, cfg.vit.image_size), generator=g) * 2 - 1).to(torch.bfloat16)
        text = torch.randint(0, 128000, (n_text,), generator=g)
        ids = torch.cat([text[:24], torch.full((n,), cfg.image_token_index), text[24:]])
        S = ids.numel() + n * (cfg.tokens_per_crop - 1)
        reqs.append((ids.to(dev), pix.to(dev)))
        meta.append({"images": len(sizes), "crops": n, "S": S, "gflop": round(dp.request_cost(n, S), 1)})
    return reqs, meta
