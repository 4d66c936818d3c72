"""CPU arm of the bench: the reference's own (torch-cpu, fp32) implementation of the prefill path = the oracle port
(oracle/model.py — neither of the reference's two implementations imports in this image), timed on the GPU
box's host cores on a BOUNDED SAMPLE of the benchmark workload.

The sample is a proportional slice of the request at FULL DEPTH (27 ViT + 32 decoder layers, real widths): `n_crops` of the
workload's crops and the matching share of its merged rows (cfg3: 1 of 28 crops -> 169 image rows + 2 text ids = 171 of
4796 rows).  Every step is a complete, un-extrapolated forward of that small request; tokens/s = sample rows / step time.
What the slice leaves out is stated with the number: causal attention is quadratic in S (6.7 % of cfg3's FLOPs, 0.3 % of the
sample's), so the sample flatters the CPU by a few percent per token.

Weights: one random fp32 tensor per distinct SHAPE, shared by all layers (dense CPU timing does not depend on values; the
8.5 G-parameter fp32 model would need 34 GB of host memory and minutes of generation).
"""
from __future__ import annotations

import os
import time

import torch


class ShapeCheckpoint:
    def __init__(self, cfg):
        from leopard_b200.weights import tensor_specs
        self.specs = {n: s for n, s, _ in tensor_specs(cfg)}
        self.cache = {}

    def f32(self, name):
        s = self.specs[name]
        if s not in self.cache:
            g = torch.Generator().manual_seed(len(self.cache))
            self.cache[s] = torch.randn(s, generator=g) * 0.02
        return self.cache[s]


class CpuSample:
    def __init__(self, cfg, n_crops_total: int, S_total: int, sample_crops: int = 1, threads: int | None = None):
        from leopard_b200.weights import synth_inputs
        self.cfg = cfg
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.ck = ShapeCheckpoint(cfg)
        text_total = S_total - n_crops_total * cfg.tokens_per_crop
        n_text = max(2, round(text_total * sample_crops / n_crops_total))
        self.pix, self.ids = synth_inputs(cfg, n_crops=sample_crops, n_text=n_text, seed=5)
        self.rows = sample_crops * cfg.tokens_per_crop + n_text
        self.desc = (f"full depth ({cfg.vit.layers} ViT + {cfg.dec.layers} decoder layers, real widths): {sample_crops} of "
                     f"{n_crops_total} crops -> {self.rows} of {S_total} merged rows per step, last-row logits; oracle port "
                     f"(torch-cpu fp32), {self.threads} threads; every step is a complete forward of this slice")

    def step(self) -> float:
        from oracle import model as om
        t0 = time.perf_counter()
        out = om.prefill_logits(self.ck, self.cfg, self.ids, self.pix, last_only=True)
        assert out.shape[1] == 1
        return time.perf_counter() - t0
