"""Time the SFT step (leopard_b200/train.py) on one B200: cfg5 shape (tokens [1, 4096], N crops), synthetic weights.
    python tools/gpu_train_speed.py [--layers 32] [--crops 20] [--seq 4096] [--steps 3] [--freeze-llm]
Prints one JSON line: ms per step, tokens/s, peak memory, launches per step, and the share of the step spent in the
materialised attention backward (the part a flash-style kernel replaces)."""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leopard_b200 import _lib  # noqa: E402
from leopard_b200.config import leopard_llava, reduced  # noqa: E402
from leopard_b200.engine import PrefillEngine  # noqa: E402
from leopard_b200.native_weights import random_native  # noqa: E402
from leopard_b200.train import SFTTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--vit-layers", type=int, default=27)
    ap.add_argument("--crops", type=int, default=20)
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--freeze-llm", action="store_true")
    ap.add_argument("--recompute", default="full", choices=["full", "none"])
    a = ap.parse_args()
    cfg = leopard_llava()
    if a.layers != 32 or a.vit_layers != 27:
        cfg = reduced(cfg, vit_layers=a.vit_layers, dec_layers=a.layers)
    dev = "cuda:0"
    eng = PrefillEngine(cfg, random_native(cfg, dev))
    tr = SFTTrainer(eng, lr=1e-5, freeze_llm=a.freeze_llm, recompute=a.recompute)
    rows = a.crops * cfg.tokens_per_crop
    assert rows < a.seq
    g = torch.Generator().manual_seed(0)
    text = torch.randint(0, 128000, (a.seq - rows,), generator=g)
    tokens = torch.cat([text[:8], torch.full((rows,), cfg.megatron_image_index), text[8:]])[None]
    labels = tokens.clone()
    labels[tokens < 0] = -100
    weights = torch.zeros(tokens.shape)
    weights[0, -(a.seq - rows) // 2:] = 1.0
    images = (torch.rand((a.crops, 3, 364, 364), generator=g) * 2 - 1).to(dev, torch.bfloat16)
    L = _lib.lib()
    out = []
    for it in range(a.steps + 1):
        torch.cuda.synchronize()
        n0, t0 = L.lp_launch_count(), time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss, norm = tr.step(tokens, labels, weights, images)
        e1.record()
        torch.cuda.synchronize()
        out.append(dict(ms=e0.elapsed_time(e1), wall_ms=(time.perf_counter() - t0) * 1e3, loss=loss, grad_norm=norm,
                        launches=L.lp_launch_count() - n0))
    timed = out[1:] or out          # --steps 0: a single (cold) step, for profiler runs
    ms = sum(o["ms"] for o in timed) / len(timed)
    print(json.dumps(dict(metric="sft_tokens_per_s", value=a.seq / ms * 1e3, ms_per_step=ms, seq=a.seq, crops=a.crops,
                          layers=a.layers, freeze_llm=a.freeze_llm, recompute=a.recompute, launches_per_step=timed[-1]["launches"],
                          peak_gb=torch.cuda.max_memory_allocated() / 2 ** 30, steps=out)))


if __name__ == "__main__":
    main()
