"""A/B of whole-step prefill time on the cfg3 workload for different kernel switches, interleaved in ONE process so that
box-to-box clock / power-cap differences cancel.  Usage: python tools/gpu_ab_prefill.py attn=2 attn=80 pair=0 ..."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
from leopard_b200 import _lib  # noqa: E402
from leopard_b200.config import leopard_llava  # noqa: E402
from leopard_b200.engine import PrefillEngine  # noqa: E402
from leopard_b200.native_weights import random_native  # noqa: E402

cfg = leopard_llava()
eng = PrefillEngine(cfg, random_native(cfg, "cuda:0", seed=1234))
pix, ids, n_crops, S, _ = bench.make_workload(cfg, os.environ.get("LEOPARD_AB_WORKLOAD", "cfg3"))   # cfg3 (default) or cfg2
pix, ids = pix.cuda(), ids[0].cuda()
L = _lib.lib()


def apply(setting):
    k, v = setting.split("=")
    v = int(v)
    if k == "attn":
        L.lp_model_set_attention_variant(v)
    elif k == "pair":
        L.lp_gemm_set_pair_mode(v)
    elif k == "pdl":
        L.lp_set_prefill_pdl(v)
    elif k == "norm":
        L.lp_norm_set_variant(v)
    elif k == "fused":
        L.lp_model_set_fused_norm(v)
    elif k == "chain":
        L.lp_model_set_chained_projector(v)
    elif k == "sk":
        L.lp_gemm_set_streamk(v)
    elif k == "episleep":
        L.lp_gemm_set_epilogue_sleep(v)
    elif k == "bal":
        L.lp_gemm_set_raster_balance(v)
    elif k == "tmares":
        L.lp_gemm_set_tma_residual(v)
    else:
        raise SystemExit(f"unknown switch {k}")


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        eng.prefill([(ids, pix)], "last")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


settings = sys.argv[1:] or ["attn=2", "attn=80"]
for _ in range(3):
    eng.prefill([(ids, pix)], "last")
res = {s: [] for s in settings}
for rnd in range(3):
    for s in settings:
        apply(s)
        timed(2)
        res[s].append(round(timed(6), 3))
for s in settings:
    print(json.dumps({"setting": s, "ms_per_step": res[s], "best": min(res[s]), "tokens_per_s": round(S / min(res[s]) * 1e3, 1)}))
L.lp_model_set_attention_variant(336)
L.lp_gemm_set_pair_mode(1)
L.lp_model_set_fused_norm(1)
L.lp_gemm_set_streamk(1)
L.lp_gemm_set_tma_residual(1)
