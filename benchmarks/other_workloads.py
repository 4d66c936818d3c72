"""bench.py --workload cfg4 | cfg5: the two BASELINE.json configurations that are not the headline metric.

cfg4  Leopard-Idefics2, 8 images 980 x 980 -> 8 x 4900 patches -> perceiver -> 512 image rows + 64 text ids (S = 576):
      prefill tokens/s, one request per GPU per step (replicas: the path has no exchange step).
cfg5  Leopard-LLaVA SFT step: tokens [1, 4096] = 21 crops x 169 image slots + 547 text ids, loss on the
      last 256 positions, bf16 compute / fp32 main params + Adam, full recompute (the reference's setting), vision tower
      trainable unless --sft-freeze-vit (the launch script passes no freeze flag).  N > 1 = data parallel over NCCL: one
      micro-batch per rank, bucketed gradient all-reduce overlapped with the backward pass
      (megatron/core/distributed/param_and_grad_buffer.py:127-141).
Both print one JSON line in the bench contract's shape (metric / value / unit / config / e2e / roofline ...); the CPU
reference arm is defined for the default workload only.
"""
from __future__ import annotations

import json
import os

import torch


def _dist():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: leopard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return dist, world, rank, local, dev


def _timed(dist, world, dev, fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps


def _peaks():
    import bench
    return bench.load_peaks()


# ------------------------------------------------------------------------------------------------ cfg4: Idefics2 prefill
def run_idefics2(args):
    import bench
    from leopard_b200 import _lib
    from leopard_b200.config import leopard_idefics2
    from leopard_b200.idefics2 import Idefics2ForConditionalGeneration
    from leopard_b200.weights import idefics2_tensor_specs
    dist, world, rank, local, dev = _dist()
    cfg = leopard_idefics2()
    g = torch.Generator(device=dev).manual_seed(3)
    specs = {n: (s, k) for n, s, k in idefics2_tensor_specs(cfg)}

    def get(name):      # device-side random init with the seeded checkpoint's distributions
        shape, kind = specs[name]
        if kind in ("linear", "conv"):
            fan = 1
            for s in shape[1:]:
                fan *= s
            return (torch.randn(shape, device=dev, generator=g) * fan ** -0.5).to(torch.bfloat16)
        if kind in ("norm", "latents"):
            return (1.0 + torch.randn(shape, device=dev, generator=g) * 0.02).to(torch.bfloat16)
        return (torch.randn(shape, device=dev, generator=g) * (0.05 if kind == "embed" else 0.02)).to(torch.bfloat16)

    model = Idefics2ForConditionalGeneration(cfg, get, dev)
    eng = model.engine
    gc = torch.Generator().manual_seed(1)
    n_img, side = 8, 980
    pv_h = (torch.rand(1, n_img, 3, side, side, generator=gc) * 2 - 1).to(torch.bfloat16).pin_memory()
    text = torch.randint(0, 32000, (64,), generator=gc)
    ids_h = torch.cat([text[:8]] + [torch.full((64,), cfg.image_token_id)] * n_img + [text[8:]])[None].pin_memory()
    pv_d, ids_d = pv_h.to(dev), ids_h.to(dev)
    S = ids_h.shape[1]
    L = _lib.lib()

    def step():
        return eng.prefill(ids_d, pv_d, None, "last")[0]

    def step_e2e():
        return model.generate(input_ids=ids_h, pixel_values=pv_h.to(dev, non_blocking=True), max_new_tokens=1).cpu()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.lp_launch_count()
    ms = _timed(dist, world, dev, step, args.steps)
    launches = L.lp_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = _timed(dist, world, dev, step_e2e, args.steps)
    T = (side // 14) ** 2
    vit = n_img * (27 * (30_449_664 * T + 4608 * T * T) + 1_354_752 * T)
    conn = n_img * (0.899e12 + 0.182e12)
    dec = S * 2 * 32 * 218_103_808 + 262_144 * S * S + 2 * 4096 * cfg.dec.vocab
    flops = vit + conn + dec
    pk = _peaks()
    if rank == 0:
        ach = flops / (ms * 1e-3) / 1e12
        print(json.dumps({
            "metric": "prefill tokens/sec (Leopard-Idefics2 8-image prefill, whole job)", "value": world * S / (ms * 1e-3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg4: Idefics2 8 x (980 x 980) -> 8 x 4900 patches (NaViT SigLIP tower) -> modality projection "
                                   "+ perceiver resampler -> 512 image rows + 64 text ids = S 576 -> Mistral-7B decoder, 1 request "
                                   "per GPU per step, synthetic weights", "S": S,
                       "l2_policy": "inputs larger than L2 (16 GB of weights per step)", "parallelism": f"dp{world}: replicas"},
            "e2e": {"value": world * S / (ms_e2e * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": pv_h.numel() * 2 + ids_h.numel() * 8, "d2h_bytes_per_step": (S + 1) * 8,
                    "api": "leopard_b200.idefics2.Idefics2ForConditionalGeneration.generate(host ids, pixel_values=pinned host bf16)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "whole step (ViT at T = 4900 is 77 % of the FLOPs, 43 % of it attention)",
                         "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                         "traffic": None, "algorithmic_tflop_per_step": flops / 1e12},
            "cpu_baseline": None}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------ cfg5: SFT step
def run_sft(args):
    import bench
    from leopard_b200 import _lib
    from leopard_b200.config import leopard_llava
    from leopard_b200.engine import PrefillEngine
    from leopard_b200.native_weights import random_native
    from leopard_b200.train import SFTTrainer
    dist, world, rank, local, dev = _dist()
    cfg = leopard_llava()
    eng = PrefillEngine(cfg, random_native(cfg, dev, seed=1234))
    tr = SFTTrainer(eng, lr=1e-5, freeze_vision_tower=bool(args.sft_freeze_vit), recompute=args.sft_recompute)
    seq, crops = 4096, 21
    rows = crops * cfg.tokens_per_crop
    g = torch.Generator().manual_seed(rank)              # every rank draws its own micro-batch
    text = torch.randint(0, 128000, (seq - rows,), generator=g)
    tokens = torch.cat([text[:8], torch.full((rows,), cfg.megatron_image_index), text[8:]])[None]
    labels = tokens.clone()
    labels[tokens < 0] = -100
    weights = torch.zeros(tokens.shape)
    weights[0, -256:] = 1.0
    images_h = (torch.rand((crops, 3, 364, 364), generator=g) * 2 - 1).to(torch.bfloat16).pin_memory()
    images_d = images_h.to(dev)
    L = _lib.lib()
    out = {}

    def step():
        out["loss"], out["norm"] = tr.step(tokens, labels, weights, images_d)

    def step_e2e():      # host batch (what mm_collate_fn hands over): pinned crops -> device inside the timed region
        out["loss"], out["norm"] = tr.step(tokens, labels, weights, images_h.to(dev, non_blocking=True))

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.lp_launch_count()
    ms = _timed(dist, world, dev, step, args.steps)
    launches = (L.lp_launch_count() - n0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = _timed(dist, world, dev, step_e2e, args.steps)
    comm = None
    if world > 1:       # the same step without the gradient reduce: the difference is the EXPOSED all-reduce time
        tr.world = 1
        ms_local = _timed(dist, world, dev, step, max(2, args.steps // 2))
        tr.world = world
        nbytes = tr.flat_grad.numel() * 4
        comm = {"allreduce_bytes_per_step": nbytes, "buckets": len(tr.buckets), "exposed_ms": round(ms - ms_local, 3),
                "step_without_reduce_ms": round(ms_local, 3),
                "what": "fp32 flat gradient buffer, one NCCL all-reduce per bucket (lm_head+final norm, each decoder layer, "
                        "embedding, projector, tower) issued when the bucket's last kernel is enqueued"}
    T = cfg.vit.tokens
    f_vit = crops * (27 * (30_449_664 * T + 4608 * T * T) + 1_354_752 * T)
    f_proj = crops * 2 * 169 * (4608 * 4096 + 4096 * 4096)
    f_dec = seq * 2 * 32 * 218_103_808 + 262_144 * seq * seq + seq * 2 * 4096 * cfg.dec.vocab
    fwd = f_vit + f_proj + f_dec
    model_flops = 3 * (f_proj + f_dec) + (1 if args.sft_freeze_vit else 3) * f_vit       # fwd + 2x bwd, no recompute counted
    pk = _peaks()
    if rank == 0:
        ach = model_flops / (ms * 1e-3) / 1e12
        print(json.dumps({
            "metric": "SFT tokens/sec (Leopard-LLaVA SFT step, seq 4096, whole job)", "value": world * seq / (ms * 1e-3),
            "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"cfg5: Leopard-LLaVA SFT step, tokens [1, 4096] = 21 crops x 169 image slots + 547 text ids, "
                                   f"loss on the last 256 positions, micro-batch 1 per GPU, Adam fp32 main params, recompute "
                                   f"{args.sft_recompute}, vision tower {'frozen' if args.sft_freeze_vit else 'trainable'}, "
                                   "synthetic weights", "S": seq,
                       "l2_policy": "inputs larger than L2 (weights + optimizer state = 144 GB touched per step)",
                       "parallelism": f"dp{world}: bucketed NCCL gradient all-reduce" if world > 1 else "dp1"},
            "e2e": {"value": world * seq / (ms_e2e * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": images_h.numel() * 2 + 3 * seq * 8, "d2h_bytes_per_step": 8,
                    "api": "leopard_b200.train.SFTTrainer.step(host tokens / labels / loss mask, pinned host crops)"},
            "gpu_launches": int(launches), "clocks": clocks, "loss": out["loss"], "grad_norm": out["norm"],
            "peak_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
            "roofline": {"bound": "tensor", "kernel": "whole step: forward + backward (dgrad + wgrad) + optimizer",
                         "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                         "traffic": None, "model_tflop_per_step": model_flops / 1e12, "forward_tflop": fwd / 1e12,
                         "note": "model FLOPs = 3 x forward of the trainable parts (+1 x the tower when frozen); the recomputed "
                                 "forward of --recompute full is work done but not counted"},
            "comm": comm, "cpu_baseline": None}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run(args):
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm is defined for the default workload "
                                                                  "(cfg3) and cfg2 only"}), flush=True)
        return 0
    return run_idefics2(args) if args.workload == "cfg4" else run_sft(args)
