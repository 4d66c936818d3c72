"""The reference's training loop body (Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py + Megatron's `train`) with
leopard_b200 under it — what a maintainer's script looks like after the switch described in INTEGRATION.md §2.
NOT YET RUN ON A B200: the step's kernels are verified under CPU emulation only.

Reference                                   here
  :37-50   model_provider                    leopard_b200.megatron.model_provider (+ configure_provider for args / weights)
  :163-191 mm_collate_fn                     leopard_b200.megatron.mm_collate_fn (same dict)
  dataset  mm_pretrain_dataset.py            synthetic samples in the dataset's output format: input_ids with one -200 per image
                                             feature row, labels, weights (loss mask), image crops [n, 3, 364, 364] bf16
  training.py:544-607 train_step             leopard_b200.megatron.train_step over leopard_b200.train.SFTTrainer
Run:  python examples/llava_sft_loop.py --layers 2 --iters 5            (needs a B200; no CPU fallback)
      torchrun --nproc-per-node 2 examples/llava_sft_loop.py ...        (data parallel: gradients averaged per bucket)
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from leopard_b200 import megatron  # noqa: E402
from leopard_b200.config import leopard_llava, reduced  # noqa: E402
from leopard_b200.train import SFTTrainer  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint  # noqa: E402


def synthetic_sample(cfg, n_crops: int, n_text: int, seed: int):
    """One dataset item as mm_pretrain_dataset.py yields it: the image block first, then prompt + answer; only the answer
    tokens carry loss weight."""
    g = torch.Generator().manual_seed(seed)
    rows = n_crops * cfg.tokens_per_crop
    text = torch.randint(0, min(cfg.dec.vocab, 128000), (n_text,), generator=g)
    ids = torch.cat([text[:4], torch.full((rows,), cfg.megatron_image_index), text[4:]])
    labels = ids.clone()
    labels[ids < 0] = megatron.IGNORE_INDEX
    weights = torch.zeros(ids.shape)
    weights[-(n_text // 2):] = 1.0
    image = (torch.rand((n_crops, 3, 364, 364), generator=g) * 2 - 1).to(torch.bfloat16)
    return {"input_ids": ids, "labels": labels, "weights": weights, "preferences": torch.zeros(ids.shape), "image": image}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=2, help="reduced ViT / decoder depth (0 = the full 27 + 32 layers)")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--crops", type=int, default=4)
    ap.add_argument("--text", type=int, default=64)
    ap.add_argument("--micro-batches", type=int, default=1)
    ap.add_argument("--lr", type=float, default=1e-5)
    ap.add_argument("--train-vision-tower", action="store_true")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")

    cfg = leopard_llava()
    if args.layers:
        cfg = reduced(cfg, vit_layers=args.layers, dec_layers=args.layers)
    ck = SyntheticCheckpoint(cfg, seed=1234)                       # stands in for Megatron's checkpoint loading
    megatron.configure_provider(cfg, hf_getter=lambda n: ck[n], device=f"cuda:{local}")
    model = [megatron.model_provider()]
    optimizer = SFTTrainer(model[0].engine, lr=args.lr, min_lr=args.lr / 10, warmup_iters=1, decay_iters=max(args.iters, 2),
                           clip_grad=1.0, freeze_vision_tower=not args.train_vision_tower)

    def batches():
        k = 0
        while True:                                               # every rank draws its own samples (data parallel)
            yield megatron.mm_collate_fn([synthetic_sample(cfg, args.crops, args.text, seed=1000 * rank + k)])
            k += 1

    it = batches()
    for i in range(args.iters):
        loss, skipped, grad_norm, _ = megatron.train_step(None, it, model, optimizer, None, None,
                                                          num_microbatches=args.micro_batches)
        if rank == 0:
            print(f"iteration {i + 1}: lm loss {loss['lm loss']:.4f}  grad norm {grad_norm:.3f}", flush=True)
    state = {"model": model[0].state_dict_for_save_checkpoint(), "optimizer": optimizer.state_dict()}
    if rank == 0:
        print("checkpoint entries:", sorted(state["model"]["language_model"]), "| optimizer step", state["optimizer"]["iteration"])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
