"""Full-depth oracle fixtures for BASELINE configs 2 and 3 (27 ViT + 32 decoder layers, real widths).

Run in the authoring container (CPU only, ~7 + ~13 minutes on 8 cores, layer-streamed weights):
    python tests/golden/gen_fullsize_goldens.py            # writes tests/golden/fullsize_goldens.npz

What is pinned: the CPU oracle (oracle/model.py — itself pinned to reference-executed goldens and to the installed
transformers modules, see its header) on the seeded synthetic checkpoint (seed 1234) and seeded inputs:
  cfg2  1 x 1344x1344  -> 17 crops, S = 2937        cfg3  4 x 896x1152 -> 28 crops, S = 4796
Stored per config (fp32 unless noted; the full [S, V] logits would be 1.5-2.5 GB):
  last_logits [V]                 the row greedy decoding reads
  rows [R], cols [C]              every 64th row (+ the last), every 97th vocabulary column
  sub_logits [R, C]               logits at rows x cols
  row_argmax [R] (int64), row_top5_idx/val [R, 5], row_absmax [R]
  feat_sub [N, 13, 64]            image_features[:, ::13, ::64] (tower + shuffle + projector)
  hidden_sub [R, 64]              final-norm hidden state at rows, every 64th channel
tests/test_gpu_fullsize.py compares the CUDA path with these on the GPU box (where /root/reference and the minutes of
CPU time are not available)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from leopard_b200.config import leopard_llava  # noqa: E402
from leopard_b200.preprocess import plan_crops  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint, synth_inputs  # noqa: E402
from oracle import model as om  # noqa: E402

CONFIGS = {
    "cfg2": dict(image_sizes=[(1344, 1344)], n_text=64, seed=21),
    "cfg3": dict(image_sizes=[(896, 1152)] * 4, n_text=64, seed=31),
}
COL_STRIDE, ROW_STRIDE = 97, 64


def fixture_inputs(cfg, name):
    """The seeded request of a config — shared by this generator and the GPU test."""
    c = CONFIGS[name]
    n = sum(plan_crops(c["image_sizes"]))
    pix, ids = synth_inputs(cfg, n_crops=n, n_text=c["n_text"], seed=c["seed"])
    return pix.to(torch.bfloat16), ids


@torch.no_grad()
def run(cfg, ck, name):
    t0 = time.time()
    pix, ids = fixture_inputs(cfg, name)
    stages = {}
    vit_out = om.vit_forward(ck, cfg, pix.float())
    feats = om.projector(ck, cfg, vit_out)
    print(f"[{name}] tower+projector {time.time() - t0:.0f}s", flush=True)
    embeds, pos = om.merge_hf_expand(ck, cfg, ids, feats)
    hidden = om.decoder_forward(ck, cfg, embeds[0], pos[0], return_hidden=True)
    S = hidden.shape[0]
    print(f"[{name}] decoder S={S} {time.time() - t0:.0f}s", flush=True)
    rows = torch.unique(torch.cat([torch.arange(0, S, ROW_STRIDE), torch.tensor([S - 1])]))
    w = ck.f32("language_model.lm_head.weight")
    logits = F.linear(hidden[rows], w)                      # [R, V]
    cols = torch.arange(0, w.shape[0], COL_STRIDE)
    top = logits.topk(5, dim=-1)
    out = {
        "S": np.int64(S), "n_crops": np.int64(pix.shape[0]),
        "last_logits": logits[-1].numpy(), "rows": rows.numpy(), "cols": cols.numpy(),
        "sub_logits": logits[:, cols].numpy(), "row_argmax": logits.argmax(-1).numpy(),
        "row_top5_idx": top.indices.numpy(), "row_top5_val": top.values.numpy(),
        "row_absmax": logits.abs().amax(-1).numpy(),
        "feat_sub": feats[:, ::13, ::64].contiguous().numpy(), "feat_absmax": np.float32(feats.abs().max()),
        "hidden_sub": hidden[rows][:, ::64].contiguous().numpy(),
    }
    print(f"[{name}] done {time.time() - t0:.0f}s; |logit|max {float(logits.abs().max()):.3f}", flush=True)
    return {f"{name}.{k}": v for k, v in out.items()}


def main():
    torch.manual_seed(0)
    cfg = leopard_llava()
    ck = SyntheticCheckpoint(cfg, seed=1234)
    names = sys.argv[1:] or list(CONFIGS)
    dst = Path(__file__).with_name("fullsize_goldens.npz")
    blob = dict(np.load(dst)) if dst.exists() else {}
    for name in names:
        blob.update(run(cfg, ck, name))
        np.savez_compressed(dst, **blob)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
