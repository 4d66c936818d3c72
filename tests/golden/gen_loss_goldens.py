"""Golden masked-LM losses produced by EXECUTING the reference's `loss_func`
(Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py:101-121, extracted with `ast`, run unmodified) with
single-rank stand-ins for the two Megatron helpers it calls:
  tensor_parallel.vocab_parallel_cross_entropy -> fp32 logsumexp - logit[label]; labels outside [0, V) contribute a
      predicted logit of 0, which is what Megatron's vocab-parallel kernel does with its target mask at TP = 1
  average_losses_across_data_parallel_group    -> identity (one data-parallel rank)
Authoring container only.  Writes tests/golden/ref_loss.json (inputs are regenerated from the seeds in the file)."""
import ast
import json
import types
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference/Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py")

CASES = [  # (seed, batch, seq, vocab, fraction of positions that carry weight)
    (1, 1, 9, 17, 0.5),
    (2, 2, 33, 101, 0.3),
    (3, 3, 16, 64, 1.0),
    (4, 1, 128, 257, 0.05),
]


def make_case(seed, b, s, v, frac):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn((b, s, v), generator=g) * 3).to(torch.bfloat16)
    labels = torch.randint(0, v, (b, s), generator=g)
    weights = (torch.rand((b, s), generator=g) < frac).float()
    weights[:, -1] = 1.0                                  # at least one weighted target after the shift
    labels = torch.where(weights > 0, labels, torch.full_like(labels, -100))   # IGNORE_INDEX where unweighted
    labels[0, 0] = -200                                   # an image slot id in the label stream (never weighted)
    weights[0, 0] = 0.0
    return logits, labels, weights


def vocab_parallel_cross_entropy(logits, target):
    v = logits.shape[-1]
    outside = (target < 0) | (target >= v)
    t = target.masked_fill(outside, 0)
    picked = logits.gather(-1, t.unsqueeze(-1)).squeeze(-1).masked_fill(outside, 0.0)
    return torch.logsumexp(logits, -1) - picked


def main():
    tree = ast.parse(SRC.read_text())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "loss_func"]
    mod = types.ModuleType("ref_loss")
    mod.__dict__.update(torch=torch,
                        tensor_parallel=types.SimpleNamespace(vocab_parallel_cross_entropy=vocab_parallel_cross_entropy),
                        average_losses_across_data_parallel_group=lambda xs: torch.stack([x.detach() for x in xs]))
    exec(compile(ast.Module(body=fn, type_ignores=[]), str(SRC), "exec"), mod.__dict__)
    out = []
    for seed, b, s, v, frac in CASES:
        logits, labels, weights = make_case(seed, b, s, v, frac)
        loss, logged = mod.loss_func(weights, labels, logits)
        out.append({"seed": seed, "batch": b, "seq": s, "vocab": v, "frac": frac, "loss": float(loss),
                    "logged": float(logged["lm loss"])})
        print(out[-1])
    (HERE / "ref_loss.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
