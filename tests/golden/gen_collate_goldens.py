"""Golden collated batches produced by EXECUTING the reference's `mm_collate_fn`
(Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py:163-191, extracted with `ast`, run unmodified).
Authoring container only.  Writes tests/golden/ref_collate.json; inputs are regenerated from `make_batch(case)`."""
import ast
import contextlib
import io
import json
import types
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference/Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py")

CASES = [  # per sample: (length, crops or None) ; None = a sample the dataset failed to load
    [(13, 2), (8, 1)],
    [(16, None), None, (5, 3)],
    [(7, None), (9, None)],
    [None, None],
    [(24, 1)],
]


def make_batch(case, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for spec in case:
        if spec is None:
            out.append(None)
            continue
        n, crops = spec
        item = {"input_ids": torch.randint(-200, 1000, (n,), generator=g), "labels": torch.randint(0, 1000, (n,), generator=g),
                "weights": torch.rand(n, generator=g), "preferences": torch.rand(n, generator=g)}
        if crops is not None:
            item["image"] = torch.randn(crops, 3, 4, 4, generator=g).to(torch.bfloat16)
        out.append(item)
    return out


def main():
    tree = ast.parse(SRC.read_text())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "mm_collate_fn"]
    mod = types.ModuleType("ref_collate")
    mod.__dict__.update(torch=torch, IGNORE_INDEX=-100)
    exec(compile(ast.Module(body=fn, type_ignores=[]), str(SRC), "exec"), mod.__dict__)
    recs = []
    for k, case in enumerate(CASES):
        with contextlib.redirect_stdout(io.StringIO()):
            r = mod.mm_collate_fn(make_batch(case, seed=k))
        recs.append(None if r is None else {key: {"shape": list(v.shape), "dtype": str(v.dtype), "data": v.float().flatten().tolist()}
                                            for key, v in r.items()})
        print(k, None if r is None else {key: tuple(v.shape) for key, v in r.items()})
    (HERE / "ref_collate.json").write_text(json.dumps(recs))


if __name__ == "__main__":
    main()
