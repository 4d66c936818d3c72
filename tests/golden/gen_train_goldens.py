"""Golden training samples produced by EXECUTING the reference's `preprocess_llama3`
(Pai-Megatron-Patch/megatron_patch/data/llava/mm_pretrain_dataset.py:769-874, extracted with `ast`, run unmodified) with
the deterministic toy tokenizer of toy_tokenizer.py.  Authoring container only.  Writes tests/golden/ref_train_samples.json."""
import ast
import copy
import json
import sys
import types
from pathlib import Path
from typing import Dict

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from toy_tokenizer import ToyTokenizer  # noqa: E402

SRC = Path("/root/reference/Pai-Megatron-Patch/megatron_patch/data/llava/mm_pretrain_dataset.py")

CASES = [  # (sources, num_image, per_image_token, num_patches_per_image)
    ([[{"role": "user", "content": "<image>What is on slide 2?"}, {"role": "assistant", "content": "A chart."}]], 1, 3, [2]),
    ([[{"role": "user", "content": "Compare them."}, {"role": "assistant", "content": "Same."}]], 2, 2, [1, 0]),
    ([[{"role": "user", "content": "<image> and <image>"}, {"role": "assistant", "content": "ok"},
       {"role": "user", "content": "more?"}, {"role": "assistant", "content": "no"}]], 2, 2, []),
    ([[{"role": "system", "content": "be brief"}, {"role": "user", "content": "<image>hi"},
       {"role": "assistant", "content": "hello"}]], 1, 4, [0]),
    ([[{"role": "user", "content": "<image><image><image>too many"}, {"role": "assistant", "content": "x"}]], 2, 2, [1, 1]),
    ([[{"from": "gpt", "value": "skipped"}, {"from": "human", "value": "<image>q"}, {"from": "gpt", "value": "a"}]], 1, 2, [3]),
    # quirk: a missing tag is prepended through ['content'] — a from/value sample raises KeyError there
    ([[{"from": "human", "value": "no tag"}, {"from": "gpt", "value": "a"}]], 1, 2, [0]),
]


def main():
    tree = ast.parse(SRC.read_text())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "preprocess_llama3"]
    mod = types.ModuleType("ref_train")
    mod.__dict__.update(torch=torch, Dict=Dict, transformers=types.SimpleNamespace(PreTrainedTokenizer=object),
                        DEFAULT_IMAGE_TOKEN="<image>", IMAGE_TOKEN_INDEX=-200)
    exec(compile(ast.Module(body=fn, type_ignores=[]), str(SRC), "exec"), mod.__dict__)
    tok = ToyTokenizer()
    out = []
    for sources, n_img, per, tiles in CASES:
        rec = {"sources": sources, "num_image": n_img, "per_image_token": per, "num_patches_per_image": tiles}
        try:
            r = mod.preprocess_llama3(copy.deepcopy(sources), tok, num_image=n_img, per_image_token=per,
                                      num_patches_per_image=list(tiles))
            rec["result"] = None if r is None else {k: v.tolist() for k, v in r.items()}
        except Exception as e:  # noqa: BLE001
            rec["raises"] = type(e).__name__
        out.append(rec)
        print(rec.get("raises") or (rec["result"] and len(rec["result"]["input_ids"][0])))
    (HERE / "ref_train_samples.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
