"""Greedy decode speed of the native generate() on the cfg3 request: prefill + n new tokens."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
from leopard_b200.config import leopard_llava  # noqa: E402
from leopard_b200.hf import myLlavaForConditionalGeneration  # noqa: E402
from leopard_b200.native_weights import random_native  # noqa: E402

cfg = leopard_llava()
model = myLlavaForConditionalGeneration.from_native(cfg, random_native(cfg, "cuda:0", seed=1234))
pix, ids, n_crops, S, _ = bench.make_workload(cfg)
for n_new in (1, 33):
    model.generate(ids, pixel_values=pix, max_new_tokens=n_new)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.generate(ids, pixel_values=pix, max_new_tokens=n_new)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"new_tokens": int(out.shape[1] - ids.shape[1]), "seconds": dt}))
