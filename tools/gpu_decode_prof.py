"""Decode-step kernels in isolation for ncu (eager launches, GEMV mode): one cfg3 prefill, then a few greedy steps.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'gemv_kernel|decode_' -s 400 -c 200 --csv --log-file gpurun_out/decode_launches.csv \
        python tools/gpu_decode_prof.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["LEOPARD_DECODE_GRAPH"] = os.environ.get("LEOPARD_DECODE_GRAPH", "0")
import torch  # noqa: E402

import bench  # noqa: E402
from leopard_b200.config import leopard_llava  # noqa: E402
from leopard_b200.hf import myLlavaForConditionalGeneration  # noqa: E402
from leopard_b200.native_weights import random_native  # noqa: E402

cfg = leopard_llava()
model = myLlavaForConditionalGeneration.from_native(cfg, random_native(cfg, "cuda:0", seed=1234))
pix, ids, n_crops, S, _ = bench.make_workload(cfg)
out = model.generate(ids, pixel_values=pix, max_new_tokens=int(os.environ.get("N_NEW", "5")))
torch.cuda.synchronize()
print("generated", out.shape[1] - ids.shape[1])
