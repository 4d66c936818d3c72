"""Phase timeline of the single-kernel decode step (mode 21): CTA 0's globaltimer stamps around every grid barrier."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ["LEOPARD_DECODE_GRAPH"] = "0"
os.environ["LEOPARD_DECODE_MODE"] = "21"
import torch  # noqa: E402

import bench  # noqa: E402
from leopard_b200 import _lib  # noqa: E402
from leopard_b200.config import leopard_llava  # noqa: E402
from leopard_b200.hf import myLlavaForConditionalGeneration  # noqa: E402
from leopard_b200.native_weights import random_native  # noqa: E402

cfg = leopard_llava()
model = myLlavaForConditionalGeneration.from_native(cfg, random_native(cfg, "cuda:0", seed=1234))
pix, ids, n_crops, S, _ = bench.make_workload(cfg)
model.generate(ids, pixel_values=pix, max_new_tokens=4)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
_lib.lib().lp_decode_debug_stamps(_lib.ptr(buf))
model.generate(ids, pixel_values=pix, max_new_tokens=4)   # last step's stamps survive
torch.cuda.synchronize()
_lib.lib().lp_decode_debug_stamps(None)
t = buf.cpu().tolist()
n_bar = 5 * cfg.dec.layers + 3
names = ["embed"] + ["qkv", "attn", "oproj", "gateup", "down"] * cfg.dec.layers + ["lm_head", "argmax"]
work, wait = {}, {}
prev_out = None
for e in range(1, n_bar):
    t_in, t_out = t[2 * e], t[2 * e + 1]
    nm = names[e - 1]
    if prev_out is not None:
        work.setdefault(nm, []).append(t_in - prev_out)
    wait.setdefault(nm, []).append(t_out - t_in)
    prev_out = t_out
print("total step (first barrier in -> last barrier out): %.1f us" % ((t[2 * (n_bar - 1) + 1] - t[2]) / 1e3))
for nm in ["qkv", "attn", "oproj", "gateup", "down", "lm_head"]:
    w = work.get(nm, [0]); b = wait.get(nm, [0])
    print(f"{nm:8s} CTA0 work avg {sum(w)/len(w)/1e3:7.2f} us   barrier wait avg {sum(b)/len(b)/1e3:6.2f} us   (n={len(w)})")
