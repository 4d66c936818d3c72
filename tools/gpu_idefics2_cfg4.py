"""BASELINE.json configs[3] (Leopard-Idefics2, 8 images 980x980 -> 8 x 4900 patches -> 512 image rows + 64 text ids,
S = 576) on the native path with device-side random weights: timing of one request per step (parity of this path is
tests/test_gpu_idefics2.py; this is a measurement aid, not the bench line)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from leopard_b200.config import leopard_idefics2  # noqa: E402
from leopard_b200.idefics2 import Idefics2Engine, Idefics2Weights  # noqa: E402
from leopard_b200.weights import idefics2_tensor_specs  # noqa: E402

cfg = leopard_idefics2()
g = torch.Generator(device="cuda").manual_seed(3)
specs = {n: (s, k) for n, s, k in idefics2_tensor_specs(cfg)}


def get(name):
    shape, kind = specs[name]
    if kind in ("linear", "conv"):
        fan = 1
        for s in shape[1:]:
            fan *= s
        return (torch.randn(shape, device="cuda", generator=g) * fan ** -0.5).to(torch.bfloat16)
    if kind in ("norm", "latents"):
        return (1.0 + torch.randn(shape, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    return (torch.randn(shape, device="cuda", generator=g) * (0.05 if kind == "embed" else 0.02)).to(torch.bfloat16)


eng = Idefics2Engine(cfg, Idefics2Weights(cfg, get, "cuda"))
gc = torch.Generator().manual_seed(1)
pv = (torch.rand(1, 8, 3, 980, 980, generator=gc) * 2 - 1).to(torch.bfloat16).cuda()
text = torch.randint(0, 32000, (64,), generator=gc)
ids = torch.cat([text[:8]] + [torch.full((64,), cfg.image_token_id)] * 8 + [text[8:]])[None]
for _ in range(2):
    eng.prefill(ids, pv, None, "last")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    logits, _, S = eng.prefill(ids, pv, None, "last")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flops = 8 * (7.022e12 + 0.899e12 + 0.182e12) + 8.13e12
print(json.dumps({"config": "cfg4 idefics2 8x980x980", "S": S, "ms_per_request": ms, "algorithmic_tflop": flops / 1e12,
                  "tflops": flops / ms / 1e9, "finite": bool(torch.isfinite(logits).all())}))
