"""Decode step with the attention kernel's L2 prefetch of the next projections' weights (lp_decode_set_prefetch): ms per generated
token on the cfg3 request for several (o_proj MiB, gate/up MiB) settings, default launch mode (GEMV kernels + CUDA graph + PDL).
Also checks that the generated tokens do not depend on the setting."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
from leopard_b200 import _lib  # noqa: E402
from leopard_b200.config import leopard_llava  # noqa: E402
from leopard_b200.hf import myLlavaForConditionalGeneration  # noqa: E402
from leopard_b200.native_weights import random_native  # noqa: E402

cfg = leopard_llava()
model = myLlavaForConditionalGeneration.from_native(cfg, random_native(cfg, "cuda:0", seed=1234))
pix, ids, n_crops, S, _ = bench.make_workload(cfg)
L = _lib.lib()
N_NEW = 65


def gen(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = model.generate(ids, pixel_values=pix, max_new_tokens=n)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, o


ref = None
settings = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(0, 0), (32, 0), (32, 48), (32, 96), (0, 48), (32, 24)]
for rnd in range(2):
    for o_mb, gu_mb in settings:
        L.lp_decode_set_prefetch(o_mb, gu_mb)
        st = model.engine._decode_state
        if st is not None:
            st.drop_graph()                       # the prefetch ranges are kernel arguments baked into the captured graph
        gen(8)
        t1 = min(gen(1)[0] for _ in range(2))
        tn, out = min((gen(N_NEW) for _ in range(2)), key=lambda r: r[0])
        if ref is None:
            ref = out
        print(json.dumps({"o_proj_mb": o_mb, "gateup_mb": gu_mb, "round": rnd, "ms_per_token": round((tn - t1) * 1e3 / (N_NEW - 1), 4),
                          "same_tokens": bool(torch.equal(out, ref))}), flush=True)
L.lp_decode_set_prefetch(32, 24)
