"""Greedy decode speed of the native generate() on the cfg3 request: ms per generated token for each
decode implementation, against the HBM floor (every decoder weight + the KV cache read once per token).

  mode 0 eager : first implementation, M = 1 rows through the tcgen05 GEMM, one launch per op
  mode 1 graph : warp-per-row-pair GEMV kernels, one CUDA-graph replay per token
  mode 5 eager : block-cooperative GEMV kernels, launched one by one
  mode 5 graph : block-cooperative GEMV kernels, CUDA-graph replay (default)
  mode 7 graph : + programmatic dependent launch edges
  mode 21      : one persistent cooperative kernel per token (eager / graph)
Writes one JSON line per variant to stdout (and to gpurun_out/decode_speed.json when that directory exists).
"""
import json
import os
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
d = cfg.dec
w_bytes = 2 * (d.layers * (d.hidden * (d.heads + 2 * d.kv_heads) * d.head_dim + d.hidden * d.hidden +
                           3 * d.hidden * d.ffn) + model.engine.w.vocab_padded * d.hidden)
kv_bytes = 2 * d.layers * 2 * S * d.kv_heads * d.head_dim
peaks = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())
hbm = peaks.get("hbm_gbs", 6500.0)
N_NEW = 65
out = []
for mode, graph in ((0, "0"), (1, "1"), (5, "0"), (5, "1"), (7, "1"), (21, "0"), (21, "1")):
    os.environ["LEOPARD_DECODE_MODE"] = str(mode)
    os.environ["LEOPARD_DECODE_GRAPH"] = graph if mode else "0"
    if mode == 0:
        # the first implementation has no device-resident loop: time lp_decode_step directly
        import ctypes as C
        eng = model.engine
        L = _lib.lib()
        L.lp_decode_set_mode(0)
        logits, kv, cu = eng.prefill([(ids[0].cuda(), pix.cuda())], "last", want_kv=True, kv_capacity=N_NEW)
        ws = eng._workspace("decode", L.lp_decode_workspace_bytes(C.byref(eng.w.dec_dims)))
        lg = torch.empty((eng.w.vocab_padded,), dtype=torch.float32, device="cuda")
        tok = logits.argmax(-1).to(torch.long)

        def steps(n):
            global tok
            for i in range(n):
                rc = L.lp_decode_step(C.byref(eng.w.dec_dims), C.byref(eng.w.dec), _lib.ptr(eng.w.embed),
                                      eng.w.embed.stride(0), _lib.ptr(tok), cu[-1] + i, _lib.ptr(kv), kv.shape[2],
                                      _lib.ptr(lg), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
                _lib.check(rc, "lp_decode_step")
                tok = lg[: d.vocab].argmax(-1, keepdim=True).to(torch.long)
                int(tok.item())
        steps(4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps(N_NEW - 1)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / (N_NEW - 1)
        L.lp_decode_set_mode(7)
        del kv
    else:
        def gen(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o = model.generate(ids, pixel_values=pix, max_new_tokens=n)
            torch.cuda.synchronize()
            assert o.shape[1] == ids.shape[1] + n
            return time.perf_counter() - t0
        gen(N_NEW)  # warm-up (graph capture, buffers)
        t1 = min(gen(1) for _ in range(2))
        tn = min(gen(N_NEW) for _ in range(2))
        ms = (tn - t1) * 1e3 / (N_NEW - 1)
    rec = {"mode": mode, "graph": graph == "1" and mode != 0, "S": S, "ms_per_token": round(ms, 4),
           "weights_GB": round(w_bytes / 1e9, 3), "kv_GB": round(kv_bytes / 1e9, 3),
           "achieved_GBs": round((w_bytes + kv_bytes) / ms / 1e6, 1), "hbm_peak_GBs": hbm,
           "frac_of_hbm_peak": round((w_bytes + kv_bytes) / ms / 1e6 / hbm, 3)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
_lib.lib().lp_decode_set_mode(7)
od = Path(__file__).resolve().parents[1] / "gpurun_out"
if od.exists():
    (od / "decode_speed.json").write_text("\n".join(json.dumps(r) for r in out) + "\n")
