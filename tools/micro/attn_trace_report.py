"""Summarise a phase trace written by tools/micro/attn_trace (see its header): mean cycles of every stamped interval."""
import json
import sys
from collections import defaultdict

SM = {(1, 2): "wait S (s_full)", (2, 3): "tcgen05.ld S + wait", (3, 4): "mask / row max / rescale decision", (3, 31): "  mask", (31, 32): "  row max (33 x 3-input max) + scale", (32, 4): "  rescale decision (vote) / rescale", (4, 5): "turn barrier (wait for the other tile's exp)",
      (5, 6): "exp section (fma2 / ex2 / add2 / pack / tcgen05.st)", (6, 7): "wait::st + fence", (7, 8): "wait P.V(t-1) + arrive p_full", (8, 1): "loop overhead"}
for path in sys.argv[1:]:
    d = json.load(open(path))
    print(f"== {path}: T={d['T']} dh={d['head_dim']} causal={d['causal']} pingpong={d['pingpong']} block={d['block']} {d['ms_untraced']:.4f} ms")
    for who in ("softmax_A", "softmax_B"):
        ev = d["trace"][who]
        if not ev:
            continue
        acc, cnt = defaultdict(float), defaultdict(int)
        starts = [c for i, c in ev if i == 1]
        for (i0, c0), (i1, c1) in zip(ev, ev[1:]):
            acc[(i0, i1)] += c1 - c0
            cnt[(i0, i1)] += 1
        n = len(starts)
        span = (ev[-1][1] - ev[0][1])
        print(f" {who}: {n} half steps, {span} cycles from first stamp to o_done, {span / max(n, 1):.0f} per half step")
        for k in sorted(acc, key=lambda k: (k[0], k[1])):
            print(f"   {SM.get(k, str(k)):58s} n={cnt[k]:4d}  mean {acc[k] / cnt[k]:8.0f}  total {acc[k]:9.0f} ({100 * acc[k] / span:4.1f} %)")
    for who in ("issuer_A", "issuer_B"):
        ev = d["trace"][who]
        if not ev:
            continue
        acc, cnt = defaultdict(float), defaultdict(int)
        for (i0, c0), (i1, c1) in zip(ev, ev[1:]):
            acc[(i0, i1)] += c1 - c0
            cnt[(i0, i1)] += 1
        span = ev[-1][1] - ev[0][1]
        print(f" {who}: {span} cycles")
        for k in sorted(acc):
            nm = {10: "before wait P.lo", 11: "before wait P.hi", 12: "P.lo ready", 13: "P.hi ready", 14: "S.lo issued", 15: "S.hi issued"}
            print(f"   {nm[k[0]]:18s} -> {nm[k[1]]:18s} n={cnt[k]:4d} mean {acc[k] / cnt[k]:8.0f} total {acc[k]:9.0f} ({100 * acc[k] / span:4.1f} %)")
