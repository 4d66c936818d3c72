"""Greedy generation on top of the native prefill: prefill (last-row logits + KV cache) then lp_decode_step per token.

Mirrors what `llava.generate(..., max_new_tokens=128, use_cache=True)` does through transformers' GenerationMixin in the
reference (evaluations/models/llava_multiimg_siglip_anyres.py:448-452): greedy arg-max, stop on any eos id.
Sampling/stopping is host logic; every tensor op on the path is a C-ABI call.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Set

import torch

from . import _lib


class GreedyDecoder:
    def __init__(self, engine):
        self.eng = engine

    @torch.no_grad()
    def run(self, ids: torch.Tensor, pixel_values: Optional[torch.Tensor], max_new_tokens: int, eos: Set[int]):
        eng = self.eng
        cfg = eng.cfg
        L = _lib.lib()
        want_cache = max_new_tokens > 1
        if not want_cache:
            logits, _, _ = eng.prefill([(ids, pixel_values)], "last")
            return logits.argmax(-1).to(torch.long)
        logits, kv, cu = eng.prefill([(ids, pixel_values)], "last", want_kv=True, kv_capacity=max_new_tokens)
        S = cu[-1]
        cap = kv.shape[2]
        ws = eng._workspace("decode", L.lp_decode_workspace_bytes(C.byref(eng.w.dec_dims)))
        step_logits = torch.empty((eng.w.vocab_padded,), dtype=torch.float32, device=eng.device)
        out = []
        tok = logits.argmax(-1).to(torch.long)  # [1] on the device
        for i in range(max_new_tokens):
            t = int(tok.item())                 # host read: stopping criterion, like GenerationMixin
            out.append(t)
            if t in eos or i == max_new_tokens - 1:
                break
            rc = L.lp_decode_step(C.byref(eng.w.dec_dims), C.byref(eng.w.dec), _lib.ptr(eng.w.embed),
                                  eng.w.embed.stride(0), _lib.ptr(tok), S + i, _lib.ptr(kv), cap,
                                  _lib.ptr(step_logits), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "lp_decode_step")
            tok = step_logits[: cfg.dec.vocab].argmax(-1, keepdim=True).to(torch.long)
        return torch.tensor(out, dtype=torch.long, device=eng.device)
