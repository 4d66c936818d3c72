"""Greedy generation on top of the native prefill: prefill (last-row logits + KV cache), then one native step per token.

Mirrors what `llava.generate(..., max_new_tokens=128, use_cache=True)` does through transformers' GenerationMixin in the
reference (evaluations/models/llava_multiimg_siglip_anyres.py:448-452): greedy arg-max, stop on any eos id.

The loop state (last token, position) lives on the device: `lp_decode_greedy_step` embeds the token, runs the 32 layers
as HBM-streaming GEMV kernels over the KV cache, picks the arg-max and advances the position.  By default the step is
captured once into a CUDA graph (`lp_decode_graph_create`) and replayed per token; the only host work per token is
reading the 8-byte token back for the stopping criterion, like GenerationMixin does.
  LEOPARD_DECODE_GRAPH=0   launch the step's kernels one by one instead of replaying the graph
  LEOPARD_DECODE_MODE=n    lp_decode_set_mode(n): bit 0 GEMV kernels, bit 1 programmatic dependent launch, bit 2 block-
                           cooperative GEMV (default 7 = all three), 21 = one persistent kernel per token
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Set

import torch

from . import _lib

KV_GRANULE = 2048  # cache capacities are rounded up to this many rows so that one captured graph serves many requests


class _DecodeState:
    """Per-engine device buffers the captured graph points at (stable addresses across generate() calls)."""

    def __init__(self, eng):
        dev = eng.device
        self.tok = torch.zeros(1, dtype=torch.long, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.logits = torch.empty((eng.w.vocab_padded,), dtype=torch.float32, device=dev)
        self.graph = None      # lp_decode_graph*
        self.graph_key = None

    def drop_graph(self):
        if self.graph is not None:
            _lib.lib().lp_decode_graph_destroy(self.graph)
            self.graph, self.graph_key = None, None

    def __del__(self):  # pragma: no cover
        try:
            self.drop_graph()
        except Exception:
            pass


class GreedyDecoder:
    def __init__(self, engine):
        self.eng = engine
        if getattr(engine, "_decode_state", None) is None:
            engine._decode_state = _DecodeState(engine)
        self.st = engine._decode_state

    def _kv_buffer(self, rows_needed: int):
        """A persistent cache buffer [layers, 2, capacity, kv_w] whose capacity only grows (in KV_GRANULE steps)."""
        d = self.eng.cfg.dec
        cap = -(-rows_needed // KV_GRANULE) * KV_GRANULE
        kv = getattr(self.eng, "_decode_kv", None)
        if kv is None or kv.shape[2] < cap:
            self.st.drop_graph()
            self.eng._decode_kv = None
            kv = torch.empty((d.layers, 2, cap, d.kv_heads * d.head_dim), dtype=torch.bfloat16, device=self.eng.device)
            self.eng._decode_kv = kv
        return kv

    @torch.no_grad()
    def run(self, ids: torch.Tensor, pixel_values: Optional[torch.Tensor], max_new_tokens: int, eos: Set[int]):
        eng, st = self.eng, self.st
        cfg = eng.cfg
        L = _lib.lib()
        if max_new_tokens <= 1:
            logits, _, _ = eng.prefill([(ids, pixel_values)], "last")
            return logits.argmax(-1).to(torch.long)
        mode = os.environ.get("LEOPARD_DECODE_MODE")
        if mode is not None:
            L.lp_decode_set_mode(int(mode))
        n_rows = eng.merged_length(ids, pixel_values)
        kv = self._kv_buffer(n_rows + max_new_tokens)
        logits, kv, cu = eng.prefill([(ids, pixel_values)], "last", want_kv=True, kv_buffer=kv)
        return self.continue_from(logits, kv, cu[-1], max_new_tokens, eos, mode)

    @torch.no_grad()
    def continue_from(self, logits: torch.Tensor, kv: torch.Tensor, S: int, max_new_tokens: int, eos: Set[int],
                      mode: Optional[str] = None):
        """Greedy loop after ANY prefill that left `S` rows in the cache `kv` [layers, 2, capacity, kv_w] and produced the
        last row's `logits` (shared by the LLaVA and Idefics2 shims).  Returns the new token ids (1-D, on the device)."""
        eng, st = self.eng, self.st
        cfg = eng.cfg
        L = _lib.lib()
        cap = kv.shape[2]
        if S + max_new_tokens - 1 > cap:
            raise ValueError(f"KV cache capacity {cap} cannot hold {S} prompt rows + {max_new_tokens - 1} new tokens")
        ws = eng._workspace("decode", L.lp_decode_workspace_bytes(C.byref(eng.w.dec_dims)))
        st.tok.copy_(logits.reshape(-1, logits.shape[-1])[-1:, : cfg.dec.vocab].argmax(-1).to(torch.long))
        st.pos.fill_(S)
        args = (C.byref(eng.w.dec_dims), C.byref(eng.w.dec), _lib.ptr(eng.w.embed), eng.w.embed.stride(0),
                _lib.ptr(st.tok), _lib.ptr(st.pos), _lib.ptr(kv), cap, _lib.ptr(st.logits), cfg.dec.vocab,
                _lib.ptr(ws), ws.numel())
        use_graph = os.environ.get("LEOPARD_DECODE_GRAPH", "1") != "0"
        if use_graph:
            key = (kv.data_ptr(), cap, ws.data_ptr(), st.logits.data_ptr(), mode)
            if st.graph_key != key:
                st.drop_graph()
                h = C.c_void_p()
                _lib.check(L.lp_decode_graph_create(*args, C.byref(h)), "lp_decode_graph_create")
                st.graph, st.graph_key = h, key
        out = []
        for i in range(max_new_tokens):
            t = int(st.tok.item())                 # host read: stopping criterion, like GenerationMixin
            out.append(t)
            if t in eos or i == max_new_tokens - 1:
                break
            if use_graph:
                _lib.check(L.lp_decode_graph_launch(st.graph, _lib.stream_ptr()), "lp_decode_graph_launch")
            else:
                _lib.check(L.lp_decode_greedy_step(*args, _lib.stream_ptr()), "lp_decode_greedy_step")
        return torch.tensor(out, dtype=torch.long, device=eng.device)
