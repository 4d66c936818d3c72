"""Fused collective + GEMM launches for the tensor-parallel decoder (leopard_b200/tp.py, comm="fused"): symmetric (peer-mapped)
buffers, signal words and the two kernels of csrc/lp_gemm.cu / lp_tp.cu.

torch's symmetric-memory allocator is used for what it is — plumbing: it allocates buffers every rank can address over NVLink and
hands out the peers' pointers.  The data path is ours: `lp_gemm_bf16_tp` pulls the peers' rows inside the GEMM kernel
(all-gather) or TMA-stores partial tiles straight into the owner's memory from the epilogue (reduce-scatter); `lp_tp_reduce` sums
the landed partials in rank order.  No NCCL call on the layer path.

Buffers per rank (S = sequence rows, sl = S / tp, H = hidden):
    gather   bf16 [S, H]        rows [rank*sl, (rank+1)*sl) are written locally (RMSNorm output), the rest is pulled from the peers
    stage    bf16 [tp, sl, H]   slot [src] receives rank src's partial sums for this rank's rows
    signal   u32  [16]          word [r] all-gather readiness of rank r, word [8 + r] reduce-scatter completion of rank r
    counters u32  [256]         per-call CTA counters (zeroed slots are recycled by the kernels themselves)
Why no extra barriers: a rank cannot run ahead by more than one collective — a reduce-scatter completes only after EVERY rank has
finished the GEMM that follows its previous all-gather — so a buffer is never overwritten while a peer still reads it.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, ops


class FusedComm:
    def __init__(self, S: int, hidden: int, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.tp, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if S % self.tp:
            raise ValueError(f"S={S} must be a multiple of tp={self.tp}")
        self.S, self.sl, self.H = S, S // self.tp, hidden
        self.dev = torch.device(device if device is not None else torch.cuda.current_device())
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        self.gather = symm.empty((S, hidden), **bf)
        self.stage = symm.empty((self.tp, self.sl, hidden), **bf)
        self.signal = symm.empty((16,), dtype=torch.int32, device=self.dev)
        self.signal.zero_()
        self.gather.zero_()
        self._h = [symm.rendezvous(t, self.group) for t in (self.gather, self.stage, self.signal)]
        self.peer = [[int(h.buffer_ptrs[r]) for r in range(self.tp)] for h in self._h]
        self.counters = torch.zeros(256, dtype=torch.int32, device=self.dev)
        self.call = 0            # epoch: identical on every rank because every rank issues the same sequence of launches
        torch.cuda.synchronize()
        dist.barrier(self.group)

    def local_rows(self) -> torch.Tensor:
        """Where this rank's RMSNorm writes its rows of the next all-gather."""
        return self.gather[self.rank * self.sl:(self.rank + 1) * self.sl]

    def _desc(self, mode: int) -> _lib.TpDesc:
        self.call += 1
        t = _lib.TpDesc()
        t.tp_size, t.tp_rank, t.seg_rows, t.mode, t.epoch = self.tp, self.rank, self.sl, mode, self.call
        slot = self.call % 256
        t.counter = C.c_void_p(self.counters.data_ptr() + 4 * slot)
        t.recycle = C.c_void_p(self.counters.data_ptr() + 4 * ((slot + 128) % 256))
        t.signal_local = C.c_void_p(self.signal.data_ptr())
        for r in range(self.tp):
            t.peer_A[r] = self.peer[0][r]
            t.peer_stage[r] = self.peer[1][r]
            t.signal_peer[r] = self.peer[2][r]
        return t

    def allgather_gemm(self, w: torch.Tensor, act: int = ops.ACT_NONE) -> torch.Tensor:
        """act(all_gather(rows) @ w^T): the rank's own rows must already sit in `local_rows()`."""
        N = w.shape[0]
        out = torch.empty((self.S, N // 2 if act == ops.ACT_SWIGLU else N), dtype=torch.bfloat16, device=self.dev)
        d = _lib.GemmDesc()
        d.A, d.lda, d.W, d.ldw, d.C, d.ldc = _lib.ptr(self.gather), self.gather.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(out), out.stride(0)
        d.M, d.N, d.K, d.act = self.S, N, w.shape[1], act
        t = self._desc(_lib.TP_ALLGATHER)
        _lib.check(_lib.lib().lp_gemm_bf16_tp(C.byref(d), C.byref(t), _lib.stream_ptr()), "lp_gemm_bf16_tp(all-gather)")
        return out

    def gemm_reduce_scatter(self, a: torch.Tensor, w: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        """residual + reduce_scatter(a @ w^T) over ranks: a [S, K_local], w [H, K_local] -> this rank's rows [sl, H]."""
        d = _lib.GemmDesc()
        d.A, d.lda, d.W, d.ldw, d.C, d.ldc = _lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(self.stage), self.stage.stride(1)
        d.M, d.N, d.K = self.S, w.shape[0], w.shape[1]
        t = self._desc(_lib.TP_REDUCE_SCATTER)
        L = _lib.lib()
        _lib.check(L.lp_gemm_bf16_tp(C.byref(d), C.byref(t), _lib.stream_ptr()), "lp_gemm_bf16_tp(reduce-scatter)")
        out = torch.empty((self.sl, self.H), dtype=torch.bfloat16, device=self.dev)
        _lib.check(L.lp_tp_reduce(C.byref(t), _lib.ptr(self.stage), self.stage.stride(1), _lib.ptr(residual), residual.stride(0),
                                  _lib.ptr(out), out.stride(0), self.H, _lib.stream_ptr()), "lp_tp_reduce")
        return out
