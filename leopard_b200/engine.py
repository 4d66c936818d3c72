"""PrefillEngine — host-side sequencing of the native prefill path (allocation, workspaces, shapes only).

Everything arithmetic is a C-ABI call into lib/libleopard_b200.so; this class owns torch buffers and mirrors the order
of operations of `myLlavaForConditionalGeneration.forward`'s prefill branch
(evaluations/models/llava_multiimg_siglip_anyres.py:259-333): vision tower -> pixel-shuffle + projector -> merge ->
decoder -> logits.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .config import LeopardConfig
from .native_weights import NativeWeights


class PrefillEngine:
    def __init__(self, cfg: LeopardConfig, weights: NativeWeights):
        if not torch.cuda.is_available():
            raise _lib.NativeError("leopard_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.w = weights
        self.device = weights.device
        self._ws = {}
        self._decode_state = None   # device-resident greedy-loop state + captured CUDA graph (leopard_b200/decode.py)
        self._decode_kv = None      # persistent KV-cache buffer of generate() (capacity grows in KV_GRANULE steps)
        _lib.lib()  # fail early and loudly when the extension is missing
        _lib.ensure_streamk_workspace(self.device)

    # ------------------------------------------------------------------ buffers
    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    # ------------------------------------------------------------------ stages
    def vit(self, pixel_values: torch.Tensor, shuffle: bool = True) -> torch.Tensor:
        """[N,3,364,364] fp32/bf16 (device) -> pixel-shuffled ln_post output [N*169, 4608] (or [N*676, 1152])."""
        v = self.cfg.vit
        if pixel_values.dim() != 4 or pixel_values.shape[1:] != (3, v.image_size, v.image_size):
            raise ValueError(f"pixel_values must be [N,3,{v.image_size},{v.image_size}], got {tuple(pixel_values.shape)}")
        if pixel_values.dtype not in (torch.float32, torch.bfloat16):
            pixel_values = pixel_values.float()
        pixel_values = pixel_values.to(self.device).contiguous()
        n = pixel_values.shape[0]
        L = _lib.lib()
        self.w.sync_folded("vit")
        if shuffle:
            out = torch.empty((n * self.cfg.tokens_per_crop, self.cfg.proj_in), dtype=torch.bfloat16, device=self.device)
        else:
            out = torch.empty((n * v.tokens, v.hidden), dtype=torch.bfloat16, device=self.device)
        nbytes = L.lp_vit_workspace_bytes(C.byref(self.w.vit_dims), n)
        ws = self._workspace("vit", nbytes)
        rc = L.lp_vit_forward(C.byref(self.w.vit_dims), C.byref(self.w.vit), _lib.ptr(pixel_values),
                              1 if pixel_values.dtype == torch.float32 else 0, n, _lib.ptr(out), out.stride(0),
                              1 if shuffle else 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "lp_vit_forward")
        return out

    def project(self, shuffled: torch.Tensor) -> torch.Tensor:
        """[rows, 4608] -> [rows, 4096]."""
        rows = shuffled.shape[0]
        hid = self.cfg.dec.hidden
        L = _lib.lib()
        out = torch.empty((rows, hid), dtype=torch.bfloat16, device=self.device)
        ws = self._workspace("proj", L.lp_project_workspace_bytes(rows, hid))
        rc = L.lp_project(C.byref(self.w.proj), _lib.ptr(shuffled), rows, shuffled.shape[1], hid, _lib.ptr(out),
                          out.stride(0), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "lp_project")
        return out

    def image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """multi_modal_projector(vision_tower(pixel_values).last_hidden_state) -> [N, 169, 4096]."""
        n = pixel_values.shape[0]
        return self.project(self.vit(pixel_values, shuffle=True)).view(n, self.cfg.tokens_per_crop, -1)

    def merge(self, ids: torch.Tensor, feats: Optional[torch.Tensor], image_id: int, rows_per_slot: int,
              out: Optional[torch.Tensor] = None, n_slots: Optional[int] = None) -> Tuple[torch.Tensor, int]:
        """One sequence: ids [L] int64, feats [rows, D] -> (embeds [S, D], n_slots).  `n_slots` = number of image slots when the
        caller already knows it (counting them on device ids costs a device -> host sync)."""
        if n_slots is None:
            n_slots = int((ids == image_id).sum())      # free for host ids; one sync for device ids
        ids = ids.to(self.device, torch.long).contiguous()
        Lt = ids.numel()
        n_feat = 0 if feats is None else feats.shape[0]
        S = Lt + n_slots * (rows_per_slot - 1)
        D = self.cfg.dec.hidden
        if out is None:
            out = torch.empty((S, D), dtype=torch.bfloat16, device=self.device)
        plan = torch.empty(2 * Lt + 2, dtype=torch.int32, device=self.device)
        rc = _lib.lib().lp_embed_merge(_lib.ptr(ids), Lt, image_id, rows_per_slot, _lib.ptr(self.w.embed),
                                       self.w.embed.stride(0), self.w.embed.shape[0], _lib.ptr(feats),
                                       feats.stride(0) if feats is not None else 0, n_feat, _lib.ptr(out),
                                       out.stride(0), out.shape[0], D, _lib.ptr(plan), _lib.stream_ptr())
        _lib.check(rc, "lp_embed_merge")
        return out[:S], n_slots

    def decoder(self, x: torch.Tensor, pos_ids: torch.Tensor, cu_seqlens: Sequence[int], logits_mode: str = "last",
                logits_dtype=torch.float32, want_kv: bool = False, kv_capacity: int = 0,
                kv_buffer: Optional[torch.Tensor] = None):
        """x [R, D] packed merged embeddings (overwritten) -> (logits, kv_cache or None).
        logits: 'last' -> [n_seqs, V]; 'all' -> [R, V]; 'none' -> None.
        kv_buffer: caller-owned cache [layers, 2, capacity >= R, kv_w] to fill instead of allocating one."""
        d = self.cfg.dec
        R = x.shape[0]
        n = len(cu_seqlens) - 1
        L = _lib.lib()
        self.w.sync_folded("dec")
        mode = {"none": _lib.LOGITS_NONE, "last": _lib.LOGITS_LAST, "all": _lib.LOGITS_ALL}[logits_mode]
        vp = self.w.vocab_padded
        logits = None
        if mode == _lib.LOGITS_LAST:
            logits = torch.empty((n, vp), dtype=logits_dtype, device=self.device)
        elif mode == _lib.LOGITS_ALL:
            logits = torch.empty((R, vp), dtype=logits_dtype, device=self.device)
        kv = None
        if want_kv and kv_buffer is not None:
            if (kv_buffer.dim() != 4 or kv_buffer.shape[0] != d.layers or kv_buffer.shape[2] < R
                    or kv_buffer.shape[3] != d.kv_heads * d.head_dim or not kv_buffer.is_contiguous()):
                raise ValueError(f"kv_buffer {tuple(kv_buffer.shape)} cannot hold {R} rows of this decoder")
            kv = kv_buffer
        elif want_kv:
            kv = torch.empty((d.layers, 2, max(R, kv_capacity), d.kv_heads * d.head_dim), dtype=torch.bfloat16,
                             device=self.device)
        cu = (C.c_int * (n + 1))(*[int(c) for c in cu_seqlens])
        ws = self._workspace("dec", L.lp_decoder_workspace_bytes(C.byref(self.w.dec_dims), R, n))
        pos = pos_ids.to(self.device, torch.int32).contiguous()
        rc = L.lp_decoder_prefill(C.byref(self.w.dec_dims), C.byref(self.w.dec), _lib.ptr(x), _lib.ptr(pos), cu, n,
                                  _lib.ptr(kv), kv.shape[2] if kv is not None else 0, _lib.ptr(logits), logits.stride(0) if logits is not None else 0, mode,
                                  1 if logits_dtype == torch.float32 else 0, _lib.ptr(ws), ws.numel(),
                                  _lib.stream_ptr())
        _lib.check(rc, "lp_decoder_prefill")
        if logits is not None and vp != d.vocab:
            logits = logits[:, : d.vocab]
        return logits, kv

    # ------------------------------------------------------------------ whole path
    @torch.no_grad()
    def prefill(self, samples: List[Tuple[torch.Tensor, Optional[torch.Tensor]]], logits_mode: str = "last",
                logits_dtype=torch.float32, want_kv: bool = False, convention: str = "hf", kv_capacity: int = 0,
                kv_buffer: Optional[torch.Tensor] = None, slots: Optional[Sequence[int]] = None):
        """samples: [(input_ids [L] or [1,L], pixel_values [N_i,3,364,364] or None), ...] — independent requests.
        All crops go through the tower as one batch, every request becomes one packed decoder sequence.
        convention 'hf': one placeholder id per crop, expanded to 169 rows; 'megatron': one -200 per feature row.
        Returns (logits, kv, cu_seqlens)."""
        cfg = self.cfg
        pix = [p for _, p in samples if p is not None and p.shape[0] > 0]
        feats_all = None
        if pix:
            allp = torch.cat([p.to(self.device) for p in pix], 0) if len(pix) > 1 else pix[0]
            feats_all = self.project(self.vit(allp, shuffle=True))  # [sum N * 169, D]
        image_id, per_slot = (cfg.image_token_index, cfg.tokens_per_crop) if convention == "hf" else (
            cfg.megatron_image_index, 1)
        # plan lengths on the host (shapes depend on the ids, exactly as in the reference's merge): ids that arrive as host
        # tensors — every caller on the plugin surfaces — are counted on the host, so no device -> host sync happens per request;
        # device ids can bring their slot counts (`slots`), else they cost one sync each
        host_ids = [i.reshape(-1) for i, _ in samples]
        if slots is None:
            slots = [int((i == image_id).sum()) for i in host_ids]
        ids_list = [i.to(self.device, torch.long, non_blocking=True) for i in host_ids]
        lens = [i.numel() + s * (per_slot - 1) for i, s in zip(ids_list, slots)]
        cu = [0]
        for ln in lens:
            cu.append(cu[-1] + ln)
        x = torch.empty((cu[-1], cfg.dec.hidden), dtype=torch.bfloat16, device=self.device)
        f0 = 0
        for k, ((_, p), ids) in enumerate(zip(samples, ids_list)):
            n_rows = 0 if p is None else p.shape[0] * cfg.tokens_per_crop
            if slots[k] * per_slot != n_rows:
                if convention == "hf":
                    raise ValueError(
                        f"The input provided to the model are wrong. The number of image tokens is {slots[k]} while "
                        f"the number of image given to the model is {0 if p is None else p.shape[0]}. "
                        "This prevents correct indexing and breaks batch generation.")
                raise AssertionError("The number of image features doesn't match the number of image tokens. "
                                     f"{n_rows} != {slots[k]}")
            f = feats_all[f0:f0 + n_rows] if n_rows else None
            self.merge(ids, f, image_id, per_slot, out=x[cu[k]:cu[k + 1]], n_slots=slots[k])
            f0 += n_rows
        pos = torch.cat([torch.arange(ln, dtype=torch.int32) for ln in lens]).to(self.device, non_blocking=True) \
            if len(lens) > 1 else torch.arange(lens[0], dtype=torch.int32, device=self.device)
        logits, kv = self.decoder(x, pos, cu, logits_mode, logits_dtype, want_kv, cu[-1] + kv_capacity, kv_buffer)
        return logits, kv, cu

    def merged_length(self, ids: torch.Tensor, pixel_values: Optional[torch.Tensor], convention: str = "hf") -> int:
        """Decoder rows one request occupies after the merge (L + 168 per image placeholder in the HF convention)."""
        cfg = self.cfg
        image_id, per_slot = (cfg.image_token_index, cfg.tokens_per_crop) if convention == "hf" else (
            cfg.megatron_image_index, 1)
        flat = ids.reshape(-1)
        return int(flat.numel() + int((flat == image_id).sum()) * (per_slot - 1))
