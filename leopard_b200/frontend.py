"""GPU image front-end: raw RGB uint8 images -> `pixel_values [N, 3, 364, 364]` in the reference's crop order.

Replaces lines 385-405 of evaluations/models/llava_multiimg_siglip_anyres.py (allocate_patches ->
select_best_resolution -> resize_and_pad_image -> divide_to_patches -> `[origin] + patches` ->
SiglipImageProcessor.preprocess per crop -> torch.cat) with host planning (integers only, `preprocess.py`) plus one
`lp_image_frontend` call per image; the uint8 pixels cross PCIe once (3.6x fewer bytes than fp32 pixel_values) and the
resize / pad / tile / normalise arithmetic runs in hand-written kernels that are bit-identical to Pillow + the
PIL-backed SiglipImageProcessor.  There is no CPU fallback: without the native library a `NativeError` is raised.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .preprocess import PATCH, TOTAL_CROP_BUDGET, allocate_patches, select_best_resolution

SIGLIP_RESCALE = 1 / 255          # preprocessor defaults of SigLIP (the reference's preprocessor_config.json is not in
SIGLIP_MEAN = (0.5, 0.5, 0.5)     # its tree; these are SiglipImageProcessor's class defaults)
SIGLIP_STD = (0.5, 0.5, 0.5)


def _as_uint8_hwc(img) -> torch.Tensor:
    """PIL image / numpy array / torch tensor -> contiguous uint8 [H, W, 3] torch tensor (host or device)."""
    if isinstance(img, torch.Tensor):
        t = img
    else:
        if hasattr(img, "convert"):          # PIL.Image: same conversion the reference applies at load (EV :385)
            img = np.array(img.convert("RGB"))   # a writable copy (PIL exposes a read-only buffer)
        img = np.ascontiguousarray(img)
        if not img.flags.writeable:
            img = img.copy()
        t = torch.from_numpy(img)
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError(f"image must be uint8 [H, W, 3], got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def plan_request(sizes: Sequence[Tuple[int, int]]) -> List[Optional[Tuple[int, int]]]:
    """Target (width, height) of every image's tile grid, or None for 'no tiles' — the reference's planning at
    EV/...anyres.py:387-392 (budget = 50 - n_images; no budget left -> originals only)."""
    budget = TOTAL_CROP_BUDGET - len(sizes)
    if budget <= 0:
        return [None] * len(sizes)
    counts = allocate_patches(list(sizes), patch_budget=budget)
    return [select_best_resolution(s, n) for s, n in zip(sizes, counts)]


class ImageFrontend:
    """`fe = ImageFrontend('cuda:0'); pixel_values, per_image = fe(images)`"""

    def __init__(self, device="cuda:0", rescale: float = SIGLIP_RESCALE, mean=SIGLIP_MEAN, std=SIGLIP_STD):
        if not torch.cuda.is_available():
            raise _lib.NativeError("leopard_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.rescale = float(rescale)
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)
        self._ws = None
        _lib.lib()

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    @torch.no_grad()
    def __call__(self, images, dtype=torch.float32, train_order: bool = False):
        """images: list of PIL images / uint8 [H, W, 3] arrays or tensors.  Returns (pixel_values [N, 3, 364, 364] on
        the device, crops-per-image).  Order per image: [original] + tiles (eval, EV/...anyres.py:395-399) or
        tiles + [original] (train_order, mm_pretrain_dataset.py:247-253)."""
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be float32 or bfloat16")
        L = _lib.lib()
        raws = [_as_uint8_hwc(im) for im in images]
        sizes = [(int(t.shape[1]), int(t.shape[0])) for t in raws]          # PIL convention: (width, height)
        targets = plan_request(sizes)
        per_image = [1 + (0 if t is None else (t[0] // PATCH) * (t[1] // PATCH)) for t in targets]
        out = torch.empty((sum(per_image), 3, PATCH, PATCH), dtype=dtype, device=self.device)
        n0 = 0
        for raw, (w, h), tgt, n in zip(raws, sizes, targets, per_image):
            dev = raw.to(self.device, non_blocking=True)
            tw, th = tgt if tgt is not None else (0, 0)
            ws = self._workspace(L.lp_image_frontend_workspace_bytes(h, w, tw, th))
            g_idx, t_idx = (n0 + n - 1, n0) if train_order else (n0, n0 + 1)
            tiles_ptr = _lib.ptr(out[t_idx]) if n > 1 else None
            rc = L.lp_image_frontend(_lib.ptr(dev), h, w, tw, th, _lib.ptr(out[g_idx]), tiles_ptr,
                                     1 if dtype == torch.float32 else 0, self.rescale, self._mean, self._std,
                                     _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "lp_image_frontend")
            n0 += n
        return out, per_image


# ---- Idefics2 (evaluations/models/idefics2_multiimg.py:22-25: AutoProcessor(size={'longest_edge': R, 'shortest_edge': 0},
# do_image_splitting=False)) -------------------------------------------------------------------------------------------------------
def idefics2_resize_size(w: int, h: int, longest_edge: int, shortest_edge: int = 0) -> Tuple[int, int]:
    """transformers' `get_resize_output_image_size` for Idefics2 (host integers, same float arithmetic): only an image whose
    longer side exceeds `longest_edge` shrinks (aspect ratio kept, the other side truncated); both sides are then raised to
    `shortest_edge`.  Returns (width, height)."""
    aspect = w / h
    if w >= h and w > longest_edge:
        w = longest_edge
        h = int(w / aspect)
    elif h > w and h > longest_edge:
        h = longest_edge
        w = int(h * aspect)
    return max(w, shortest_edge), max(h, shortest_edge)


class Idefics2Frontend:
    """`pixel_values, pixel_attention_mask = Idefics2Frontend('cuda:0', longest_edge=980)(images)` — the image half of the
    reference's `processor(text=..., images=images, return_tensors='pt')` (idefics2_multiimg.py:88-90) for ONE sample: BILINEAR
    resize by the longest-edge rule, rescale, normalise, zero-pad to the largest image, pixel mask.  uint8 pixels cross PCIe, the
    arithmetic runs in `lp_image_resize` and is bit-identical to transformers' PIL-backed Idefics2ImageProcessor
    (tests/golden/ref_idefics2_frontend.json).  Returns pixel_values [1, n, 3, Hmax, Wmax] and pixel_attention_mask [1, n, Hmax, Wmax]
    on the device — the tensors `Idefics2ForConditionalGeneration.generate` takes."""

    def __init__(self, device="cuda:0", longest_edge: int = 980, shortest_edge: int = 0, rescale: float = 1 / 255,
                 mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
        if not torch.cuda.is_available():
            raise _lib.NativeError("leopard_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.longest_edge, self.shortest_edge = int(longest_edge), int(shortest_edge)
        self.rescale = float(rescale)
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)
        self._ws = None
        _lib.lib()

    @torch.no_grad()
    def __call__(self, images, dtype=torch.float32):
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be float32 or bfloat16")
        L = _lib.lib()
        raws = [_as_uint8_hwc(im) for im in images]
        sizes = [idefics2_resize_size(int(t.shape[1]), int(t.shape[0]), self.longest_edge, self.shortest_edge) for t in raws]
        wm, hm = max(s[0] for s in sizes), max(s[1] for s in sizes)
        pv = torch.empty((1, len(raws), 3, hm, wm), dtype=dtype, device=self.device)
        mask = torch.empty((1, len(raws), hm, wm), dtype=torch.long, device=self.device)
        for i, (raw, (ow, oh)) in enumerate(zip(raws, sizes)):
            h, w = int(raw.shape[0]), int(raw.shape[1])
            dev = raw.to(self.device, non_blocking=True)
            need = L.lp_image_resize_workspace_bytes(h, w, ow, oh, 1)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
            rc = L.lp_image_resize(_lib.ptr(dev), h, w, ow, oh, 1, wm, hm, 0.0, _lib.ptr(pv[0, i]),
                                   1 if dtype == torch.float32 else 0, _lib.ptr(mask[0, i]), self.rescale, self._mean, self._std,
                                   _lib.ptr(self._ws), self._ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "lp_image_resize")
        return pv, mask
