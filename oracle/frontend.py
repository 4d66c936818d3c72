"""ORACLE (test infrastructure, never on the product path): CPU restatement of Leopard's image front-end.

What the reference does per image before the model sees it (evaluations/models/llava_multiimg_siglip_anyres.py):
  resize_and_pad_image  :102-140   aspect-preserving `PIL.Image.resize` (default resample = BICUBIC), centred on black
  divide_to_patches     :143-162   row-major 364 x 364 tiles
  crop order            :395-399   per image [original] + tiles
  image_processor.preprocess :403-404   SiglipImageProcessor: resize to 364 x 364 (bicubic), x 1/255, (x - 0.5) / 0.5
The arithmetic lives in two un-vendored third-party packages:
  * Pillow (requirements.txt: unpinned) — `ImagingResample` (src/libImaging/Resample.c): separable two-pass
    convolution, horizontal pass first, uint8 intermediate, coefficients computed in double precision and quantised to
    22 fractional bits (PRECISION_BITS = 32 - 8 - 2), each pass rounds to nearest and clips to [0, 255].
    Restated here from the published algorithm; pinned bit-exactly against the Pillow installed in the authoring
    container (12.2.0) by tests/test_frontend_oracle.py over random sizes (up- and down-scaling).
  * transformers' slow (PIL/numpy) SiglipImageProcessor — `rescale` multiplies in float64 and rounds to float32,
    `normalize` is (x - mean) / std in float32.  Pinned against transformers 5.5.0 `SiglipImageProcessorPil`.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CROP = 364


def bicubic_filter(x: float) -> float:
    """Keys cubic, a = -0.5 (Pillow Resample.c `bicubic_filter`)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bilinear_filter(x: float) -> float:
    """Triangle filter (Pillow Resample.c `bilinear_filter`, support 1.0) — Idefics2ImageProcessor's resample."""
    if x < 0.0:
        x = -x
    if x < 1.0:
        return 1.0 - x
    return 0.0


FILTERS = {"bicubic": (bicubic_filter, 2.0), "bilinear": (bilinear_filter, 1.0)}


def precompute_coeffs(in_size: int, out_size: int, filt: str = "bicubic"):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` for the whole-image box (in0 = 0, in1 = in_size).
    Returns (bounds int32 [out, 2] = (xmin, count), coeffs int32 [out, ksize], ksize)."""
    filter_fn, support0 = FILTERS[filt]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = filter_fn((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        bounds[xx] = (xmin, xmax)
        for x in range(ksize):
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
    return bounds, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis0(img: np.ndarray, out_size: int, filt: str = "bicubic") -> np.ndarray:
    """1-D pass along axis 0 of a uint8 array [n, ...] -> [out_size, ...]."""
    bounds, kk, _ = precompute_coeffs(img.shape[0], out_size, filt)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        w = kk[xx, :cnt].astype(np.int64).reshape((cnt,) + (1,) * (img.ndim - 1))
        acc = (src[xmin:xmin + cnt] * w).sum(0) + (1 << (PRECISION_BITS - 1))
        out[xx] = _clip8(acc)
    return out


def pil_resize(img: np.ndarray, out_w: int, out_h: int, filt: str = "bicubic") -> np.ndarray:
    """`PIL.Image.fromarray(img).resize((out_w, out_h), resample)` for an RGB uint8 array [H, W, 3] (no box, no
    reducing_gap): horizontal pass first, then vertical, each skipped when that size is unchanged — except for
    images more than 100x taller than wide (see below)."""
    h, w = img.shape[:2]
    cur = img
    if h > w * 100 and out_h < h:
        # Pillow >= 11 (Image.resize, Python side): very tall images that shrink vertically are resampled vertically
        # FIRST at full width, then horizontally — observable in the result because of the uint8 intermediate
        cur = _resample_axis0(cur, out_h, filt)
        if out_w != w:
            cur = _resample_axis0(cur.transpose(1, 0, 2), out_w, filt).transpose(1, 0, 2)
        return np.ascontiguousarray(cur)
    if out_w != w:
        cur = _resample_axis0(cur.transpose(1, 0, 2), out_w, filt).transpose(1, 0, 2)
    if out_h != h:
        cur = _resample_axis0(cur, out_h, filt)
    return np.ascontiguousarray(cur)


def pil_resize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Default resample of `Image.resize` = BICUBIC (what resize_and_pad_image and SiglipImageProcessor use)."""
    return pil_resize(img, out_w, out_h, "bicubic")


# ---- Idefics2 (evaluations/models/idefics2_multiimg.py:22-25: AutoProcessor(size={'longest_edge': R, 'shortest_edge': 0},
# do_image_splitting=False) -> transformers' Idefics2ImageProcessor, un-vendored third-party code restated here and pinned against the
# installed PIL-backed implementation, tests/golden/gen_idefics2_frontend_goldens.py) ------------------------------------------------
def idefics2_resize_size(w: int, h: int, longest_edge: int, shortest_edge: int = 0) -> Tuple[int, int]:
    """`get_resize_output_image_size`: only images whose longer side exceeds `longest_edge` shrink (aspect ratio kept, the
    other side truncated), then both sides are raised to `shortest_edge`.  Returns (width, height)."""
    aspect = w / h
    if w >= h and w > longest_edge:
        w = longest_edge
        h = int(w / aspect)
    elif h > w and h > longest_edge:
        h = longest_edge
        w = int(h * aspect)
    return max(w, shortest_edge), max(h, shortest_edge)


def idefics2_preprocess(images: List[np.ndarray], longest_edge: int = 980, shortest_edge: int = 0):
    """One sample's images (RGB uint8 [H, W, 3]) -> (pixel_values fp32 [n, 3, Hmax, Wmax], pixel_attention_mask int64
    [n, Hmax, Wmax]): BILINEAR `Image.resize`, x * (1/255) in float64 rounded to float32, (x - 0.5) / 0.5 in float32, then zero
    padding (AFTER the normalisation: padded pixels are 0.0, not normalised black) at the bottom / right up to the largest image."""
    lut = normalize_lut()
    outs = []
    for img in images:
        h, w = img.shape[:2]
        nw, nh = idefics2_resize_size(w, h, longest_edge, shortest_edge)
        r = pil_resize(img, nw, nh, "bilinear")
        outs.append(np.ascontiguousarray(lut[r].transpose(2, 0, 1)))
    hm, wm = max(o.shape[1] for o in outs), max(o.shape[2] for o in outs)
    pv = np.zeros((len(outs), 3, hm, wm), np.float32)
    mask = np.zeros((len(outs), hm, wm), np.int64)
    for i, o in enumerate(outs):
        pv[i, :, :o.shape[1], :o.shape[2]] = o
        mask[i, :o.shape[1], :o.shape[2]] = 1
    return pv, mask


def pad_geometry(w: int, h: int, target: Optional[Tuple[int, int]]):
    """resize_and_pad_image's sizes (EV/...anyres.py:116-136): -> (new_w, new_h, paste_x, paste_y) or None."""
    if target is None:
        return None
    tw, th = target
    scale_w = tw / w
    scale_h = th / h
    if scale_w < scale_h:
        nw = tw
        nh = min(math.ceil(h * scale_w), th)
    else:
        nh = th
        nw = min(math.ceil(w * scale_h), tw)
    return nw, nh, (tw - nw) // 2, (th - nh) // 2


def resize_and_pad(img: np.ndarray, target: Optional[Tuple[int, int]]) -> Optional[np.ndarray]:
    if target is None:
        return None
    h, w = img.shape[:2]
    nw, nh, px, py = pad_geometry(w, h, target)
    out = np.zeros((target[1], target[0], 3), np.uint8)
    out[py:py + nh, px:px + nw] = pil_resize_bicubic(img, nw, nh)
    return out


def divide_to_tiles(img: np.ndarray, size: int = CROP) -> List[np.ndarray]:
    h, w = img.shape[:2]
    return [img[i:i + size, j:j + size] for i in range(0, h, size) for j in range(0, w, size)]


def normalize_lut() -> np.ndarray:
    """fp32 value of every uint8 level after SiglipImageProcessor's rescale (float64 multiply, rounded to float32) and
    normalize ((x - 0.5) / 0.5 in float32)."""
    lv = np.arange(256, dtype=np.uint8)
    x = (lv.astype(np.float64) * (1 / 255)).astype(np.float32)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)


def siglip_preprocess(img: np.ndarray, size: int = CROP) -> np.ndarray:
    """SiglipImageProcessor.preprocess of one RGB uint8 image [H, W, 3] -> fp32 [3, size, size]."""
    r = pil_resize_bicubic(img, size, size)
    return normalize_lut()[r].transpose(2, 0, 1).copy()


def frontend(img: np.ndarray, target: Optional[Tuple[int, int]], train_order: bool = False) -> np.ndarray:
    """One image -> pixel_values [1 + tiles, 3, 364, 364] fp32 in the eval order [original] + tiles
    (EV/...anyres.py:395-399); train_order puts the original last (mm_pretrain_dataset.py:247-253)."""
    crops = [siglip_preprocess(img)]
    padded = resize_and_pad(img, target)
    tiles = [siglip_preprocess(t) for t in divide_to_tiles(padded)] if padded is not None else []
    crops = tiles + crops if train_order else crops + tiles
    return np.stack(crops, 0)
