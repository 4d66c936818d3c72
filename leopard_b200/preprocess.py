"""Host-side adaptive-resolution crop planning (drives every downstream shape of the prefill path).

Same names, argument meaning and results as the reference functions in
evaluations/models/llava_multiimg_siglip_anyres.py:26-162 (training twin:
Pai-Megatron-Patch/megatron_patch/data/llava/mm_pretrain_dataset.py:65-201); written from their behaviour, pinned by
tests/golden/ref_crops.json which was produced by executing the reference functions themselves.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

PATCH = 364
TOTAL_CROP_BUDGET = 50  # `budget = 50 - len(images)`, EV/llava_multiimg_siglip_anyres.py:387


def allocate_patches(image_sizes: Sequence[Tuple[int, int]], patch_size: int = PATCH, patch_budget: int = 50) -> List[int]:
    """Sub-crop count per image: round(a/patch)*round(b/patch) (a single tile counts as 0), scaled down
    proportionally (floor) when the sum exceeds the budget.  Uses Python's round() (banker's rounding) like the
    reference (EV/...anyres.py:26-58)."""
    wanted = []
    for a, b in image_sizes:
        n = round(a / patch_size) * round(b / patch_size)
        wanted.append(0 if n == 1 else n)
    total = sum(wanted)
    if total <= patch_budget:
        return wanted
    ratio = patch_budget / total
    scaled = [int(n * ratio) for n in wanted]
    # floor scaling can never exceed the budget, but keep the reference's trimming loop semantics
    while sum(scaled) > patch_budget:
        over = sum(scaled) - patch_budget
        for i, n in enumerate(scaled):
            if n > 0:
                scaled[i] = n - 1
                over -= 1
            if over == 0:
                break
    return scaled


def select_best_resolution(original_size: Tuple[int, int], num_patches: int, patch_size: int = PATCH) -> Optional[Tuple[int, int]]:
    """Best (width, height) grid of `patch_size` tiles with rows*cols <= num_patches, (1,1) excluded: maximise the
    effective (downscaled, capped) resolution, break ties by least wasted area, first candidate wins among equals in
    row-major (row, col) order (EV/...anyres.py:61-99)."""
    if num_patches == 0:
        return None
    ow, oh = original_size
    best = None
    best_eff, best_waste = 0, float("inf")
    for rows in range(1, num_patches + 1):
        for cols in range(1, num_patches + 1):
            if rows * cols > num_patches or (rows == 1 and cols == 1):
                continue
            w, h = cols * patch_size, rows * patch_size
            s = min(w / ow, h / oh)
            dw, dh = int(ow * s), int(oh * s)
            eff = min(dw * dh, ow * oh)
            waste = w * h - eff
            if eff > best_eff or (eff == best_eff and waste < best_waste):
                best_eff, best_waste, best = eff, waste, (w, h)
    if best == (patch_size, patch_size):
        return None
    return best


def resize_and_pad_image(image, target_resolution):
    """Aspect-preserving PIL resize (default filter) centred on a black canvas (EV/...anyres.py:102-140)."""
    if target_resolution is None:
        return None
    from PIL import Image
    ow, oh = image.size
    tw, th = target_resolution
    sw, sh = tw / ow, th / oh
    if sw < sh:
        nw, nh = tw, min(math.ceil(oh * sw), th)
    else:
        nh, nw = th, min(math.ceil(ow * sh), tw)
    canvas = Image.new("RGB", (tw, th), (0, 0, 0))
    canvas.paste(image.resize((nw, nh)), ((tw - nw) // 2, (th - nh) // 2))
    return canvas


def divide_to_patches(image, patch_size: int = PATCH):
    """Row-major patch_size x patch_size tiles (EV/...anyres.py:143-162)."""
    w, h = image.size
    return [image.crop((x, y, x + patch_size, y + patch_size))
            for y in range(0, h, patch_size) for x in range(0, w, patch_size)]


def plan_crops(image_sizes: Sequence[Tuple[int, int]]) -> List[int]:
    """Number of ViT inputs per image in the eval convention ([global] + tiles, EV/...anyres.py:387-401):
    budget = 50 - n_images; with no budget left every image is a single global crop."""
    budget = TOTAL_CROP_BUDGET - len(image_sizes)
    if budget <= 0:
        return [1] * len(image_sizes)
    counts = allocate_patches(list(image_sizes), patch_budget=budget)
    out = []
    for size, n in zip(image_sizes, counts):
        res = select_best_resolution(size, n)
        tiles = 0 if res is None else (res[0] // PATCH) * (res[1] // PATCH)
        out.append(1 + tiles)
    return out


def preprocess_images(images, image_processor=None):
    """PIL images -> (list of PIL crops in eval order, crops-per-image).  `image_processor` (SiglipImageProcessor in
    the reference, EV/...anyres.py:378,403-404) is applied by the caller; without one, `to_pixel_values` below does
    the same resize -> 1/255 -> (x-0.5)/0.5."""
    budget = TOTAL_CROP_BUDGET - len(images)
    if budget <= 0:
        return list(images), [1] * len(images)
    counts = allocate_patches([im.size for im in images], patch_budget=budget)
    crops, per_image = [], []
    for im, n in zip(images, counts):
        padded = resize_and_pad_image(im, select_best_resolution(im.size, n))
        tiles = divide_to_patches(padded, PATCH) if padded is not None else []
        crops += [im] + tiles
        per_image.append(1 + len(tiles))
    return crops, per_image


def to_pixel_values(crops, size: int = PATCH):
    """Host restatement of SiglipImageProcessor (PIL backend) with SigLIP's defaults as the reference uses it
    (EV/...anyres.py:403-404): bicubic resize to 364, x * (1/255) in float64 rounded to float32, (x - 0.5) / 0.5 in
    float32.  The reference's preprocessor_config.json is not in its tree ("parity unpinned" for the settings; the
    arithmetic is pinned by tests/golden/ref_frontend.json).  The GPU path is leopard_b200.frontend.ImageFrontend."""
    import numpy as np
    import torch
    from PIL import Image
    lut = (np.arange(256).astype(np.float64) * (1 / 255)).astype(np.float32)
    lut = (lut - np.float32(0.5)) / np.float32(0.5)
    arrs = []
    for c in crops:
        c = c.convert("RGB").resize((size, size), resample=Image.BICUBIC)
        arrs.append(lut[np.asarray(c)].transpose(2, 0, 1))
    return torch.from_numpy(np.ascontiguousarray(np.stack(arrs, 0)))
