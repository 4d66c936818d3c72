"""Golden vectors for the image front-end, produced by EXECUTING the reference's own functions
(evaluations/models/llava_multiimg_siglip_anyres.py: allocate_patches, select_best_resolution, resize_and_pad_image,
divide_to_patches — extracted with `ast`, run unmodified on Pillow) followed by the PIL-backed SiglipImageProcessor of
the installed transformers (`SiglipImageProcessorPil`, the implementation transformers <= 4.4x shipped as
SiglipImageProcessor, which is what the reference calls at :378,403-404) with SigLIP's default preprocessor settings
at size 364 (the reference's preprocessor_config.json is not in its tree).

Run in the authoring container only (needs /root/reference).  Writes tests/golden/ref_frontend.json: per case the
seeded image recipe, the planned resolutions, crop counts, and for every crop a sha256 of the fp32 pixel_values bytes
plus a few probe values, in the eval order `[origin] + patches` (:395-399).
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from gen_reference_goldens import extract  # noqa: E402

CASES = [  # (name, [(width, height), ...], seed)
    ("cfg1_384", [(384, 384)], 1),
    ("cfg3_slide_x2", [(896, 1152)] * 2, 2),
    ("wide_and_tall", [(900, 300), (250, 1000)], 3),
    ("upscale_small", [(120, 90)], 4),
    ("big_downscale", [(2200, 1700)], 5),
    ("exact_tiles", [(728, 364)], 6),
]


def make_image(w, h, seed):
    """smooth gradients + noise + hard edges, uint8 RGB (deterministic from numpy's PCG64)"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) % 256)], -1).astype(np.int64)
    noise = rng.integers(-40, 41, (h, w, 3))
    img = np.clip(base + noise, 0, 255).astype(np.uint8)
    img[h // 3: h // 3 + max(h // 10, 1), :, :] = 255   # white bar
    img[:, w // 2: w // 2 + max(w // 16, 1), :] = 0     # black bar
    return img


def main():
    from PIL import Image
    from transformers.models.siglip.image_processing_pil_siglip import SiglipImageProcessorPil
    ref = extract(["allocate_patches", "select_best_resolution", "resize_and_pad_image", "divide_to_patches"])
    proc = SiglipImageProcessorPil(size={"height": 364, "width": 364})
    out = []
    for name, sizes, seed in CASES:
        images = [Image.fromarray(make_image(w, h, seed * 100 + i), "RGB") for i, (w, h) in enumerate(sizes)]
        budget = 50 - len(images)
        alloc = ref.allocate_patches([im.size for im in images], patch_budget=budget)
        best = [ref.select_best_resolution(im.size, n) for im, n in zip(images, alloc)]
        padded = [ref.resize_and_pad_image(im, r) for im, r in zip(images, best)]
        patches = [ref.divide_to_patches(p, 364) if p is not None else [] for p in padded]
        crops = []
        for origin, pt in zip(images, patches):
            crops += [origin] + pt
        pv = np.concatenate([proc.preprocess(c, return_tensors="np")["pixel_values"] for c in crops], 0)
        assert pv.dtype == np.float32 and pv.shape[1:] == (3, 364, 364)
        out.append({
            "name": name, "sizes": sizes, "seed": seed, "best": best, "per_image": [1 + len(p) for p in patches],
            "sha256": [hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest() for c in pv],
            "probe": [[float(c[0, 0, 0]), float(c[1, 181, 200]), float(c[2, 363, 363]), float(c.sum(dtype=np.float64))]
                      for c in pv],
        })
        print(name, best, [1 + len(p) for p in patches])
    (HERE / "ref_frontend.json").write_text(json.dumps({"pillow": __import__("PIL").__version__, "cases": out}, indent=1))


if __name__ == "__main__":
    main()
