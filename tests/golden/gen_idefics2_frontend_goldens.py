"""Golden vectors for the Idefics2 image front-end, produced by EXECUTING transformers' PIL-backed Idefics2ImageProcessor — the
processor the reference builds at evaluations/models/idefics2_multiimg.py:22-25 (AutoProcessor(checkpoint, size={'longest_edge':
args.resolution, 'shortest_edge': 0}, do_image_splitting=False); transformers <= 4.4x shipped the PIL / numpy implementation as
Idefics2ImageProcessor, the installed 5.5.0 keeps it as `Idefics2ImageProcessorPil`).  The installed version treats
shortest_edge = 0 as "missing", so the goldens are made with shortest_edge = 1, which is the same rule for every image whose
resized sides are >= 1.  Run in the authoring container; writes tests/golden/ref_idefics2_frontend.json: per case the seeded image
recipe, the resized sizes, the padded shape, a sha256 of the fp32 pixel_values bytes and of the int64 pixel_attention_mask bytes.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from gen_frontend_goldens import make_image  # noqa: E402

CASES = [  # (name, [(width, height), ...], longest_edge, seed)
    ("cfg4_like_x2", [(1400, 1000), (980, 980)], 980, 11),
    ("mixed_shrink_keep", [(350, 420), (460, 100), (200, 200)], 300, 12),
    ("small_kept", [(48, 64)], 980, 13),
    ("tall", [(40, 900)], 378, 14),
]


def main():
    from PIL import Image
    from transformers.models.idefics2.image_processing_pil_idefics2 import Idefics2ImageProcessorPil
    out = []
    for name, sizes, longest, seed in CASES:
        proc = Idefics2ImageProcessorPil(size={"longest_edge": longest, "shortest_edge": 1}, do_image_splitting=False)
        imgs = [make_image(w, h, seed + i) for i, (w, h) in enumerate(sizes)]
        r = proc.preprocess([[Image.fromarray(i, "RGB") for i in imgs]], return_tensors="np")
        pv, mask = np.ascontiguousarray(r["pixel_values"][0], np.float32), np.ascontiguousarray(r["pixel_attention_mask"][0], np.int64)
        out.append({"name": name, "sizes": sizes, "longest_edge": longest, "seed": seed, "shape": list(pv.shape),
                    "resized": [[int(m[0].sum()), int(m[:, 0].sum())] for m in mask],     # (width, height) of every image
                    "pixel_values_sha256": hashlib.sha256(pv.tobytes()).hexdigest(),
                    "mask_sha256": hashlib.sha256(mask.tobytes()).hexdigest(),
                    "probe": [float(pv[0, c, 1, 2]) for c in range(3)]})
    (HERE / "ref_idefics2_frontend.json").write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {len(out)} cases")


if __name__ == "__main__":
    main()
