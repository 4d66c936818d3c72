"""CPU: the image front-end oracle (oracle/frontend.py) pinned against (a) goldens produced by executing the
reference's own resize_and_pad_image / divide_to_patches + the PIL-backed SiglipImageProcessor
(tests/golden/gen_frontend_goldens.py), (b) the Pillow installed here, bit for bit, and the host planning in
leopard_b200/frontend.py against the same goldens."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

G = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(G))
from gen_frontend_goldens import make_image  # noqa: E402  (pure numpy recipe; does not touch /root/reference)
from oracle import frontend as fe  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return json.loads((G / "ref_frontend.json").read_text())


def test_oracle_matches_reference_executed_goldens(gold):
    from leopard_b200.frontend import plan_request
    for case in gold["cases"]:
        sizes = [tuple(s) for s in case["sizes"]]
        targets = plan_request(sizes)
        assert [None if t is None else list(t) for t in targets] == case["best"], case["name"]
        crops = []
        for i, ((w, h), tgt) in enumerate(zip(sizes, targets)):
            crops.append(fe.frontend(make_image(w, h, case["seed"] * 100 + i), tgt))
        assert [c.shape[0] for c in crops] == case["per_image"]
        pv = np.concatenate(crops, 0)
        assert pv.dtype == np.float32
        got = [hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest() for c in pv]
        assert got == case["sha256"], case["name"]
        for c, probe in zip(pv, case["probe"]):
            assert [float(c[0, 0, 0]), float(c[1, 181, 200]), float(c[2, 363, 363])] == probe[:3]


@pytest.mark.parametrize("w,h,ow,oh", [(500, 700, 364, 364), (100, 80, 364, 364), (896, 1152, 728, 936),
                                       (37, 991, 364, 364), (300, 300, 301, 299), (364, 364, 364, 364),
                                       (1000, 10, 10, 1000), (1277, 958, 617, 329), (3, 2, 364, 364), (9, 2000, 364, 364),
                                       (728, 364, 728, 364), (2200, 1700, 2184, 1688), (1, 1, 5, 7)])
def test_resize_is_bit_identical_to_pillow(w, h, ow, oh):
    from PIL import Image
    rng = np.random.default_rng(w * 7 + h)
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(fe.pil_resize_bicubic(a, ow, oh), np.asarray(Image.fromarray(a).resize((ow, oh))))


def test_normalise_table_and_pad_geometry():
    lut = fe.normalize_lut()
    assert lut.dtype == np.float32 and lut[0] == -1.0 and lut[255] == 1.0
    # the float64-multiply-then-round rule differs from a float32 divide for some levels: pin the rule itself
    assert np.array_equal(lut, ((np.arange(256).astype(np.float64) * (1 / 255)).astype(np.float32) - 0.5) / 0.5)
    assert fe.pad_geometry(896, 1152, (728, 1092)) == (728, 936, 0, 78)
    assert fe.pad_geometry(250, 1000, (364, 1092)) == (273, 1092, 45, 0)
    assert fe.pad_geometry(100, 100, None) is None


def test_host_shim_to_pixel_values_follows_the_processor_rule():
    """leopard_b200.preprocess.to_pixel_values (host path used when images are already PIL crops) == oracle."""
    from PIL import Image
    from leopard_b200.preprocess import to_pixel_values
    img = make_image(500, 420, 9)
    got = to_pixel_values([Image.fromarray(img, "RGB")]).numpy()[0]
    assert np.array_equal(got, fe.siglip_preprocess(img))


# ---- Idefics2 processor (evaluations/models/idefics2_multiimg.py:22-25) ---------------------------------------------------------
@pytest.mark.parametrize("w,h,ow,oh", [(500, 700, 250, 350), (100, 80, 364, 291), (1400, 1000, 980, 700), (37, 991, 14, 378),
                                       (16, 2100, 3, 400), (300, 300, 300, 300), (1277, 958, 617, 329), (3, 2, 64, 43)])
def test_bilinear_resize_is_bit_identical_to_pillow(w, h, ow, oh):
    from PIL import Image
    rng = np.random.default_rng(w * 5 + h)
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(fe.pil_resize(a, ow, oh, "bilinear"), np.asarray(Image.fromarray(a).resize((ow, oh), Image.BILINEAR)))


def test_idefics2_oracle_matches_the_installed_pil_processor_and_its_goldens():
    """oracle/frontend.py::idefics2_preprocess == transformers' PIL-backed Idefics2ImageProcessor, bit for bit: against committed
    goldens made by executing the processor (tests/golden/gen_idefics2_frontend_goldens.py) and against the processor itself;
    the product's host planning (leopard_b200.frontend.idefics2_resize_size) gives the processor's sizes."""
    from leopard_b200.frontend import idefics2_resize_size
    gold = json.loads((G / "ref_idefics2_frontend.json").read_text())
    for case in gold:
        imgs = [make_image(w, h, case["seed"] + i) for i, (w, h) in enumerate(case["sizes"])]
        pv, mask = fe.idefics2_preprocess(imgs, case["longest_edge"], 0)
        assert list(pv.shape) == case["shape"] and pv.dtype == np.float32 and mask.dtype == np.int64
        assert hashlib.sha256(pv.tobytes()).hexdigest() == case["pixel_values_sha256"], case["name"]
        assert hashlib.sha256(mask.tobytes()).hexdigest() == case["mask_sha256"], case["name"]
        assert [list(idefics2_resize_size(w, h, case["longest_edge"], 0)) for w, h in case["sizes"]] == case["resized"]
    from PIL import Image
    from transformers.models.idefics2.image_processing_pil_idefics2 import Idefics2ImageProcessorPil
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(420, 350), (100, 460), (200, 200), (301, 299)]]
    r = Idefics2ImageProcessorPil(size={"longest_edge": 300, "shortest_edge": 1}, do_image_splitting=False).preprocess(
        [[Image.fromarray(i, "RGB") for i in imgs]], return_tensors="np")
    pv, mask = fe.idefics2_preprocess(imgs, 300, 0)
    assert np.array_equal(pv, r["pixel_values"][0]) and np.array_equal(mask, r["pixel_attention_mask"][0])
