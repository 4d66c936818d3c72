"""-m gpu: the CUDA image front-end (lp_image_frontend through the C ABI) against the CPU oracle — bit-exact: the
path is integer / byte arithmetic plus a table lookup."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(G))
from gen_frontend_goldens import make_image  # noqa: E402
from leopard_b200.frontend import ImageFrontend, plan_request  # noqa: E402
from oracle import frontend as ofe  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def fe(native_lib):
    return ImageFrontend("cuda:0")


def oracle_request(imgs, train_order=False):
    targets = plan_request([(im.shape[1], im.shape[0]) for im in imgs])
    return np.concatenate([ofe.frontend(im, t, train_order) for im, t in zip(imgs, targets)], 0)


def test_golden_cases_bit_exact(fe):
    gold = json.loads((G / "ref_frontend.json").read_text())
    import hashlib
    for case in gold["cases"]:
        imgs = [make_image(w, h, case["seed"] * 100 + i) for i, (w, h) in enumerate(case["sizes"])]
        pv, per_image = fe(imgs)
        assert per_image == case["per_image"] and pv.dtype == torch.float32
        got = [hashlib.sha256(c.contiguous().cpu().numpy().tobytes()).hexdigest() for c in pv]
        assert got == case["sha256"], case["name"]   # == the reference functions + PIL-backed SiglipImageProcessor


@pytest.mark.parametrize("sizes", [
    [(896, 1152)] * 4,                   # cfg3 (BASELINE.json configs[2]): 28 crops
    [(1344, 1344)],                      # cfg2: 17 crops
    [(980, 980)] * 8,                    # cfg4 sizes through the LLaVA planner
    [(37, 991), (991, 37)],              # extreme aspect ratios (one-sample-wide intermediate rows)
    [(364, 364)], [(365, 363)],          # identity / near-identity passes
    [(5, 5)],                            # tiny image, pure upscaling
    [(3001, 700)],                       # odd widths: unaligned source rows, wide downscale
    [(9, 2000)], [(5, 900), (41, 4101)],  # > 100x taller than wide: Pillow resamples vertically FIRST
])
def test_random_images_bit_exact_full_sizes(fe, sizes):
    rng = np.random.default_rng(len(sizes) * 1000 + sizes[0][0])
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in sizes]
    ref = oracle_request(imgs)
    pv, per_image = fe(imgs)
    assert pv.shape == ref.shape and sum(per_image) == ref.shape[0]
    assert np.array_equal(pv.cpu().numpy(), ref)


def test_train_order_bf16_and_input_kinds(fe):
    from PIL import Image
    imgs = [make_image(900, 300, 1), make_image(250, 1000, 2)]
    ref = oracle_request(imgs, train_order=True)
    pv, _ = fe(imgs, train_order=True)
    assert np.array_equal(pv.cpu().numpy(), ref)
    # bf16 output == fp32 output rounded once (what lp_im2col_patch14 would do with fp32 pixels)
    pb, _ = fe(imgs, dtype=torch.bfloat16, train_order=True)
    assert torch.equal(pb, pv.to(torch.bfloat16))
    # PIL images and device tensors are accepted too
    p2, _ = fe([Image.fromarray(im, "RGB") for im in imgs], train_order=True)
    p3, _ = fe([torch.from_numpy(im).cuda() for im in imgs], train_order=True)
    assert torch.equal(p2, pv) and torch.equal(p3, pv)


def test_properties_constant_image_and_determinism(fe):
    """size-independent properties: a constant image stays constant through any resize (taps sum to one up to the
    22-bit quantisation, which the rounding absorbs), padding is exactly -1, results are deterministic."""
    img = np.full((1152, 896, 3), 200, np.uint8)
    pv, per = fe([img])
    lut = ofe.normalize_lut()
    assert per == [7]
    assert bool((pv[0] == float(lut[200])).all())
    tiles = pv[1:].cpu().numpy().reshape(3, 2, 3, 364, 364)  # rows x cols of the 728 x 1092 canvas
    canvas = np.concatenate([np.concatenate(list(r), -1) for r in tiles], -2)  # [3, 1092, 728]
    assert (canvas[:, :78] == -1.0).all() and (canvas[:, 78 + 936:] == -1.0).all()
    assert (canvas[:, 78:78 + 936] == lut[200]).all()
    pv2, _ = fe([img])
    assert torch.equal(pv, pv2)


def test_argument_errors(fe):
    import ctypes as C
    from leopard_b200 import _lib
    L = _lib.lib()
    img = torch.zeros((10, 10, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((1, 3, 364, 364), device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    m = (C.c_float * 3)(0.5, 0.5, 0.5)
    rc = L.lp_image_frontend(_lib.ptr(img), 10, 10, 500, 364, _lib.ptr(out), _lib.ptr(out), 1, 1 / 255, m, m,
                             _lib.ptr(ws), ws.numel(), None)
    assert rc == -1 and b"364-pixel tiles" in L.lp_last_error()
    rc = L.lp_image_frontend(_lib.ptr(img), 10, 10, 0, 0, _lib.ptr(out), None, 1, 1 / 255, m, m, _lib.ptr(ws), 16, None)
    assert rc == -1 and b"workspace too small" in L.lp_last_error()
    with pytest.raises(ValueError):
        fe([np.zeros((4, 4), np.uint8)])


def test_idefics2_processor_on_the_gpu_is_bit_exact():
    """Idefics2Frontend (lp_image_resize: BILINEAR resize by the longest-edge rule + rescale + normalise + zero padding + pixel mask)
    against goldens made by executing transformers' PIL-backed Idefics2ImageProcessor, and against the oracle on random images."""
    import hashlib
    import json
    import sys
    from pathlib import Path
    G = Path(__file__).resolve().parent / "golden"
    sys.path.insert(0, str(G))
    from gen_frontend_goldens import make_image
    from leopard_b200.frontend import Idefics2Frontend
    from oracle import frontend as ofe
    for case in json.loads((G / "ref_idefics2_frontend.json").read_text()):
        imgs = [make_image(w, h, case["seed"] + i) for i, (w, h) in enumerate(case["sizes"])]
        pv, mask = Idefics2Frontend("cuda:0", longest_edge=case["longest_edge"])(imgs)
        assert list(pv.shape) == [1] + case["shape"] and list(mask.shape) == [1, case["shape"][0]] + case["shape"][2:]
        assert hashlib.sha256(pv[0].cpu().numpy().tobytes()).hexdigest() == case["pixel_values_sha256"], case["name"]
        assert hashlib.sha256(mask[0].cpu().numpy().tobytes()).hexdigest() == case["mask_sha256"], case["name"]
    rng = np.random.default_rng(9)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(1300, 977), (640, 1500), (2100, 16)]]
    fe2 = Idefics2Frontend("cuda:0", longest_edge=980)
    pv, mask = fe2(imgs)
    want_pv, want_mask = ofe.idefics2_preprocess(imgs, 980, 0)
    assert np.array_equal(pv[0].cpu().numpy(), want_pv) and np.array_equal(mask[0].cpu().numpy(), want_mask)
    pv16, _ = fe2(imgs, dtype=torch.bfloat16)                      # bf16 output == fp32 output rounded once
    assert torch.equal(pv16, pv.to(torch.bfloat16))
