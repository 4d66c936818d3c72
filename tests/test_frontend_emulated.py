"""The whole image front-end (leopard_b200/csrc/lp_frontend.cu: the kernels AND the host function that sequences them — pad
geometry, Pillow's vertical-first rule, workspace carving) executed on the CPU: launches rewritten by tests/emu/launch_rewrite.py,
CUDA runtime names from tests/emu/host_emu.h, fp64 coefficient arithmetic compiled with -ffp-contract=off.  Result must be
BIT-identical to the Pillow-pinned oracle (oracle/frontend.py), as it is on the B200 (tests/test_gpu_frontend.py)."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import frontend as OF

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu_fe")
    sys.path.insert(0, str(ROOT / "tests/emu"))
    from launch_rewrite import rewrite_launches
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// numerics helpers"))
    (out / "elem_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    src = (ROOT / "leopard_b200/csrc/lp_frontend.cu").read_text()
    for inc in ('#include "lp_common.cuh"\n', '#include "lp_host.h"\n', '#include "../../include/leopard_b200.h"\n'):
        assert inc in src
        src = src.replace(inc, "")
    decl = "  extern __shared__ uint4 fe_smem[];\n"           # -> the static buffer of tests/emu/frontend_emu.cpp
    assert decl in src
    src = rewrite_launches(src.replace(decl, ""))
    assert "<<<" not in src
    (out / "frontend_src.inc").write_text(src)
    lib = out / "libfe.so"
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I",
                        str(ROOT / "tests/emu"), str(ROOT / "tests/emu/frontend_emu.cpp"), "-o", str(lib)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(str(lib))
    L.lp_image_frontend_workspace_bytes.restype = C.c_size_t
    return L


@pytest.mark.parametrize("h,w,target", [
    (50, 70, None),            # up-scaling in both directions, original crop only
    (400, 3, None),            # h > 100 w and shrinking vertically: Pillow's vertical-first order
    (90, 61, (728, 364)),      # aspect-preserving resize, centred black padding, two tiles + the original
])
def test_frontend_source_is_bit_exact_under_emulation(emu, h, w, target):
    img = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = OF.frontend(img, target)
    n = want.shape[0]
    buf = torch.from_numpy(img.copy())
    out = torch.full((n, 3, 364, 364), float("nan"))
    tw, th = target if target else (0, 0)
    ws = torch.empty(emu.lp_image_frontend_workspace_bytes(h, w, tw, th), dtype=torch.uint8)
    mean, std = (C.c_float * 3)(0.5, 0.5, 0.5), (C.c_float * 3)(0.5, 0.5, 0.5)
    rc = emu.lp_image_frontend(C.c_void_p(buf.data_ptr()), h, w, tw, th, C.c_void_p(out[0].data_ptr()),
                               C.c_void_p(out[1].data_ptr()) if n > 1 else None, 1, C.c_double(1 / 255), mean, std,
                               C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), None)
    assert rc == 0
    assert np.array_equal(out.numpy(), want)


@pytest.mark.parametrize("sizes,longest", [
    ([(420, 350), (100, 460), (200, 200)], 300),     # two images shrink (one by height, one by width), one is kept: zero padding to the largest
    ([(64, 48)], 980),                               # smaller than the limit: identity resize, no padding
    ([(2100, 16)], 400),                             # > 100x taller than wide and shrinking: Pillow's vertical-first order, bilinear
])
def test_idefics2_processor_source_is_bit_exact_under_emulation(emu, sizes, longest):
    """lp_image_resize (BILINEAR Image.resize + rescale + normalise + zero padding + pixel mask) against the oracle's restatement of
    transformers' PIL-backed Idefics2ImageProcessor, which tests/test_frontend_oracle.py pins to the installed processor itself."""
    rng = np.random.default_rng(len(sizes) + longest)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    want_pv, want_mask = OF.idefics2_preprocess(imgs, longest, 0)
    n, _, hm, wm = want_pv.shape
    pv = torch.full((n, 3, hm, wm), float("nan"))
    mask = torch.full((n, hm, wm), -1, dtype=torch.long)
    mean, std = (C.c_float * 3)(0.5, 0.5, 0.5), (C.c_float * 3)(0.5, 0.5, 0.5)
    emu.lp_image_resize_workspace_bytes.restype = C.c_size_t
    for i, img in enumerate(imgs):
        h, w = img.shape[:2]
        ow, oh = OF.idefics2_resize_size(w, h, longest, 0)
        buf = torch.from_numpy(img.copy())
        ws = torch.empty(emu.lp_image_resize_workspace_bytes(h, w, ow, oh, 1), dtype=torch.uint8)
        rc = emu.lp_image_resize(C.c_void_p(buf.data_ptr()), h, w, ow, oh, 1, wm, hm, C.c_float(0.0), C.c_void_p(pv[i].data_ptr()), 1,
                                 C.c_void_p(mask[i].data_ptr()), C.c_double(1 / 255), mean, std, C.c_void_p(ws.data_ptr()),
                                 C.c_size_t(ws.numel()), None)
        assert rc == 0
    assert np.array_equal(pv.numpy(), want_pv) and np.array_equal(mask.numpy(), want_mask)
