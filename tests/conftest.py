import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def native_lib():
    """The C-ABI library, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from leopard_b200 import _lib, build
    if not _lib.LIB_PATH.exists():
        build.build_native()
    return _lib.lib()


@pytest.fixture(scope="session")
def seeded_full_engine(native_lib):
    """Full-size Leopard-LLaVA (27 + 32 layers) on the seeded synthetic checkpoint (seed 1234) — the weights the
    committed full-depth oracle fixtures (tests/golden/fullsize_goldens.npz) and the cfg1 full-depth test were computed with.
    Built once per session: materialising 8.5 G parameters on the host takes a minute or two."""
    from leopard_b200.config import leopard_llava
    from leopard_b200.engine import PrefillEngine
    from leopard_b200.native_weights import load_hf
    from leopard_b200.weights import SyntheticCheckpoint
    cfg = leopard_llava()
    ck = SyntheticCheckpoint(cfg, seed=1234)
    return cfg, ck, PrefillEngine(cfg, load_hf(cfg, lambda n: ck[n], "cuda"))
