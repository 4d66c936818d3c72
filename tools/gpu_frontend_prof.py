"""Image front-end kernels in isolation for ncu: the cfg3 request's four 896 x 1152 images, three times."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from leopard_b200.frontend import ImageFrontend  # noqa: E402

rng = np.random.default_rng(7)
imgs = [torch.from_numpy(rng.integers(0, 256, (1152, 896, 3), dtype=np.uint8)).cuda() for _ in range(4)]
fe = ImageFrontend("cuda:0")
for _ in range(3):
    pv, per = fe(imgs, dtype=torch.bfloat16)
torch.cuda.synchronize()
print("crops", per, tuple(pv.shape))
