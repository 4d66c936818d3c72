"""Generate golden vectors by EXECUTING the reference's own code (run in the authoring container only;
/root/reference does not exist on the GPU box).  Outputs are committed next to this script:

  ref_crops.json      allocate_patches / select_best_resolution / resize+pad+divide results for a table of inputs
                      (evaluations/models/llava_multiimg_siglip_anyres.py:26-162)
  ref_shuffle_proj.npz  pixel_shuffle index map and myLlavaMultiModalProjector outputs on seeded weights
                      (same file :165-192)
  ref_rope.npz        _compute_llama3_parameters and apply_rotary_pos_emb_bshd
                      (Pai-Megatron-Patch/Megatron-LM-240603/megatron/core/models/common/embeddings/
                       rotary_pos_embedding.py:48-83,197-239)

The reference module itself cannot be imported here (it needs rouge/editdistance and transformers <= 4.4x), so the
pure functions/classes are extracted from its source with `ast` and executed unmodified.
"""
import ast
import json
import math
import sys
import types
from pathlib import Path

import numpy as np
import torch
from torch import nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
EV = REF / "evaluations/models/llava_multiimg_siglip_anyres.py"


def extract(names):
    src = EV.read_text()
    tree = ast.parse(src)
    picked = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(picked) == len(names), [n.name for n in picked]
    mod = types.ModuleType("ref_extract")
    from PIL import Image
    from transformers.activations import ACT2FN
    mod.__dict__.update(dict(math=math, Image=Image, torch=torch, nn=nn, ACT2FN=ACT2FN, LlavaConfig=object))
    code = compile(ast.Module(body=picked, type_ignores=[]), str(EV), "exec")
    exec(code, mod.__dict__)
    return mod


def crops(ref):
    from PIL import Image
    size_sets = [
        [(384, 384)], [(1344, 1344)], [(896, 1152)] * 4, [(980, 980)] * 8, [(364, 364)], [(400, 1800)],
        [(3000, 200)], [(728, 364), (364, 728)], [(5000, 5000)], [(2000, 1500), (300, 300), (1000, 3000)],
        [(1456, 1456)] * 6, [(910, 910)] * 49, [(910, 910)] * 50, [(545, 545)], [(546, 546)], [(1274, 1638)] * 12,
        [(1092, 728), (728, 1092), (1456, 364)], [(640, 480)] * 3, [(1920, 1080)] * 5, [(2550, 3300)] * 20,
    ]
    rows = []
    for sizes in size_sets:
        budget = 50 - len(sizes)
        if budget > 0:
            alloc = ref.allocate_patches(list(sizes), patch_budget=budget)
            best = [ref.select_best_resolution(s, n) for s, n in zip(sizes, alloc)]
        else:
            alloc, best = None, None
        rows.append({"sizes": sizes, "budget": budget, "alloc": alloc, "best": best})
    # select_best_resolution sweep
    sweep = []
    for (w, h) in [(100, 100), (364, 364), (500, 400), (1344, 1344), (896, 1152), (3000, 500), (333, 2999), (728, 728)]:
        for n in [0, 1, 2, 3, 4, 5, 6, 9, 12, 16, 25]:
            sweep.append({"size": [w, h], "n": n, "best": ref.select_best_resolution((w, h), n)})
    # allocate with explicit budgets (exercises the scale-down path)
    alloc_cases = []
    for sizes, b in [([(1456, 1456)] * 4, 49), ([(2000, 2000), (400, 400), (1000, 500)], 10), ([(5000, 5000)], 49),
                     ([(1820, 1820), (1456, 728)], 7), ([(546, 546), (545, 545)], 48)]:
        alloc_cases.append({"sizes": sizes, "budget": b, "alloc": ref.allocate_patches(list(sizes), patch_budget=b)})
    # resize/pad/divide on synthetic images: sizes + a checksum of the pixels
    tiles = []
    rng = np.random.RandomState(0)
    for (w, h), n in [((900, 600), 6), ((300, 1000), 4), ((1344, 1344), 16), ((500, 500), 4)]:
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
        res = ref.select_best_resolution(img.size, n)
        padded = ref.resize_and_pad_image(img, res)
        pt = ref.divide_to_patches(padded, 364)
        tiles.append({"size": [w, h], "n": n, "res": res, "n_tiles": len(pt),
                      "padded_sum": int(np.asarray(padded, dtype=np.int64).sum()),
                      "tile_sums": [int(np.asarray(t, dtype=np.int64).sum()) for t in pt]})
    (OUT / "ref_crops.json").write_text(json.dumps(
        {"plans": rows, "best_sweep": sweep, "alloc_cases": alloc_cases, "tiles": tiles}, indent=1))


def shuffle_proj(ref):
    # index map: feed x[b, t, e] = b*1e6 + t*10 + e
    b, t, e = 2, 676, 3
    x = (torch.arange(b)[:, None, None] * 1_000_000 + torch.arange(t)[None, :, None] * 10 + torch.arange(e)[None, None, :]).float()
    sh = ref.pixel_shuffle(x)
    cfg = types.SimpleNamespace(vision_config=types.SimpleNamespace(hidden_size=32),
                                text_config=types.SimpleNamespace(hidden_size=48), projector_hidden_act="gelu")
    torch.manual_seed(0)
    proj = ref.myLlavaMultiModalProjector(cfg)
    xin = torch.randn(2, 676, 32)
    with torch.no_grad():
        y = proj(xin)
    np.savez_compressed(OUT / "ref_shuffle_proj.npz", shuffle_out=sh.numpy().astype(np.int64),
                        proj_in=xin.numpy(), proj_out=y.numpy(),
                        w1=proj.linear_1.weight.detach().numpy(), b1=proj.linear_1.bias.detach().numpy(),
                        w2=proj.linear_2.weight.detach().numpy(), b2=proj.linear_2.bias.detach().numpy())


def rope():
    sys.path.insert(0, str(REF / "Pai-Megatron-Patch/Megatron-LM-240603"))
    from megatron.core.models.common.embeddings import rotary_pos_embedding as R
    dim = 128
    inv = 1.0 / (500000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    inv3 = R._compute_llama3_parameters(inv)
    inv_m = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    S = 40
    pos = torch.arange(S, dtype=torch.float32)
    freqs = torch.outer(pos, inv3)
    emb = torch.cat((freqs, freqs), -1)[:, None, None, :]  # [s,1,1,d] as RotaryEmbedding.forward returns
    torch.manual_seed(1)
    t = torch.randn(S, 1, 5, dim)  # [s, b, h, d]
    out = R.apply_rotary_pos_emb_bshd(t, emb, rotary_interleaved=False)
    np.savez_compressed(OUT / "ref_rope.npz", inv_llama31=inv3.numpy(), inv_base=inv.numpy(), inv_mistral=inv_m.numpy(),
                        t=t.numpy(), out=out.numpy())


if __name__ == "__main__":
    ref = extract(["allocate_patches", "select_best_resolution", "resize_and_pad_image", "divide_to_patches",
                   "pixel_shuffle", "myLlavaMultiModalProjector"])
    crops(ref)
    shuffle_proj(ref)
    rope()
    print("wrote", sorted(p.name for p in OUT.glob("ref_*")))
