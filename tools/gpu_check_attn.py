"""GPU-side check + timing of lp_attention_fwd against torch fp32 SDPA (run under gpurun)."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from leopard_b200 import ops  # noqa: E402

CASES = {
    # name: (B, T, heads, kv_heads, dh, hs, causal)
    "dec_t128": (1, 128, 4, 1, 128, 128, False),
    "dec_t128_c": (1, 128, 4, 1, 128, 128, True),
    "dec_t256": (1, 256, 8, 2, 128, 128, False),
    "dec_t300_c": (2, 300, 8, 2, 128, 128, True),
    "vit_t128": (1, 128, 2, 2, 72, 80, False),
    "vit_t676": (3, 676, 16, 16, 72, 80, False),
    "dec_2937_c": (1, 2937, 32, 8, 128, 128, True),
    "dec_4796_c": (1, 4796, 32, 8, 128, 128, True),
    "vit_28": (28, 676, 16, 16, 72, 80, False),
    "vit_17": (17, 676, 16, 16, 72, 80, False),
    "vit_5x420": (5, 420, 16, 16, 72, 80, False),
    "vit_2x1024": (2, 1024, 16, 16, 72, 80, False),
    "vit_4900": (8, 4900, 16, 16, 72, 80, False),      # cfg4: Idefics2 NaViT tower, 8 images of 980 x 980
}


def make_qkv(B, T, H, HK, dh, hs, g):
    wq, wk = H * hs, HK * hs
    ld = wq + 2 * wk
    qkv = torch.zeros(B * T, ld, device="cuda", dtype=torch.bfloat16)
    q = torch.randn(B, T, H, dh, device="cuda", generator=g)
    k = torch.randn(B, T, HK, dh, device="cuda", generator=g)
    v = torch.randn(B, T, HK, dh, device="cuda", generator=g)
    # make a few rows "peaky" so the lazy-rescale path and large maxima are exercised
    q[:, ::7] *= 3.0
    qv = qkv.view(B, T, ld)
    qv[:, :, :wq].view(B, T, H, hs)[..., :dh] = q.to(torch.bfloat16)
    qv[:, :, wq:wq + wk].view(B, T, HK, hs)[..., :dh] = k.to(torch.bfloat16)
    qv[:, :, wq + wk:].view(B, T, HK, hs)[..., :dh] = v.to(torch.bfloat16)
    if hs > dh:  # v pad columns may hold anything; put garbage there to prove they are ignored
        qv[:, :, wq + wk:].view(B, T, HK, hs)[..., dh:] = 7.0
    return qkv, (0, wq, wq + wk)


def reference(qkv, offs, B, T, H, HK, dh, hs, causal):
    wq, wk = H * hs, HK * hs
    qv = qkv.view(B, T, -1).float()
    q = qv[:, :, :wq].view(B, T, H, hs)[..., :dh].transpose(1, 2)
    k = qv[:, :, wq:wq + wk].view(B, T, HK, hs)[..., :dh].transpose(1, 2)
    v = qv[:, :, wq + wk:].view(B, T, HK, hs)[..., :dh].transpose(1, 2)
    rep = H // HK
    k = k.repeat_interleave(rep, dim=1)
    v = v.repeat_interleave(rep, dim=1)
    outs = []
    for h0 in range(0, H, 8):  # chunk heads to bound the fp32 score matrix
        s = (q[:, h0:h0 + 8] @ k[:, h0:h0 + 8].transpose(-1, -2)) * dh ** -0.5
        if causal:
            s = s.masked_fill(torch.ones(T, T, device=s.device, dtype=torch.bool).triu(1), float("-inf"))
        outs.append(torch.softmax(s, -1) @ v[:, h0:h0 + 8])
    o = torch.cat(outs, 1)
    return o.transpose(1, 2).reshape(B * T, H * dh)


def run_case(name, variant, do_time):
    B, T, H, HK, dh, hs, causal = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(1234)
    qkv, offs = make_qkv(B, T, H, HK, dh, hs, g)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = reference(qkv, offs, B, T, H, HK, dh, hs, causal)
    out = ops.attention(qkv, B, T, H, HK, dh, hs, offs[0], offs[1], offs[2], causal, variant=variant)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    tol = 2e-2 * ref.abs().clamp(min=0.5)
    bad = err > tol
    nbad = int(bad.sum())
    rec = {"case": name, "variant": variant, "max_err": float(err.max()), "ref_absmax": float(ref.abs().max()),
           "bad": nbad, "ok": nbad == 0 and bool(torch.isfinite(out.float()).all())}
    if variant & (128 | 256):  # the two-stream / persistent kernels do v3's arithmetic row by row: results must be bit-identical
        out3 = ops.attention(qkv, B, T, H, HK, dh, hs, offs[0], offs[1], offs[2], causal, variant=(variant & ~(128 | 256)) | 64 | 16)
        rec["same_as_v3"] = bool(torch.equal(out, out3))
    if not rec["ok"]:
        eb = err.view(B, T, H, dh)
        rec["err_by_head"] = [round(float(x), 3) for x in eb.amax(dim=(0, 1, 3))[:16]]
        rec["err_by_rowblk"] = [round(float(eb[:, i:i + 32].max()), 3) for i in range(0, min(T, 256), 32)]
        rec["err_by_col16"] = [round(float(eb[..., i:i + 16].max()), 3) for i in range(0, dh, 16)]
        rec["nan"] = int(torch.isnan(out.float()).sum())
        rec["sample_got"] = [round(float(x), 3) for x in out[0, :6]]
        rec["sample_ref"] = [round(float(x), 3) for x in ref[0, :6]]
    if do_time and rec["ok"]:
        for _ in range(3):
            ops.attention(qkv, B, T, H, HK, dh, hs, offs[0], offs[1], offs[2], causal, out=out, variant=variant)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.attention(qkv, B, T, H, HK, dh, hs, offs[0], offs[1], offs[2], causal, out=out, variant=variant)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 4.0 * B * H * T * T * dh * (0.5 if causal else 1.0)
        rec["ms"] = ms
        rec["tflops"] = flops / ms / 1e9
    print(json.dumps(rec), flush=True)
    return rec["ok"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", action="append")
    ap.add_argument("--variant", type=int, action="append")
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    ok = True
    for v in (args.variant or [0]):
        for n in (args.case or list(CASES)):
            try:
                ok &= run_case(n, v, args.time)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"case": n, "variant": v, "ok": False, "exception": repr(e)[:600]}), flush=True)
                return 1
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
