"""GPU-side check + timing of lp_gemm_bf16 against torch fp32 matmul (run under gpurun).

Usage: python tools/gpu_check_gemm.py [--case NAME] [--list] [--time]
Each case prints one JSON line; failures print an error-pattern summary to help debug descriptor/layout bugs.
"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from leopard_b200 import ops  # noqa: E402

CASES = {
    # name: (M, N, K, bias, act, residual, res_row_mod, out_fp32, block_n)
    "tile1": (128, 256, 64, False, 0, False, 0, False, 256),
    "tile1_k4": (128, 256, 256, False, 0, False, 0, False, 256),
    "tile1_k9": (128, 256, 576, False, 0, False, 0, False, 256),
    "tile1_bn128": (128, 128, 576, False, 0, False, 0, False, 128),
    "ragged": (300, 520, 1160, True, 0, False, 0, False, 256),
    "ragged128": (300, 520, 1160, True, 0, True, 0, False, 128),
    "fp32out": (200, 384, 512, True, 0, False, 0, True, 256),
    "vit_qkv": (676, 3840, 1152, True, 0, False, 0, False, 0),
    "vit_fc1": (676, 4304, 1152, True, 1, False, 0, False, 0),
    "vit_fc2": (676, 1152, 4304, True, 0, True, 0, False, 0),
    "vit_patch": (676 * 2, 1152, 592, True, 0, True, 676, False, 0),
    "proj1_erf": (169 * 3, 4096, 4608, True, 2, False, 0, False, 0),
    "dec_qkv": (2937, 6144, 4096, False, 0, False, 0, False, 0),
    "dec_o": (2937, 4096, 4096, False, 0, True, 0, False, 0),
    "dec_gateup": (2937, 28672, 4096, False, 3, False, 0, False, 0),
    "dec_down": (2937, 4096, 14336, False, 0, True, 0, False, 0),
    "big_qkv": (4796, 6144, 4096, False, 0, False, 0, False, 256),
    "big_gateup": (4796, 28672, 4096, False, 3, False, 0, False, 256),
    "big_down": (4796, 4096, 14336, False, 0, True, 0, False, 256),
    "big_down128": (4796, 4096, 14336, False, 0, True, 0, False, 128),
    "vitbig_fc1": (18928, 4304, 1152, True, 1, False, 0, False, 256),
    "vitbig_fc2": (18928, 1152, 4304, True, 0, True, 0, False, 0),
    "lm_head": (64, 128256, 4096, False, 0, False, 0, True, 0),
    # --- bench shapes (cfg3): ViT M = 28*676, decoder M = 4796; block_n sweeps
    "v_patch": (18928, 1152, 592, True, 0, True, 676, False, 0),
    "v_qkv": (18928, 3840, 1152, True, 0, False, 0, False, 0),
    "v_out": (18928, 1152, 1152, True, 0, True, 0, False, 0),
    "v_out_256": (18928, 1152, 1152, True, 0, True, 0, False, 256),
    "v_fc1": (18928, 4304, 1152, True, 1, False, 0, False, 0),
    "v_fc2": (18928, 1152, 4304, True, 0, True, 0, False, 0),
    "v_fc2_256": (18928, 1152, 4304, True, 0, True, 0, False, 256),
    "p_1": (4732, 4096, 4608, True, 2, False, 0, False, 0),
    "d_qkv": (4796, 6144, 4096, False, 0, False, 0, False, 0),
    "d_qkv_192": (4796, 6144, 4096, False, 0, False, 0, False, 192),
    "d_o": (4796, 4096, 4096, False, 0, True, 0, False, 0),
    "d_o_256": (4796, 4096, 4096, False, 0, True, 0, False, 256),
    "d_o_160": (4796, 4096, 4096, False, 0, True, 0, False, 160),
    "d_o_224": (4796, 4096, 4096, False, 0, True, 0, False, 224),
    "d_o_192": (4796, 4096, 4096, False, 0, True, 0, False, 192),
    "d_down_224": (4796, 4096, 14336, False, 0, True, 0, False, 224),
    "d_down_192": (4796, 4096, 14336, False, 0, True, 0, False, 192),
    "d_qkv_224": (4796, 6144, 4096, False, 0, False, 0, False, 224),
    "d_qkv_256": (4796, 6144, 4096, False, 0, False, 0, False, 256),
    "v_qkv_224": (18928, 3840, 1152, True, 0, False, 0, False, 224),
    "v_qkv_192": (18928, 3840, 1152, True, 0, False, 0, False, 192),
    "v_fc1_224": (18928, 4304, 1152, True, 1, False, 0, False, 224),
    "v_fc1_192": (18928, 4304, 1152, True, 1, False, 0, False, 192),
    "v_fc2_192": (18928, 1152, 4304, True, 0, True, 0, False, 192),
    "v_fc2_128": (18928, 1152, 4304, True, 0, True, 0, False, 128),
    "v_out_192": (18928, 1152, 1152, True, 0, True, 0, False, 192),
    "v_out_128": (18928, 1152, 1152, True, 0, True, 0, False, 128),
    "d_gateup": (4796, 28672, 4096, False, 3, False, 0, False, 0),
    "d_gateup_192": (4796, 28672, 4096, False, 3, False, 0, False, 192),
    "d_down": (4796, 4096, 14336, False, 0, True, 0, False, 0),
    "d_down_256": (4796, 4096, 14336, False, 0, True, 0, False, 256),
    "d_down_160": (4796, 4096, 14336, False, 0, True, 0, False, 160),
    "v_out_b": (18928, 1152, 1152, True, 0, False, 0, False, 0),
    "v_out_n": (18928, 1152, 1152, False, 0, False, 0, False, 0),
    "v_out_n256": (18928, 1152, 1152, False, 0, False, 0, False, 256),
    "v_k2304_n": (18928, 1152, 2304, False, 0, False, 0, False, 0),
    "r224": (300, 520, 1160, True, 0, True, 0, False, 224),
    "r192": (300, 520, 1160, True, 1, True, 0, False, 192),
    "r160": (300, 520, 1160, True, 2, True, 0, False, 160),
    "sw192": (300, 1152, 1160, True, 3, False, 0, False, 192),
    "sw128": (300, 1152, 1160, True, 3, True, 0, False, 128),
}


def reference(a, w, bias, act, residual, res_row_mod):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y, approximate="tanh")
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    elif act == 3:
        n = y.shape[1]
        yy = y.view(y.shape[0], n // 64, 2, 32)
        y = (torch.nn.functional.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(y.shape[0], n // 2)
    if residual is not None:
        if res_row_mod:
            idx = torch.arange(y.shape[0], device=y.device) % res_row_mod
            y = y + residual.float()[idx]
        else:
            y = y + residual.float()
    return y


def run_case(name, do_time):
    M, N, K, has_bias, act, has_res, rmod, out_fp32, bn = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 2**31)
    ldk = (K + 7) // 8 * 8
    a = (torch.randn(M, ldk, device="cuda", generator=g)).to(torch.bfloat16)[:, :K]
    w = (torch.randn(N, ldk, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)[:, :K]
    bias = torch.randn(N, device="cuda", generator=g) * 0.5 if has_bias else None
    n_out = N // 2 if act == 3 else N
    rrows = rmod if rmod else M
    residual = torch.randn(rrows, n_out, device="cuda", generator=g).to(torch.bfloat16) if has_res else None
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = reference(a, w, bias, act, residual, rmod)
    out = ops.gemm(a, w, bias=bias, residual=residual, act=act, out_fp32=out_fp32, res_row_mod=rmod, block_n=bn)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    tol = 1e-2 * ref.abs().clamp(min=1.0) if not out_fp32 else 2e-3 * ref.abs().clamp(min=1.0)
    bad = err > tol
    nbad = int(bad.sum())
    rec = {"case": name, "M": M, "N": N, "K": K, "max_err": float(err.max()), "ref_absmax": float(ref.abs().max()),
           "bad": nbad, "ok": nbad == 0 and bool(torch.isfinite(out.float()).all())}
    if nbad:
        # error pattern by 32-row / 32-col blocks
        rb = bad.float().view(-1, bad.shape[1]).sum(1)
        rows_bad = torch.nonzero(rb).flatten()[:16].tolist()
        cb = bad.float().sum(0)
        cols_bad = torch.nonzero(cb).flatten()[:16].tolist()
        rec["bad_rows_head"] = rows_bad
        rec["bad_cols_head"] = cols_bad
        rec["bad_frac"] = nbad / bad.numel()
        i = torch.nonzero(bad)[0].tolist()
        rec["first_bad"] = {"idx": i, "got": float(out[i[0], i[1]]), "ref": float(ref[i[0], i[1]])}
    if do_time and rec["ok"]:
        for _ in range(3):
            ops.gemm(a, w, bias=bias, residual=residual, act=act, out=out, res_row_mod=rmod, block_n=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.gemm(a, w, bias=bias, residual=residual, act=act, out=out, res_row_mod=rmod, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rec["ms"] = ms
        rec["tflops"] = 2.0 * M * N * K / ms / 1e9
        t0 = time.time()
        for _ in range(iters):
            (a @ w.t())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            (a @ w.t())
        e1.record()
        torch.cuda.synchronize()
        rec["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / iters) / 1e9
    print(json.dumps(rec), flush=True)
    return rec["ok"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", action="append")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--pair", action="store_true", help="CTA-pair (cta_group::2) GEMM mode")
    ap.add_argument("--single", action="store_true", help="one CTA per tile")
    ap.add_argument("--quad", action="store_true", help="4-CTA clusters with A multicast (lp_gemm4.cu, mode 2)")
    ap.add_argument("--streamk", type=int, default=1, help="stream-K tail on / off (lp_gemm_set_streamk)")
    ap.add_argument("--balance", type=int, default=1, help="row-major raster for ragged N (lp_gemm_set_raster_balance)")
    ap.add_argument("--tma-res", type=int, default=1, help="residual tile by TMA on / off (lp_gemm_set_tma_residual)")
    args = ap.parse_args()
    from leopard_b200 import _lib as _l
    _l.ensure_streamk_workspace("cuda:0")
    _l.lib().lp_gemm_set_streamk(args.streamk)
    _l.lib().lp_gemm_set_tma_residual(args.tma_res)
    _l.lib().lp_gemm_set_raster_balance(args.balance)
    if args.quad:
        from leopard_b200 import _lib
        _lib.lib().lp_gemm_set_pair_mode(2)
    if args.pair:
        from leopard_b200 import _lib
        _lib.lib().lp_gemm_set_pair_mode(1)
    if args.single:
        from leopard_b200 import _lib
        _lib.lib().lp_gemm_set_pair_mode(0)
    if args.list:
        print("\n".join(CASES))
        return 0
    names = args.case or list(CASES)
    ok = True
    for n in names:
        try:
            ok &= run_case(n, args.time)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"case": n, "ok": False, "exception": repr(e)[:500]}), flush=True)
            ok = False
            if "CUDA" in repr(e) or "cuda" in repr(e):
                break
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
