"""Randomised runs of the CPU-emulated kernels that have not been on hardware yet: the flash attention backward
(lp_attn_bwd.cu) and the 4-CTA multicast GEMM (lp_gemm4.cu).  Builds the same emulation libraries as the CPU tests, then draws
shapes until the time budget is spent.  Scheduling noise of the OS threads varies from run to run, which is the point: a protocol
race shows up as a mismatch or as the emulator's deadlock abort.
    python tools/emu_stress.py [--seconds 300] [--which attn,gemm4] [--seed 1]
"""
import argparse
import ctypes as C
import random
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import native_ops as ref  # noqa: E402

BF = torch.bfloat16
LL = C.c_longlong


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _helpers(out, from_marker, fix_tanh=False):
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index(from_marker))
    text = common[a:common.rindex("}  // namespace lp")]
    if fix_tanh:
        text = text.replace('asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k0 * (x + k1 * x * x * x)));',
                            "t = tanhf(k0 * (x + k1 * x * x * x));")
    (out / "emu_helpers.inc").write_text("namespace lp {\n" + text + "\n}\n")


def _gxx(out, src, lib, *flags):
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", *flags, "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu" / src), "-o", str(out / lib)], capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr[-3000:])
    return C.CDLL(str(out / lib))


def build_attn(out):
    _helpers(out, "// UMMA descriptors")
    return _gxx(out, "attn_bwd_emu.cpp", "libattn.so")


def build_gemm(out):
    _helpers(out, "// UMMA descriptors", fix_tanh=True)

    def part(path, end):
        src = (ROOT / path).read_text()
        return src[src.index("namespace lp {"):src.index(end)].replace("extern __shared__ uint8_t smem_raw[];",
                                                                       "uint8_t* smem_raw = emu_tc::cta().smem;")

    shared = (ROOT / "leopard_b200/csrc/lp_gemm_common.cuh").read_text()
    shared = shared[shared.index("namespace lp {"):].replace('asm volatile("bar.sync 1, 256;" ::: "memory");', "named_bar(1, 256, true);")
    k1 = part("leopard_b200/csrc/lp_gemm.cu", "template <int BN, int STAGES, bool CTA2>\nstatic int launch_gemm(")
    k4 = part("leopard_b200/csrc/lp_gemm4.cu", "#ifndef LP_EMULATE\ntemplate <int BN, int STAGES>\nstatic int launch4(")
    (out / "gemm_kernel.inc").write_text(shared + "\n" + k1 + "\n}  // namespace lp\n")
    (out / "gemm4_kernel.inc").write_text(k4 + "\n}  // namespace lp\n")
    return _gxx(out, "gemm_emu.cpp", "libgemm.so", "-DLP_EMULATE=1")


def attn_case(L, n):
    hd, hs = random.choice([(128, 128), (128, 128), (72, 80)])
    causal = False if hd == 72 else random.choice([True, True, False])
    kvh = random.choice([1, 2])
    heads = kvh * random.choice([1, 2, 4])
    T = random.choice([8, 16, 127, 128, 129, 136, 255, 256, 264, 300, 392])
    B = random.choice([1, 2])
    g = torch.Generator().manual_seed(n)
    qw, kw = heads * hs, kvh * hs
    qkv = (torch.randn((B * T, qw + 2 * kw), generator=g) * random.choice([0.5, 1.0, 2.0])).to(BF)
    if hs != hd:
        v = qkv.view(B * T, heads + 2 * kvh, hs)
        v[:, :, hd:] = 0
        qkv = v.reshape(B * T, -1).contiguous()
    dout = torch.randn((B * T, heads * hd), generator=g).to(BF)
    out = ref.attention(qkv, B, T, heads, kvh, hd, hs, 0, qw, qw + kw, causal)
    want = ref.attention_bwd(qkv, out, dout, B, T, heads, kvh, hd, 0, qw, qw + kw, causal, head_stride=hs)
    got = torch.full_like(qkv, float("nan"))
    tp = -(-T // 128) * 128
    lse, dl = torch.zeros(B * heads * tp), torch.zeros(B * heads * tp)
    L.emu_attention_bwd(P(qkv), LL(qkv.stride(0)), P(out), P(dout), LL(out.stride(0)), B, T, heads, kvh, hd, hs, 0, qw, qw + kw,
                        int(causal), C.c_float(hd ** -0.5), P(got), LL(got.stride(0)), P(lse), P(dl), 3)
    e = float((got.float() - want.float()).abs().max() / (want.float().abs().max() + 1e-12))
    return (e > 2e-2 or bool(torch.isnan(got.float()).any())), f"B{B} T{T} heads {heads}/{kvh} causal={causal} hd={hd}: {e:.4f}"


def gemm4_case(L, n):
    bn = random.choice([128, 128, 192, 256])
    M, K = random.choice([130, 256, 300, 512, 777]), random.choice([64, 72, 136, 200, 320])
    N = bn * random.choice([2, 3, 4, 5]) - random.choice([0, 0, 8, 40, bn // 2])
    act = random.choice([0, 0, 1, 2, 3])
    if act == 3:
        N -= N % 64
    bias, res = random.random() < 0.5, random.random() < 0.5
    fp32 = act != 3 and random.random() < 0.3
    res = res and not fp32
    ctas, gm = random.choice([4, 8, 12]), random.choice([1, 2, 3])
    g = torch.Generator().manual_seed(n)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g) if bias else None
    no = N // 2 if act == 3 else N
    r = torch.randn((M, no), generator=g).to(BF) if res else None
    want = ref.gemm(a, w, bias=b, residual=r, act=act, out_fp32=fp32)
    out = torch.full((M, no), float("nan"), dtype=torch.float32 if fp32 else BF)
    L.emu_gemm4(P(a), LL(K), P(w), LL(K), P(out), LL(no), M, N, K, P(b), P(r), LL(no), 0, act, int(fp32), bn, gm, ctas)
    e = float((out.float() - want.float()).abs().max() / (want.float().abs().max() + 1e-12))
    return (e > 8e-3 or bool(torch.isnan(out.float()).any())), f"M{M} N{N} K{K} bn{bn} act{act} bias={bias} res={res} fp32={fp32} ctas={ctas}: {e:.2e}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=int, default=300)
    ap.add_argument("--which", default="attn,gemm4")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    random.seed(a.seed)
    which = a.which.split(",")
    with tempfile.TemporaryDirectory() as d:
        out = Path(d)
        jobs = []
        if "attn" in which:
            (out / "a").mkdir()
            jobs.append(("attn", build_attn(out / "a"), attn_case))
        if "gemm4" in which:
            (out / "g").mkdir()
            jobs.append(("gemm4", build_gemm(out / "g"), gemm4_case))
        for name, lib, fn in jobs:
            t0, n, bad = time.time(), 0, 0
            while time.time() - t0 < a.seconds / len(jobs):
                is_bad, desc = fn(lib, n)
                n += 1
                if is_bad:
                    bad += 1
                    print("BAD", name, desc, flush=True)
            print(f"{name}: {n} cases, {bad} bad")


if __name__ == "__main__":
    main()
