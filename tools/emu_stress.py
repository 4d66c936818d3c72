"""Randomised runs of the CPU-emulated kernels that have not been on hardware yet: the flash attention backward
(lp_attn_bwd.cu) and the 4-CTA multicast GEMM (lp_gemm4.cu).  Builds the same emulation libraries as the CPU tests, then draws
shapes until the time budget is spent.  Scheduling noise of the OS threads varies from run to run, which is the point: a protocol
race shows up as a mismatch or as the emulator's deadlock abort.
    python tools/emu_stress.py [--seconds 300] [--which attn,gemm4] [--seed 1]
`--which fwd` adds the forward attention kernels whose scheduling changed last (lp_attn3.cu: next S half loaded before the hand-over,
alternating pv_done barriers; lp_attn5.cu: persistent CTAs, barrier parities running across items) — 765 random cases in 600 s
passed before the round closed.
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


def build_fwd(out):
    """lp_attn3.cu / lp_attn5.cu kernel templates under the emulator (same cuts as tests/test_attention_kernels_emulated.py)."""
    _helpers(out, "// UMMA descriptors")
    host = (ROOT / "leopard_b200/csrc/lp_host.h").read_text()
    (out / "attn_args.inc").write_text(re.search(r"struct AttnArgs \{.*?\n\};", host, re.S).group(0) + "\n")
    src = (ROOT / "leopard_b200/csrc/lp_attn5.cu").read_text()
    (out / "attn5_kernel.inc").write_text(src[src.index("// ---- kernel code"):src.index("// ---- host side")])
    src = (ROOT / "leopard_b200/csrc/lp_attn3.cu").read_text()
    (out / "attn3_kernel.inc").write_text(src[src.index("template <int DH, bool CAUSAL, int EMU>\n__global__"):
                                              src.index("template <int DH, bool CAUSAL, int EMU>\nstatic int launch3(")])
    return (_gxx(out, "attn5_fwd_emu.cpp", "liba5.so"), _gxx(out, "attn3_fwd_emu.cpp", "liba3.so"))


def fwd_case(libs, n):
    """60 % persistent-kernel cases (eligible T, 1..7 CTAs over B x heads x pairs items), 40 % v3 cases (causal GQA / full / tower layout)."""
    l5, l3 = libs
    if random.random() < 0.6:
        while True:
            T = random.randint(256, 1100)
            if T % 256 == 0 or T % 256 > 128:
                break
        B, heads, grid, ping = random.randint(1, 3), random.randint(1, 3), random.randint(1, 7), random.randint(0, 1)
        hd, hs, causal, kvh = 72, 80, False, heads
    else:
        T, B, causal, grid, ping = random.randint(1, 700), random.randint(1, 2), random.random() < 0.6, 0, random.randint(0, 1)
        if random.random() < 0.4:
            hd, hs, heads, causal = 72, 80, random.randint(1, 2), False
            kvh = heads
        else:
            hd, hs, kvh = 128, 128, random.choice([1, 2])
            heads = kvh * random.choice([1, 2, 4])
    g = torch.Generator().manual_seed(random.randint(0, 1 << 30))
    q_w, kv_w = heads * hs, kvh * hs
    qkv = torch.randn((B * T, q_w + 2 * kv_w), generator=g).to(BF)
    if hs != hd:
        v = qkv.view(B * T, heads + 2 * kvh, hs)
        v[:, :, hd:] = 0
        qkv = v.reshape(B * T, -1).contiguous()
    want = ref.attention(qkv, B, T, heads, kvh, hd, hs, 0, q_w, q_w + kv_w, causal)
    got = torch.full((B * T, heads * hd), float("nan")).to(BF)
    if grid:
        l5.emu_attention5_fwd(P(qkv), LL(qkv.stride(0)), B, T, heads, kvh, hd, hs, 0, q_w, q_w + kv_w, 0, C.c_float(hd ** -0.5), P(got),
                              LL(got.stride(0)), ping, grid)
    else:
        l3.emu_attention3_fwd(P(qkv), LL(qkv.stride(0)), B, T, heads, kvh, hd, hs, 0, q_w, q_w + kv_w, int(causal), C.c_float(hd ** -0.5),
                              P(got), LL(got.stride(0)), ping)
    d = (got.float() - want.float()).abs()
    bad = bool(torch.isnan(got.float()).any()) or int((d > 2e-2 * want.float().abs().clamp(min=0.5)).sum()) > 0
    return bad, f"{'v5' if grid else 'v3'} B={B} T={T} heads={heads}/{kvh} dh={hd} causal={causal} ping={ping} grid={grid} max_err={float(d.max()):.4f}"


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
        if "fwd" in which:
            (out / "f").mkdir()
            jobs.append(("fwd", build_fwd(out / "f"), fwd_case))
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
