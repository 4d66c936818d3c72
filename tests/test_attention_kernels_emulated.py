"""The SOURCE of the tcgen05 attention kernels executed on the CPU — the flash-style backward (leopard_b200/csrc/lp_attn_bwd.cu,
compiled with -DLP_EMULATE), and as the emulator's credibility check the two GPU-verified forward kernels (lp_attn.cu, lp_attn3.cu): one OS thread per CUDA thread, the tcgen05 / TMEM / TMA / mbarrier wrappers replaced by the functional stand-ins of
tests/emu/tc_emu.h.  Checks the kernels' logic — the producer / issuer / row-thread protocol (a wedged barrier aborts the run),
tile indexing, causal and ragged masks, the GQA group accumulation, the operand roles of the seven MMAs per tile pair, the
zero-padded 80-wide head layout — against the closed form of oracle/train.py.  It does not check the hardware: the emulator's
memory layouts are the same reading of the hardware the GPU-verified forward kernels are built on."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest
import torch

from oracle import native_ops as ref

ROOT = Path(__file__).resolve().parents[1]
BF = torch.bfloat16


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu_attn")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// UMMA descriptors"))
    (out / "emu_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    lib = out / "libattn.so"
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/attn_bwd_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


def P(t):
    return C.c_void_p(t.data_ptr())


def _err(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


@pytest.mark.parametrize("B,T,heads,kv_heads,causal,hd,hs", [
    (1, 128, 1, 1, True, 128, 128),      # one tile
    (2, 200, 4, 2, True, 128, 128),      # ragged length, two sequences, GQA groups of 2
    (1, 384, 2, 2, False, 128, 128),     # full attention, three tiles
    (1, 520, 8, 2, True, 128, 128),      # the decoder's group of 4, five tiles (S buffers wrap, barrier phases flip twice)
    (2, 300, 3, 3, False, 72, 80),       # vision-tower layout: 72 used of 80 columns per head
])
def test_flash_backward_kernels_under_emulation(emu, B, T, heads, kv_heads, causal, hd, hs):
    g = torch.Generator().manual_seed(T + heads)
    q_w, kv_w = heads * hs, kv_heads * hs
    qkv = torch.randn((B * T, q_w + 2 * kv_w), generator=g).to(BF)
    if hs != hd:
        v = qkv.view(B * T, heads + 2 * kv_heads, hs)
        v[:, :, hd:] = 0
        qkv = v.reshape(B * T, -1).contiguous()
    dout = torch.randn((B * T, heads * hd), generator=g).to(BF)
    out = ref.attention(qkv, B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, causal)
    want = ref.attention_bwd(qkv, out, dout, B, T, heads, kv_heads, hd, 0, q_w, q_w + kv_w, causal, head_stride=hs)
    got = torch.full_like(qkv, float("nan"))
    t_pad = -(-T // 128) * 128
    lse, delta = torch.zeros(B * heads * t_pad), torch.zeros(B * heads * t_pad)
    emu.emu_attention_bwd(P(qkv), C.c_longlong(qkv.stride(0)), P(out), P(dout), C.c_longlong(out.stride(0)), B, T, heads,
                          kv_heads, hd, hs, 0, q_w, q_w + kv_w, int(causal), C.c_float(hd ** -0.5), P(got),
                          C.c_longlong(got.stride(0)), P(lse), P(delta), 3)
    assert not bool(torch.isnan(got.float()).any())                      # every element of dqkv is written
    for name, sl in (("dq", slice(0, q_w)), ("dk", slice(q_w, q_w + kv_w)), ("dv", slice(q_w + kv_w, None))):
        assert _err(got[:, sl], want[:, sl]) < 1.5e-2, name
    if hs != hd:
        assert float(got.view(B * T, heads + 2 * kv_heads, hs)[:, :, hd:].abs().max()) == 0.0
    # the row statistics the first kernel hands to the second: natural-log lse = lse2 * ln 2
    x = qkv.float().view(B, T, -1)
    q0 = x[0, :, :hd]
    k0 = x[0, :, q_w: q_w + hd]
    s = (q0 @ k0.t()) * hd ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool).triu(1), float("-inf"))
    assert torch.allclose(lse[:T] * 0.6931471805599453, torch.logsumexp(s, -1), atol=2e-3)


# ---- the emulator's own credibility check: the forward kernel that is parity-green on the B200 must be green here too ------
@pytest.fixture(scope="module")
def emu_fwd(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu_fwd")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// UMMA descriptors"))
    (out / "emu_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    args = re.search(r"struct AttnArgs \{.*?\n\};", (ROOT / "leopard_b200/csrc/lp_host.h").read_text(), re.S).group(0)
    src = (ROOT / "leopard_b200/csrc/lp_attn.cu").read_text()
    body = src[src.index("namespace lp {"):src.index("template <int DH, bool CAUSAL, bool P_TMEM>\nstatic int launch_attn(")]
    # the kernel's two inline-PTX statements -> functional equivalents (everything else goes through lp_common.cuh wrappers)
    ex2 = 'asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));'
    assert ex2 in body
    body = body.replace(ex2, "y = exp2f(x);")
    st16 = re.compile(r'asm volatile\(\s*"tcgen05\.st\.sync\.aligned\.32x32b\.x16\.b32.*?::\s*"r"\(([^,]*?)\),\s*"r"\((\w+)\[0\]\).*?'
                      r':\s*"memory"\);', re.S)
    body, n = st16.subn(lambda m: f"lp::tmem_st({m.group(2)}, {m.group(1)}, 16);", body)
    assert n == 2 and "asm" not in body
    (out / "attn_fwd.inc").write_text("namespace lp {\n" + args + "\n}\n" + body + "\n}  // namespace lp\n")
    lib = out / "libfwd.so"
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/attn_fwd_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


@pytest.mark.parametrize("B,T,heads,kv_heads,causal,hd,hs,p_tmem", [
    (2, 200, 4, 2, True, 128, 128, 0),      # P through shared memory (manual 128-byte swizzle writes)
    (1, 520, 4, 1, True, 128, 128, 1),      # P aliased over S in TMEM, five kv tiles
    (2, 300, 2, 2, False, 72, 80, 1),       # vision-tower layout
])
def test_gpu_verified_forward_kernel_is_green_under_the_emulator(emu_fwd, B, T, heads, kv_heads, causal, hd, hs, p_tmem):
    g = torch.Generator().manual_seed(T)
    q_w, kv_w = heads * hs, kv_heads * hs
    qkv = torch.randn((B * T, q_w + 2 * kv_w), generator=g).to(BF)
    if hs != hd:
        v = qkv.view(B * T, heads + 2 * kv_heads, hs)
        v[:, :, hd:] = 0
        qkv = v.reshape(B * T, -1).contiguous()
    want = ref.attention(qkv, B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, causal)
    got = torch.full((B * T, heads * hd), float("nan")).to(BF)
    emu_fwd.emu_attention_fwd(P(qkv), C.c_longlong(qkv.stride(0)), B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, int(causal),
                              C.c_float(hd ** -0.5), P(got), C.c_longlong(got.stride(0)), p_tmem)
    assert not bool(torch.isnan(got.float()).any()) and _err(got, want) < 1.5e-2


@pytest.fixture(scope="module")
def emu_v3(tmp_path_factory):
    """lp_attn3.cu: the kernel template is cut out verbatim (it contains no PTX); its few inline-PTX helpers (a3_*) are restated
    in tests/emu/attn3_fwd_emu.cpp."""
    out = tmp_path_factory.mktemp("emu_v3")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// UMMA descriptors"))
    (out / "emu_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    host = (ROOT / "leopard_b200/csrc/lp_host.h").read_text()
    (out / "attn_args.inc").write_text(re.search(r"struct AttnArgs \{.*?\n\};", host, re.S).group(0) + "\n")
    src = (ROOT / "leopard_b200/csrc/lp_attn3.cu").read_text()
    body = src[src.index("template <int DH, bool CAUSAL, int EMU>\n__global__"):
               src.index("template <int DH, bool CAUSAL, int EMU>\nstatic int launch3(")]
    assert "asm" not in body
    (out / "attn3_kernel.inc").write_text(body)
    lib = out / "liba3.so"
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/attn3_fwd_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


@pytest.mark.parametrize("B,T,heads,kv_heads,causal,hd,hs,pingpong", [
    (2, 200, 4, 2, True, 128, 128, 1),      # one query-tile pair, the second tile ragged; exp-section turn protocol on
    (1, 520, 4, 1, True, 128, 128, 0),      # three pairs (the last one half empty), turn protocol off
    (2, 300, 2, 2, False, 72, 80, 1),       # vision-tower layout
    (1, 640, 2, 2, False, 128, 128, 1),     # full attention, five kv tiles for both query tiles
])
def test_default_prefill_attention_kernel_is_green_under_the_emulator(emu_v3, B, T, heads, kv_heads, causal, hd, hs, pingpong):
    """attn3_fwd_kernel is what every prefill runs (variant 80); parity-green on the B200, and a CPU regression check here."""
    g = torch.Generator().manual_seed(T + 1)
    q_w, kv_w = heads * hs, kv_heads * hs
    qkv = torch.randn((B * T, q_w + 2 * kv_w), generator=g).to(BF)
    if hs != hd:
        v = qkv.view(B * T, heads + 2 * kv_heads, hs)
        v[:, :, hd:] = 0
        qkv = v.reshape(B * T, -1).contiguous()
    want = ref.attention(qkv, B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, causal)
    got = torch.full((B * T, heads * hd), float("nan")).to(BF)
    emu_v3.emu_attention3_fwd(P(qkv), C.c_longlong(qkv.stride(0)), B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, int(causal),
                              C.c_float(hd ** -0.5), P(got), C.c_longlong(got.stride(0)), pingpong)
    assert not bool(torch.isnan(got.float()).any()) and _err(got, want) < 1.5e-2


@pytest.fixture(scope="module")
def emu_v4(tmp_path_factory):
    """lp_attn4.cu: everything below its "kernel code" marker (the half-step function and the kernel template, no PTX) is cut out
    verbatim; the inline-PTX helpers (a4_*) are restated in tests/emu/attn4_fwd_emu.cpp."""
    out = tmp_path_factory.mktemp("emu_v4")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// UMMA descriptors"))
    (out / "emu_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    host = (ROOT / "leopard_b200/csrc/lp_host.h").read_text()
    (out / "attn_args.inc").write_text(re.search(r"struct AttnArgs \{.*?\n\};", host, re.S).group(0) + "\n")
    src = (ROOT / "leopard_b200/csrc/lp_attn4.cu").read_text()
    body = src[src.index("// ---- kernel code"):src.index("template <int DH, bool CAUSAL, int EMU>\nstatic int launch4(")]
    assert "asm" not in body
    (out / "attn4_kernel.inc").write_text(body)
    lib = out / "liba4.so"
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/attn4_fwd_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


@pytest.mark.parametrize("B,T,heads,kv_heads,causal,hd,hs", [
    (2, 200, 4, 2, True, 128, 128),      # one query-tile pair, the second tile ragged: dual steps + tile B's diagonal alone
    (1, 520, 4, 1, True, 128, 128),      # three pairs, the last one without a tile B (single-stream body for tile A)
    (2, 300, 2, 2, False, 72, 80),       # vision-tower layout, trailing half tile skipped
    (1, 640, 2, 2, False, 128, 128),     # full attention, five kv tiles for both query tiles
    (1, 100, 2, 1, True, 128, 128),      # a single query tile
])
def test_two_stream_attention_kernel_under_the_emulator(emu_v4, B, T, heads, kv_heads, causal, hd, hs):
    """attn4_fwd_kernel (variant bit 7): same arithmetic per row as v3, so the same bound against the fp32 reference."""
    g = torch.Generator().manual_seed(T + 1)
    q_w, kv_w = heads * hs, kv_heads * hs
    qkv = torch.randn((B * T, q_w + 2 * kv_w), generator=g).to(BF)
    if hs != hd:
        v = qkv.view(B * T, heads + 2 * kv_heads, hs)
        v[:, :, hd:] = 0
        qkv = v.reshape(B * T, -1).contiguous()
    want = ref.attention(qkv, B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, causal)
    got = torch.full((B * T, heads * hd), float("nan")).to(BF)
    emu_v4.emu_attention4_fwd(P(qkv), C.c_longlong(qkv.stride(0)), B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, int(causal),
                              C.c_float(hd ** -0.5), P(got), C.c_longlong(got.stride(0)), 0)
    assert not bool(torch.isnan(got.float()).any()) and _err(got, want) < 1.5e-2


@pytest.fixture(scope="module")
def emu_v5(tmp_path_factory):
    """lp_attn5.cu (persistent v3 for the vision tower's short sequences): everything below its "kernel code" marker up to the host
    side is cut out verbatim; the inline-PTX helpers (a5_*) are restated in tests/emu/attn5_fwd_emu.cpp."""
    out = tmp_path_factory.mktemp("emu_v5")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// UMMA descriptors"))
    (out / "emu_helpers.inc").write_text("namespace lp {\n" + common[a:common.rindex("}  // namespace lp")] + "\n}\n")
    host = (ROOT / "leopard_b200/csrc/lp_host.h").read_text()
    (out / "attn_args.inc").write_text(re.search(r"struct AttnArgs \{.*?\n\};", host, re.S).group(0) + "\n")
    src = (ROOT / "leopard_b200/csrc/lp_attn5.cu").read_text()
    body = src[src.index("// ---- kernel code"):src.index("// ---- host side")]
    assert "asm" not in body
    (out / "attn5_kernel.inc").write_text(body)
    lib = out / "liba5.so"
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-I", str(out), "-I", str(ROOT / "tests/emu"),
                        str(ROOT / "tests/emu/attn5_fwd_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


@pytest.mark.parametrize("B,T,heads,kv_heads,grid,pingpong", [
    (2, 300, 3, 3, 2, 1),     # 12 items on 2 CTAs: six items each, trailing half tile skipped, turn protocol across item boundaries
    (1, 676, 2, 2, 4, 1),     # the tower's T: 6 items on 4 CTAs (two CTAs get two items, two get one)
    (1, 512, 2, 2, 1, 0),     # T a multiple of 256, one CTA walks all four items; turn protocol off
    (3, 420, 1, 1, 8, 1),     # more CTAs than items per head; grid > items is clamped
])
def test_persistent_attention_kernel_under_the_emulator(emu_v5, B, T, heads, kv_heads, grid, pingpong):
    """attn5_fwd_kernel (variant bit 8): v3's arithmetic per row, barrier parities running across items."""
    hd, hs = 72, 80
    g = torch.Generator().manual_seed(T + grid)
    q_w, kv_w = heads * hs, kv_heads * hs
    qkv = torch.randn((B * T, q_w + 2 * kv_w), generator=g).to(BF)
    v = qkv.view(B * T, heads + 2 * kv_heads, hs)
    v[:, :, hd:] = 0
    qkv = v.reshape(B * T, -1).contiguous()
    want = ref.attention(qkv, B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, False)
    got = torch.full((B * T, heads * hd), float("nan")).to(BF)
    emu_v5.emu_attention5_fwd(P(qkv), C.c_longlong(qkv.stride(0)), B, T, heads, kv_heads, hd, hs, 0, q_w, q_w + kv_w, 0,
                              C.c_float(hd ** -0.5), P(got), C.c_longlong(got.stride(0)), pingpong, grid)
    assert not bool(torch.isnan(got.float()).any()) and _err(got, want) < 1.5e-2
