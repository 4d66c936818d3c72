"""The SOURCE of the persistent tcgen05 GEMMs (leopard_b200/csrc/lp_gemm.cu, lp_gemm4.cu) executed on the CPU under tests/emu/tc_emu.h —
one CTA per tile, CTA pairs (cluster of 2, cta_group::2 MMA over both CTAs' shared memory and TMEM, remote mbarrier arrives,
multicast commits) and the 4-CTA-cluster variant whose two pairs share the A tile by TMA multicast (not yet run on hardware:
this is its logic check).  The first two are parity-green on the B200; here it is (a) the emulator's credibility check for the cluster features and
(b) a CPU regression check of the GEMM's logic — tile scheduler, ring / accumulator barriers, ragged M / N / K, the bias /
GELU / SwiGLU / residual / fp32 epilogues and the swizzled output staging — for kernel work done without a GPU at hand."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest
import torch

from oracle import native_ops as ref

ROOT = Path(__file__).resolve().parents[1]
BF = torch.bfloat16
LL = C.c_longlong


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu_gemm")
    common = (ROOT / "leopard_b200/csrc/lp_common.cuh").read_text()
    a = common.rindex("// ----", 0, common.index("// UMMA descriptors"))
    helpers = common[a:common.rindex("}  // namespace lp")]
    tanh_asm = 'asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k0 * (x + k1 * x * x * x)));'
    assert tanh_asm in helpers
    (out / "emu_helpers.inc").write_text("namespace lp {\n" + helpers.replace(tanh_asm, "t = tanhf(k0 * (x + k1 * x * x * x));") + "\n}\n")
    def kernel_part(path, end_marker):
        src = (ROOT / path).read_text()
        body = src[src.index("namespace lp {"):src.index(end_marker)]
        decl = "extern __shared__ uint8_t smem_raw[];"                      # -> the executing CTA's own buffer
        assert body.count(decl) == 1, path
        return body.replace(decl, "uint8_t* smem_raw = emu_tc::cta().smem;")

    shared = (ROOT / "leopard_b200/csrc/lp_gemm_common.cuh").read_text()
    shared = shared[shared.index("namespace lp {"):]
    named = 'asm volatile("bar.sync 1, 256;" ::: "memory");'
    assert shared.count(named) == 1
    shared = shared.replace(named, "named_bar(1, 256, true);")
    k1 = kernel_part("leopard_b200/csrc/lp_gemm.cu", "template <int BN, int STAGES, bool CTA2>\nstatic int launch_gemm(")
    k4 = kernel_part("leopard_b200/csrc/lp_gemm4.cu", "#ifndef LP_EMULATE\ntemplate <int BN, int STAGES>\nstatic int launch4(")
    assert "asm" not in shared + k1 + k4
    (out / "gemm_kernel.inc").write_text(shared + "\n" + k1 + "\n}  // namespace lp\n")
    (out / "gemm4_kernel.inc").write_text(k4 + "\n}  // namespace lp\n")
    lib = out / "libgemm.so"
    r = subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-DLP_EMULATE=1", "-I", str(out), "-I",
                        str(ROOT / "tests/emu"), str(ROOT / "tests/emu/gemm_emu.cpp"), "-o", str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(str(lib))


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("M,N,K,bn,pair,ctas,act,bias,res,fp32,group_m,rmod", [
    (300, 256, 192, 128, False, 2, 0, True, False, False, 1, 0),      # one CTA per tile, ragged M, persistent over 6 tiles
    (600, 384, 320, 128, True, 4, 0, True, True, False, 2, 0),        # CTA pairs, residual, grouped rasterisation
    (520, 504, 200, 256, True, 2, 1, True, False, False, 1, 0),       # pairs, ragged N (narrower last MMA), ragged K, GELU-tanh
    (384, 512, 136, 192, True, 2, 3, True, True, False, 1, 0),        # pairs, SwiGLU on interleaved granules + residual
    (333, 264, 72, 160, False, 3, 2, True, False, True, 1, 0),        # fp32 output, exact GELU, the 160-wide tile
    (1352, 256, 128, 256, False, 3, 0, True, True, False, 1, 676),    # residual row = row % 676 (position embedding)
    (700, 392, 256, 128, True, 6, 0, False, False, True, 1, 0),       # pairs, fp32 output, more CTAs than tiles in the last wave
])
def test_gemm_kernel_under_emulation(emu, M, N, K, bn, pair, ctas, act, bias, res, fp32, group_m, rmod):
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g) if bias else None
    n_out = N // 2 if act == 3 else N
    r = torch.randn((rmod if rmod else M, n_out), generator=g).to(BF) if res else None
    want = ref.gemm(a, w, bias=b, residual=r, act=act, out_fp32=fp32, res_row_mod=rmod)
    out = torch.full((M, n_out), float("nan"), dtype=torch.float32 if fp32 else BF)
    emu.emu_gemm(P(a), LL(K), P(w), LL(K), P(out), LL(n_out), M, N, K, P(b), P(r), LL(n_out), rmod, act, int(fp32), bn, int(pair),
                 group_m, ctas)
    assert not bool(torch.isnan(out.float()).any())
    err = float((out.float() - want.float()).abs().max() / want.float().abs().max())
    assert err < (1e-5 if fp32 else 8e-3), err


@pytest.mark.parametrize("M,N,K,bn,ctas,act,bias,res,fp32,group_m", [
    (600, 512, 320, 128, 4, 0, True, True, False, 2),       # two column-block pairs per row block, one cluster, residual
    (520, 384, 200, 128, 8, 1, True, False, False, 1),      # three column blocks: the second pair idles on the odd one out
    (384, 1024, 136, 256, 4, 3, True, True, False, 1),      # SwiGLU + residual on 256-wide tiles, ragged K
    (300, 700, 264, 192, 4, 0, True, False, True, 1),       # ragged M and N, fp32 output, 192-wide tiles
])
def test_four_cta_multicast_gemm_under_emulation(emu, M, N, K, bn, ctas, act, bias, res, fp32, group_m):
    """lp_gemm4.cu: every CTA requests half of its A tile and the TMA multicasts it to the same row half of the other pair;
    full barriers count bytes issued by both pairs, empty barriers wait for both pairs' commits."""
    g = torch.Generator().manual_seed(M + N + 4)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g) if bias else None
    n_out = N // 2 if act == 3 else N
    r = torch.randn((M, n_out), generator=g).to(BF) if res else None
    want = ref.gemm(a, w, bias=b, residual=r, act=act, out_fp32=fp32)
    out = torch.full((M, n_out), float("nan"), dtype=torch.float32 if fp32 else BF)
    emu.emu_gemm4(P(a), LL(K), P(w), LL(K), P(out), LL(n_out), M, N, K, P(b), P(r), LL(n_out), 0, act, int(fp32), bn, group_m, ctas)
    assert not bool(torch.isnan(out.float()).any())
    err = float((out.float() - want.float()).abs().max() / want.float().abs().max())
    assert err < (1e-5 if fp32 else 8e-3), err
