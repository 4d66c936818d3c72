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
    kc = kernel_part("leopard_b200/csrc/lp_gemm_chain.cu", "#ifndef LP_EMULATE\n// counters live in")
    assert "asm" not in shared + k1 + k4 + kc
    (out / "gemm_chain_kernel.inc").write_text(kc + "\n}  // namespace lp\n")
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


# ---- epilogue fusions of lp_gemm_bf16_ex: row statistics (producer), folded LayerNorm / RMSNorm (consumer), RoPE -------------
class _Fx(C.Structure):
    _fields_ = [("norm_mode", C.c_int), ("stats_in", C.c_void_p), ("stats_in_parts", C.c_int), ("colsum", C.c_void_p),
                ("norm_eps", C.c_float), ("stats_out", C.c_void_p), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
                ("rope_heads", C.c_int)]


def _gemm_ex(emu, a, w, out, bias=None, res=None, act=0, bn=128, pair=False, ctas=2, fx=None):
    M, K = a.shape
    N = w.shape[0]
    emu.emu_gemm_ex(P(a), LL(a.stride(0)), P(w), LL(w.stride(0)), P(out), LL(out.stride(0)), M, N, K, P(bias), P(res),
                    LL(res.stride(0) if res is not None else 0), 0, act, int(out.dtype == torch.float32), bn, int(pair), 1, ctas,
                    C.byref(fx) if fx is not None else None)


def _split_stats(x, parts):
    """Row (sum, sum of squares) of x split into `parts` arbitrary column groups -> [M, parts, 2] fp32."""
    cols = torch.tensor_split(x.float(), parts, dim=1)
    return torch.stack([torch.stack([c.sum(1), (c * c).sum(1)], -1) for c in cols], 1).contiguous()


@pytest.mark.parametrize("M,N,K,bn,pair,ctas", [(300, 384, 136, 128, False, 2), (600, 1152, 200, 256, True, 4)])
def test_producer_epilogue_writes_row_statistics_of_the_rounded_outputs(emu, M, N, K, bn, pair, ctas):
    g = torch.Generator().manual_seed(7)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g)
    r = torch.randn((M, N), generator=g).to(BF)
    parts = 2 * (-(-N // bn))
    out = torch.full((M, N), float("nan"), dtype=BF)
    stats = torch.full((M, parts, 2), float("nan"))
    fx = _Fx(stats_out=stats.data_ptr())
    _gemm_ex(emu, a, w, out, bias=b, res=r, bn=bn, pair=pair, ctas=ctas, fx=fx)
    want = ref.gemm(a, w, bias=b, residual=r)
    assert float((out.float() - want.float()).abs().max() / want.float().abs().max()) < 8e-3
    assert not bool(torch.isnan(stats).any())                    # every (row, slot) written, also for ragged column blocks
    tot = stats.sum(1)
    o = out.float()
    assert torch.allclose(tot[:, 0], o.sum(1), rtol=1e-4, atol=1e-3) and torch.allclose(tot[:, 1], (o * o).sum(1), rtol=1e-4)


@pytest.mark.parametrize("mode,act,pair,bn", [(2, 1, False, 128), (2, 0, True, 256), (1, 3, True, 128), (1, 0, False, 256)])
def test_consumer_epilogue_applies_folded_norm(emu, mode, act, pair, bn):
    """norm(x) W^T (+ b) == rs * (x W'^T - mu * colsum) + (beta W^T + b) with W' = W diag(gamma): LayerNorm (+GELU) and RMSNorm
    (+SwiGLU) against the unfused arithmetic on the same bf16 inputs."""
    M, N, K, eps = 300, 512, 264, 1e-5
    g = torch.Generator().manual_seed(11 + mode)
    x = (torch.randn((M, K), generator=g) * 1.5 + 0.3).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    gamma, beta = 1 + 0.1 * torch.randn(K, generator=g), 0.1 * torch.randn(K, generator=g)
    b = torch.randn(N, generator=g) * 0.1
    xf, wf32 = x.float(), w.float()
    if mode == 2:
        h = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, eps)
        bias_f = wf32 @ beta + b
    else:
        h = xf * torch.rsqrt((xf * xf).mean(-1, keepdim=True) + eps) * gamma
        bias_f = None
    y = h @ wf32.t() + (b if mode == 2 else 0)
    if act == 1:
        y = torch.nn.functional.gelu(y, approximate="tanh")
    elif act == 3:
        yy = y.view(M, N // 64, 2, 32)
        y = (torch.nn.functional.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(M, N // 2)
    wfold = (wf32 * gamma[None]).to(BF)
    colsum = wfold.float().sum(1).contiguous()
    stats = _split_stats(x, 5)
    out = torch.full(y.shape, float("nan"), dtype=BF)
    fx = _Fx(norm_mode=mode, stats_in=stats.data_ptr(), stats_in_parts=5, colsum=colsum.data_ptr(), norm_eps=eps)
    _gemm_ex(emu, x, wfold, out, bias=bias_f, act=act, bn=bn, pair=pair, ctas=2, fx=fx)
    err = float((out.float() - y).abs().max() / y.abs().max())
    assert err < 1.5e-2, err


@pytest.mark.parametrize("pair", [False, True])
def test_rope_epilogue_rotates_the_leading_heads(emu, pair):
    """QKV GEMM with fused RMSNorm + rotate-half RoPE on the first 3 of 4 heads (128 columns each), BN = 256."""
    M, K, heads, rot, eps = 300, 136, 4, 3, 1e-5
    N = heads * 128
    g = torch.Generator().manual_seed(5)
    x = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    ang = torch.rand((M, 64), generator=g) * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    xf = x.float()
    rs = torch.rsqrt((xf * xf).mean(-1, keepdim=True) + eps)
    y = (rs * (xf @ w.float().t())).view(M, heads, 128)
    lo, hi = y[..., :64], y[..., 64:]
    rotd = torch.cat([lo * cos[:, None] - hi * sin[:, None], hi * cos[:, None] + lo * sin[:, None]], -1)
    want = torch.cat([rotd[:, :rot], y[:, rot:]], 1).reshape(M, N)
    stats = _split_stats(x, 1)
    out = torch.full((M, N), float("nan"), dtype=BF)
    fx = _Fx(norm_mode=1, stats_in=stats.data_ptr(), stats_in_parts=1, norm_eps=eps, rope_cos=cos.data_ptr(),
             rope_sin=sin.data_ptr(), rope_heads=rot)
    _gemm_ex(emu, x, w, out, bn=256, pair=pair, ctas=2, fx=fx)
    err = float((out.float() - want).abs().max() / want.abs().max())
    assert err < 8e-3, err


# ---- tensor-parallel launches (lp_gemm_bf16_tp): one rank's kernel, the peers' buffers in the same process ----------------------
def _ptr_array(tensors):
    arr = (C.c_void_p * 8)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


@pytest.mark.parametrize("tp,rank,act", [(2, 1, 0), (2, 0, 3), (4, 2, 0)])
def test_allgather_gemm_pulls_peer_rows_and_clips_segments(emu, tp, rank, act):
    """All-gather prologue + segmented rows: only the rank's own rows are valid in its A buffer on entry; the kernel copies the
    peers' segments from their buffers, multiplies own-segment tiles first, stores through per-segment maps (no row of another
    segment is ever written from a stale A tile) and posts its readiness to the peers."""
    sl, K, N = 300, 136, 512                                   # 300 rows per rank: tiles straddle nothing, last tile is ragged
    M = tp * sl
    g = torch.Generator().manual_seed(tp * 10 + rank)
    full = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    bufs = []
    for r in range(tp):                                        # rank r's buffer: its own rows valid, the rest poison
        b = torch.full((M, K), float("nan"), dtype=BF)
        b[r * sl:(r + 1) * sl] = full[r * sl:(r + 1) * sl]
        bufs.append(b)
    n_out = N // 2 if act == 3 else N
    out = torch.full((M, n_out), float("nan"), dtype=BF)
    counter = torch.zeros(1, dtype=torch.int32)
    epoch = 7
    sig_local = torch.full((16,), epoch, dtype=torch.int32)     # every peer has announced its rows
    sig_peers = [torch.zeros(16, dtype=torch.int32) for _ in range(tp)]
    ctas = 2          # ONE pair: the emulator runs pairs one after another, and the prologue's counter needs every CTA of the grid
    emu.emu_gemm_tp(P(bufs[rank]), LL(K), P(w), LL(K), P(out), LL(n_out), M, N, K, act, tp, rank, 1, _ptr_array(bufs), None,
                    P(counter), P(sig_local), _ptr_array(sig_peers), epoch, ctas)
    want = ref.gemm(full, w, act=act)
    assert torch.equal(bufs[rank], full)                       # gathered
    assert not bool(torch.isnan(out.float()).any())
    assert float((out.float() - want.float()).abs().max() / want.float().abs().max()) < 8e-3
    assert int(counter) == ctas
    for r in range(tp):
        assert int(sig_peers[r][rank]) == (epoch if r != rank else 0)


@pytest.mark.parametrize("tp,rank", [(2, 0), (4, 3)])
def test_gemm_reduce_scatter_stores_partials_into_the_owners_slots(emu, tp, rank):
    sl, K, N = 300, 200, 512
    M = tp * sl
    g = torch.Generator().manual_seed(tp + rank)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    stages = [torch.full((tp, sl, N), float("nan"), dtype=BF) for _ in range(tp)]
    counter = torch.zeros(1, dtype=torch.int32)
    sig_local = torch.zeros(16, dtype=torch.int32)
    sig_peers = [torch.zeros(16, dtype=torch.int32) for _ in range(tp)]
    epoch, ctas = 3, 4
    emu.emu_gemm_tp(P(a), LL(K), P(w), LL(K), None, LL(N), M, N, K, 0, tp, rank, 2, None, _ptr_array(stages), P(counter), P(sig_local),
                    _ptr_array(sig_peers), epoch, ctas)
    want = ref.gemm(a, w)
    for r in range(tp):                                        # owner r received this rank's partial sums for ITS rows in slot [rank]
        got = stages[r][rank]
        assert float((got.float() - want[r * sl:(r + 1) * sl].float()).abs().max() / want.float().abs().max()) < 8e-3
        others = [s for s in range(tp) if s != rank]
        assert bool(torch.isnan(stages[r][others].float()).all())              # nothing else was touched
        assert int(sig_peers[r][8 + rank]) == epoch                            # completion posted to every owner
    assert int(counter) == ctas


# ---- stream-K tail (lp_gemm_set_streamk_workspace): the accumulator of a tile cut along K is handed from unit u to unit u + 1 -----
@pytest.mark.parametrize("M,N,K,bn,pair,ctas,act,bias,res,stats", [
    (600, 384, 320, 128, True, 4, 0, True, True, True),        # pairs: 9 tiles on 2 units -> unit 0 starts tile 7, unit 1 finishes it; residual + row statistics
    (600, 384, 330, 128, False, 4, 1, True, False, False),     # one CTA per tile: 15 tiles on 4 units, ragged K, GELU
    (520, 760, 456, 256, True, 4, 0, False, True, False),      # 256-wide tiles, ragged last column block (narrow MMA) inside the stream-K range
    (384, 1024, 200, 128, True, 6, 3, True, True, False),      # SwiGLU granules, 3 units
    (200, 256, 192, 128, False, 3, 0, True, False, False),     # 4 tiles on 3 units: unit 1 has only a START and an END piece
])
def test_stream_k_tail_is_bit_identical_to_the_unsplit_schedule(emu, M, N, K, bn, pair, ctas, act, bias, res, stats):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g) if bias else None
    n_out = N // 2 if act == 3 else N
    r = torch.randn((M, n_out), generator=g).to(BF) if res else None
    parts = 2 * (-(-N // bn))
    outs, sts = [], []
    ws = torch.zeros(4096 + ctas * 128 * bn * 4, dtype=torch.uint8)
    for sk in (False, True):
        out = torch.full((M, n_out), float("nan"), dtype=BF)
        st = torch.full((M, parts, 2), float("nan"))
        fx = _Fx(stats_out=st.data_ptr()) if stats else None
        emu.emu_gemm_set_streamk(P(ws) if sk else None)
        try:
            _gemm_ex(emu, a, w, out, bias=b, res=r, act=act, bn=bn, pair=pair, ctas=ctas, fx=fx)
        finally:
            emu.emu_gemm_set_streamk(None)
        outs.append(out)
        sts.append(st)
    want = ref.gemm(a, w, bias=b, residual=r, act=act)
    assert not bool(torch.isnan(outs[1].float()).any())
    assert float((outs[1].float() - want.float()).abs().max() / want.float().abs().max()) < 8e-3
    assert torch.equal(outs[0], outs[1])                        # same summation order -> same bits
    if stats:
        assert torch.equal(sts[0], sts[1])
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0    # every hand-over flag consumed: the workspace is ready for the next launch
    assert float(ws[4096:].view(torch.float32).nan_to_num().abs().sum()) > 0   # and partial accumulators really went through it


def test_stream_k_schedule_covers_every_k_block_exactly_once():
    """Host-side restatement of gemm_schedule / gemm_item (lp_gemm_common.cuh): for the bench shapes, every (tile, k-block) is
    owned by exactly one unit, ranges differ by at most one k-block, a cut tile is shared by units u and u + 1 only."""
    def schedule(T, KB, P):
        dp = (T // P - 1) * P
        sk = T - dp
        total = sk * KB
        items = {u: [] for u in range(P)}
        for u in range(P):
            b, e = u * total // P, (u + 1) * total // P
            t_first, b_off = divmod(b, KB)
            t_last, e_off = (e - 1) // KB, e - ((e - 1) // KB) * KB
            has_end, has_start = b_off > 0, e_off < KB
            items[u] += [(u + i * P, 0, KB, "full") for i in range(dp // P)]
            if has_start:
                items[u].append((dp + t_last, 0, e_off, "start"))
            items[u] += [(dp + t, 0, KB, "full") for t in range(t_first + has_end, t_last - has_start + 1)]
            if has_end:
                items[u].append((dp + t_first, b_off, KB, "end"))
        return items, dp
    for T, KB, P in [(304, 64, 74), (304, 224, 74), (456, 64, 74), (192, 64, 74), (15, 6, 4), (9, 5, 2), (149, 18, 148)]:
        items, dp = schedule(T, KB, P)
        seen = {}
        for u, its in items.items():
            for t, k0, k1, kind in its:
                for k in range(k0, k1):
                    assert (t, k) not in seen
                    seen[(t, k)] = u
                if kind == "end":
                    start = [i for i in items[u - 1] if i[3] == "start"]
                    assert seen[(t, 0)] == u - 1 and start[0][0] == t and start[0][2] == k0
                    # the neighbour's START is finished before this unit reaches its END item (k-block times after the data-parallel waves)
                    t_start_done = start[0][2]
                    t_reach_end = sum(k1_ - k0_ for tt, k0_, k1_, kd in its if kd != "end" and tt >= dp)
                    assert t_reach_end >= t_start_done
        assert len(seen) == T * KB
        loads = [sum(k1 - k0 for _, k0, k1, _ in its) for its in items.values()]
        assert max(loads) - min(loads) <= 1


# ---- residual tile by TMA (gemm_bf16_res_kernel, lp_gemm_set_tma_residual): same bits as the register path ------------------------
@pytest.mark.parametrize("M,N,K,bn,ctas,bias,stats,sk,inplace", [
    (600, 384, 320, 128, 4, True, True, False, True),      # in-place residual stream (C == residual), row statistics
    (520, 1160, 200, 256, 2, True, False, False, False),   # 256-wide tiles, ragged M and N (half-empty last column block), persistent over 15 tiles
    (600, 760, 456, 192, 4, False, True, True, True),      # 192-wide tiles (chunk split 3 + 3), stream-K tail on top
    (300, 72, 136, 128, 2, True, False, False, False),     # N < one chunk pair: warpgroup 1 has nothing to do
])
def test_tma_residual_ring_matches_the_register_path(emu, M, N, K, bn, ctas, bias, stats, sk, inplace):
    g = torch.Generator().manual_seed(M * 3 + N)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g) if bias else None
    r = torch.randn((M, N), generator=g).to(BF)
    parts = 2 * (-(-N // bn))
    ws = torch.zeros(4096 + ctas * 128 * bn * 4, dtype=torch.uint8)
    outs, sts = [], []
    for tma in (0, 1):
        out = r.clone() if inplace else torch.full((M, N), float("nan"), dtype=BF)
        res = out if inplace else r
        st = torch.full((M, parts, 2), float("nan"))
        fx = _Fx(stats_out=st.data_ptr()) if stats else None
        emu.emu_gemm_set_tma_residual(tma)
        emu.emu_gemm_set_streamk(P(ws) if sk else None)
        try:
            _gemm_ex(emu, a, w, out, bias=b, res=res, bn=bn, pair=True, ctas=ctas, fx=fx)
        finally:
            emu.emu_gemm_set_tma_residual(0)
            emu.emu_gemm_set_streamk(None)
        outs.append(out)
        sts.append(st)
    want = ref.gemm(a, w, bias=b, residual=r)
    assert not bool(torch.isnan(outs[1].float()).any())
    assert float((outs[1].float() - want.float()).abs().max() / want.float().abs().max()) < 8e-3
    assert torch.equal(outs[0], outs[1])
    if stats:
        assert torch.equal(sts[0], sts[1])


# ---- chained projector (lp_gemm_chain.cu): both layers in one persistent launch -------------------------------------------------
@pytest.mark.parametrize("M,K1,N1,N2,act", [(300, 200, 512, 256, 2), (169, 136, 264, 520, 1)])
def test_chained_projector_kernel_under_emulation(emu, M, K1, N1, N2, act):
    g = torch.Generator().manual_seed(M + N1)
    x = torch.randn((M, K1), generator=g).to(BF)
    w1 = (torch.randn((N1, K1), generator=g) * K1 ** -0.5).to(BF)
    w2 = (torch.randn((N2, N1), generator=g) * N1 ** -0.5).to(BF)
    b1, b2 = torch.randn(N1, generator=g), torch.randn(N2, generator=g)
    mid = torch.full((M, N1), float("nan"), dtype=BF)
    out = torch.full((M, N2), float("nan"), dtype=BF)
    counters = torch.zeros(64, dtype=torch.int32)
    emu.emu_gemm_chain(P(x), LL(K1), P(w1), P(b1), act, P(mid), LL(N1), P(w2), P(b2), P(out), LL(N2), M, N1, K1, N2, P(counters), 2)
    want_mid = ref.gemm(x, w1, bias=b1, act=act)
    want = ref.gemm(want_mid, w2, bias=b2)
    assert not bool(torch.isnan(out.float()).any())
    assert float((mid.float() - want_mid.float()).abs().max() / want_mid.float().abs().max()) < 8e-3
    assert float((out.float() - want.float()).abs().max() / want.float().abs().max()) < 1.2e-2
    assert int(counters.abs().sum()) == 0          # the last CTA re-zeroed the dataflow counters
    # the chained launch is the two plain launches fused: same bits
    mid2 = torch.full((M, N1), float("nan"), dtype=BF)
    out2 = torch.full((M, N2), float("nan"), dtype=BF)
    emu.emu_gemm(P(x), LL(K1), P(w1), LL(K1), P(mid2), LL(N1), M, N1, K1, P(b1), None, LL(0), 0, act, 0, 256, 1, 1, 2)
    emu.emu_gemm(P(mid2), LL(N1), P(w2), LL(N1), P(out2), LL(N2), M, N2, N1, P(b2), None, LL(0), 0, 0, 0, 256, 1, 1, 2)
    assert torch.equal(mid, mid2) and torch.equal(out, out2)


@pytest.mark.parametrize("M,N,K,bn,pair,ctas,mode,act", [
    (700, 1152, 264, 256, True, 4, 2, 1),      # LayerNorm + GELU consumer, 3 row blocks x 5 column blocks on 2 pairs: 7.5 tiles each
    (600, 512, 200, 128, True, 6, 1, 0),       # RMSNorm consumer, 12 tiles on 3 pairs: whole tiles only (no cut)
    (333, 640, 136, 128, False, 4, 2, 0),      # one CTA per tile, 15 tiles on 4 units
])
def test_pure_stream_k_for_norm_consumers_is_bit_identical(emu, M, N, K, bn, pair, ctas, mode, act):
    """lp_gemm_set_streamk mode 2: the whole (row block, column block, k) space cut into one contiguous range per unit over a
    column-block-fastest raster — same bits as the plain schedule, for the folded LayerNorm / RMSNorm epilogues."""
    g = torch.Generator().manual_seed(M + N)
    x = (torch.randn((M, K), generator=g) * 1.3 + 0.2).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g) * 0.1
    colsum = w.float().sum(1).contiguous()
    stats = _split_stats(x, 6)
    ws = torch.zeros(4096 + ctas * 128 * bn * 4, dtype=torch.uint8)
    outs = []
    for pure in (0, 1):
        out = torch.full((M, N), float("nan"), dtype=BF)
        fx = _Fx(norm_mode=mode, stats_in=stats.data_ptr(), stats_in_parts=6, colsum=colsum.data_ptr(), norm_eps=1e-5)
        emu.emu_gemm_set_streamk(P(ws) if pure else None)
        emu.emu_gemm_set_streamk_pure(pure)
        try:
            _gemm_ex(emu, x, w, out, bias=b, act=act, bn=bn, pair=pair, ctas=ctas, fx=fx)
        finally:
            emu.emu_gemm_set_streamk(None)
            emu.emu_gemm_set_streamk_pure(0)
        outs.append(out)
    assert not bool(torch.isnan(outs[1].float()).any())
    assert torch.equal(outs[0], outs[1])
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0


# ---- transposed (MN-major) operands: the SFT step's dgrad and wgrad without transposed copies ------------------------------------
@pytest.mark.parametrize("M,N,K,a_t,b_t,fp32,bias,sk", [
    (520, 512, 200, False, True, False, False, False),     # dgrad: dX = dY W with W given as [K, N]; ragged M and K
    (512, 256, 330, True, True, True, False, False),       # wgrad: dW = dY^T X, both operands transposed, fp32 output
    (776, 768, 136, True, False, False, True, False),      # A transposed only, bias, M a multiple of 8 but not of the tile
    (600, 1280, 320, False, True, False, False, True),     # with the stream-K tail (15 tiles on 2 pairs)
    (520, 1152, 136, True, True, True, False, False),      # ragged last column block (1152 = 4.5 tiles): full-width MMA on zero fill
    (300, 200, 72, False, True, False, True, False),       # N smaller than one tile, bias
])
def test_transposed_operand_gemm_under_emulation(emu, M, N, K, a_t, b_t, fp32, bias, sk):
    g = torch.Generator().manual_seed(M + N + K + 7)
    a = torch.randn((M, K), generator=g).to(BF)
    w = (torch.randn((N, K), generator=g) * K ** -0.5).to(BF)
    b = torch.randn(N, generator=g) if bias else None
    A = a.t().contiguous() if a_t else a                         # what the caller holds: At [K, M] / Wt [K, N]
    W = w.t().contiguous() if b_t else w
    want = ref.gemm(a, w, bias=b, out_fp32=fp32)
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if fp32 else BF)
    ws = torch.zeros(4096 + 4 * 128 * 256 * 4, dtype=torch.uint8)
    emu.emu_gemm_set_streamk(P(ws) if sk else None)
    try:
        emu.emu_gemm_mn(P(A), LL(A.stride(0)), P(W), LL(W.stride(0)), P(out), LL(N), M, N, K, P(b), int(fp32), int(a_t), int(b_t), 4)
    finally:
        emu.emu_gemm_set_streamk(None)
    assert not bool(torch.isnan(out.float()).any())
    err = float((out.float() - want.float()).abs().max() / want.float().abs().max())
    assert err < (1e-5 if fp32 else 8e-3), err
    # same bits as the K-major kernel on explicitly transposed copies (same accumulation order)
    out2 = torch.full((M, N), float("nan"), dtype=torch.float32 if fp32 else BF)
    emu.emu_gemm(P(a), LL(K), P(w), LL(K), P(out2), LL(N), M, N, K, P(b), None, LL(0), 0, 0, int(fp32), 256, 1, 1, 4)
    assert torch.equal(out, out2)
