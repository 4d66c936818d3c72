"""-m gpu: every exported kernel against the CPU oracle / plain torch fp32 on the same seeded inputs (through the C ABI).

Tolerances (bf16 outputs): |err| <= 2^-7 * max(|ref|, 1) for single-rounding ops (norms, RoPE, GEMM epilogues),
2e-2 * max(|ref|, 0.5) for attention (bf16 P operand).  Pure data movement (im2col, merge, shuffle) is bit-exact.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from leopard_b200 import ops  # noqa: E402
from oracle import model as om  # noqa: E402  (checker only)

BF = 2.0 ** -7


def _bf(t):
    return t.to(torch.bfloat16)


def _close(got, ref, rel, floor):
    err = (got.float().cpu() - ref.float().cpu()).abs()
    tol = rel * ref.float().cpu().abs().clamp(min=floor)
    assert bool((err <= tol).all()), f"max err {float(err.max()):.5f} (ref absmax {float(ref.abs().max()):.3f})"


@pytest.mark.parametrize("M,N,K,bias,act,res,rmod,fp32,bn", [
    (128, 256, 64, False, 0, False, 0, False, 256),
    (300, 520, 1160, True, 0, True, 0, False, 0),
    (300, 520, 1160, True, 0, True, 0, False, 128),
    (676, 4304, 1152, True, 1, False, 0, False, 0),
    (676, 1152, 4304, True, 0, True, 0, False, 0),
    (1352, 1152, 592, True, 0, True, 676, False, 0),
    (507, 4096, 4608, True, 2, False, 0, False, 0),
    (201, 6144, 4096, False, 0, False, 0, False, 0),
    (201, 28672, 4096, False, 3, False, 0, False, 0),
    (201, 4096, 14336, False, 0, True, 0, False, 0),
    (3, 128256, 4096, False, 0, False, 0, True, 0),
    (1, 1024, 4096, False, 0, False, 0, True, 0),
])
def test_gemm_epilogues(native_lib, M, N, K, bias, act, res, rmod, fp32, bn):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = _bf(torch.randn(M, K, generator=g))
    w = _bf(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = _bf(torch.randn(N, generator=g) * 0.5).float() if bias else None
    n_out = N // 2 if act == 3 else N
    r = _bf(torch.randn(rmod if rmod else M, n_out, generator=g)) if res else None
    y = a.float() @ w.float().t()
    if b is not None:
        y = y + b
    if act == 1:
        y = torch.nn.functional.gelu(y, approximate="tanh")
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    elif act == 3:
        yy = y.view(M, N // 64, 2, 32)
        y = (torch.nn.functional.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(M, N // 2)
    if r is not None:
        y = y + (r.float()[torch.arange(M) % rmod] if rmod else r.float())
    out = ops.gemm(a.cuda(), w.cuda(), bias=b.cuda() if b is not None else None,
                   residual=r.cuda() if r is not None else None, act=act, out_fp32=fp32, res_row_mod=rmod, block_n=bn)
    _close(out, y, 2e-3 if fp32 else BF, 1.0)


def _qkv(B, T, H, HK, dh, hs, seed):
    g = torch.Generator().manual_seed(seed)
    wq, wk = H * hs, HK * hs
    qkv = torch.zeros(B, T, wq + 2 * wk)
    q, k, v = torch.randn(B, T, H, dh, generator=g), torch.randn(B, T, HK, dh, generator=g), torch.randn(B, T, HK, dh, generator=g)
    q[:, ::5] *= 3.0
    qkv[..., :wq].view(B, T, H, hs)[..., :dh] = q
    qkv[..., wq:wq + wk].view(B, T, HK, hs)[..., :dh] = k
    qkv[..., wq + wk:].view(B, T, HK, hs)[..., :dh] = v
    if hs > dh:
        qkv[..., wq + wk:].view(B, T, HK, hs)[..., dh:] = 5.0  # v pad must be ignored
    return _bf(qkv).view(B * T, -1), wq, wk


def _attn_ref(qkv, B, T, H, HK, dh, hs, wq, wk, causal):
    x = qkv.float().view(B, T, -1)
    q = x[..., :wq].view(B, T, H, hs)[..., :dh].transpose(1, 2)
    k = x[..., wq:wq + wk].view(B, T, HK, hs)[..., :dh].transpose(1, 2).repeat_interleave(H // HK, 1)
    v = x[..., wq + wk:].view(B, T, HK, hs)[..., :dh].transpose(1, 2).repeat_interleave(H // HK, 1)
    s = (q @ k.transpose(-1, -2)) * dh ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool).triu(1), float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, H * dh)


@pytest.mark.parametrize("B,T,H,HK,dh,hs,causal", [
    (1, 1, 4, 1, 128, 128, True),        # single token
    (1, 128, 4, 1, 128, 128, True),
    (2, 300, 8, 2, 128, 128, True),      # ragged tail + batch
    (1, 201, 32, 8, 128, 128, True),     # cfg1 decoder shape
    (1, 777, 8, 2, 128, 128, False),
    (1, 676, 16, 16, 72, 80, False),     # one SigLIP crop
    (3, 676, 16, 16, 72, 80, False),
    (2, 100, 16, 16, 72, 80, False),
])
@pytest.mark.parametrize("variant", [0, 1, 2, 18, 64, 80, 84, 128, 336])
def test_attention(native_lib, B, T, H, HK, dh, hs, causal, variant):
    qkv, wq, wk = _qkv(B, T, H, HK, dh, hs, seed=T + H)
    ref = _attn_ref(qkv, B, T, H, HK, dh, hs, wq, wk, causal)
    out = ops.attention(qkv.cuda(), B, T, H, HK, dh, hs, 0, wq, wq + wk, causal, variant=variant)
    _close(out, ref, 2e-2, 0.5)


@pytest.mark.parametrize("B,T,H,HK,dh,hs,causal", [
    (5, 676, 16, 16, 72, 80, False),     # 240 items on 148 persistent CTAs: item loop, Q replacement, accumulator hand-back
    (2, 1024, 16, 16, 72, 80, False),    # T a multiple of 256
    (3, 420, 4, 4, 72, 80, False),       # fewer items than SMs
    (2, 300, 8, 2, 128, 128, True),      # not eligible for the persistent form: falls through to v3
])
def test_attention_scheduling_variants_are_bit_identical(native_lib, B, T, H, HK, dh, hs, causal):
    """Variant 128 (one softmax warpgroup, two rows per thread) and 336 (persistent CTAs for short non-causal sequences) reorganise
    WHO does the work and WHEN; every row sees v3's arithmetic, so the outputs must match variant 80 bit for bit."""
    qkv, wq, wk = _qkv(B, T, H, HK, dh, hs, seed=T + H + 1)
    q = qkv.cuda()
    base = ops.attention(q, B, T, H, HK, dh, hs, 0, wq, wq + wk, causal, variant=80)
    for variant in (128, 336, 336 | 32):
        out = ops.attention(q, B, T, H, HK, dh, hs, 0, wq, wq + wk, causal, variant=variant)
        assert torch.equal(out, base), f"variant {variant}"


@pytest.mark.parametrize("M,D", [(5, 1152), (676, 1152), (3, 4096), (8, 64)])
def test_layernorm_rmsnorm(native_lib, M, D):
    g = torch.Generator().manual_seed(M + D)
    x = _bf(torch.randn(M, D, generator=g) * 3 + 0.5)
    gm = _bf(1 + 0.1 * torch.randn(D, generator=g)).float()
    bt = _bf(0.1 * torch.randn(D, generator=g)).float()
    ref = torch.nn.functional.layer_norm(x.float(), (D,), gm, bt, 1e-6)
    _close(ops.layernorm(x.cuda(), gm.cuda(), bt.cuda(), 1e-6), ref, BF, 1.0)
    ref = om.rmsnorm(x.float(), gm, 1e-5)
    _close(ops.rmsnorm(x.cuda(), gm.cuda(), 1e-5), ref, BF, 1.0)


def test_layernorm_pixel_shuffle_store(native_lib):
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(2 * 676, 1152, generator=g))
    gm = _bf(1 + 0.1 * torch.randn(1152, generator=g)).float()
    bt = _bf(0.1 * torch.randn(1152, generator=g)).float()
    ln = torch.nn.functional.layer_norm(x.float(), (1152,), gm, bt, 1e-6).view(2, 676, 1152)
    ref = om.pixel_shuffle(ln).reshape(2 * 169, 4608)
    out = ops.layernorm(x.cuda(), gm.cuda(), bt.cuda(), 1e-6, shuffle_side=26)
    _close(out, ref, BF, 1.0)
    # standalone shuffle is pure data movement: bit-exact
    out2 = ops.pixel_shuffle(x.cuda(), 26).cpu()
    assert torch.equal(out2, om.pixel_shuffle(x.view(2, 676, 1152)).reshape(2 * 169, 4608))


def test_rope_matches_oracle(native_lib):
    from leopard_b200.config import leopard_llava, leopard_idefics2_decoder
    from leopard_b200.native_weights import rope_inv_freq
    for dec in (leopard_llava().dec, leopard_idefics2_decoder()):
        inv = rope_inv_freq(dec)
        assert torch.allclose(inv, om.rope_inv_freq(dec), rtol=1e-6, atol=0)
        S = 333
        pos = torch.arange(S, dtype=torch.int32) * 29 % 9000  # positions beyond the original 8192 context too
        g = torch.Generator().manual_seed(5)
        x = _bf(torch.randn(S, 40 * 128 + 1024, generator=g))
        cos, sin = om.rope_cos_sin(dec, pos)
        ref = x.float().clone()
        ref[:, : 40 * 128] = om.apply_rope(ref[:, : 40 * 128].view(S, 40, 128), cos, sin).reshape(S, -1)
        cs, sn = ops.rope_table(pos.cuda(), inv.cuda())
        assert torch.allclose(cs.cpu(), cos[:, :64], atol=2e-3)  # fp32 angle up to ~9000 rad: sin/cos argument rounding
        xg = x.cuda()
        ops.rope_inplace(xg, 0, 40, 128, cs, sn)
        _close(xg, ref, 2 * BF, 1.0)
        assert torch.equal(xg[:, 40 * 128:].cpu(), x[:, 40 * 128:])  # v columns untouched


def test_im2col_exact(native_lib):
    g = torch.Generator().manual_seed(9)
    pix = torch.rand(3, 3, 364, 364, generator=g) * 2 - 1
    ref = torch.nn.functional.unfold(_bf(pix).float(), kernel_size=14, stride=14).transpose(1, 2).reshape(3 * 676, 588)
    for t in (pix, _bf(pix)):
        out = ops.im2col_patch14(t.cuda()).cpu()
        assert torch.equal(out[:, :588].float(), ref)
        assert bool((out[:, 588:] == 0).all())


def test_embed_merge_both_conventions(native_lib):
    g = torch.Generator().manual_seed(11)
    D, Vv = 4096, 2000
    table = _bf(torch.randn(Vv, D, generator=g))
    feats = _bf(torch.randn(3 * 169, D, generator=g))
    # HF expand: one placeholder per crop
    ids = torch.randint(0, 1900, (40,), generator=g)
    ids[[4, 5, 30]] = 1950
    out, n_slots, S = ops.embed_merge(ids.cuda(), table.cuda(), feats.cuda(), 1950, 169)
    assert (n_slots, S) == (3, 40 + 3 * 168)
    rows, k = [], 0
    for t in ids.tolist():
        if t == 1950:
            rows.append(feats[k * 169:(k + 1) * 169]); k += 1
        else:
            rows.append(table[t:t + 1])
    assert torch.equal(out[:S].cpu(), torch.cat(rows))
    # Megatron scatter: one -200 per feature row, -100 pads read row 0
    ids2 = torch.randint(0, 1900, (600,), generator=g)
    ids2[10:10 + 338] = -200
    ids2[400:569] = -200
    ids2[590:] = -100
    out2, n2, S2 = ops.embed_merge(ids2.cuda(), table.cuda(), feats.cuda(), -200, 1)
    ref2 = table[ids2.clamp(min=0)].clone()
    ref2[ids2 == -200] = feats
    assert (n2, S2) == (507, 600) and torch.equal(out2[:600].cpu(), ref2)
    # text only
    out3, n3, S3 = ops.embed_merge(ids[:4].cuda(), table.cuda(), None, 1950, 169)
    assert n3 == 0 and torch.equal(out3[:4].cpu(), table[ids[:4]])


def test_bad_arguments_report_errors(native_lib):
    from leopard_b200._lib import NativeError
    a = torch.zeros(8, 60, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(NativeError, match="multiples of 8"):
        ops.gemm(a[:, :60], torch.zeros(16, 60, dtype=torch.bfloat16, device="cuda"))
    with pytest.raises(NativeError, match="head_dim"):
        ops.attention(torch.zeros(4, 192, dtype=torch.bfloat16, device="cuda"), 1, 4, 1, 1, 64, 64, 0, 64, 128, True)


# ---- lp_gemm_bf16_ex: epilogue fusions (row statistics -> folded LayerNorm / RMSNorm, RoPE) ---------------------------------
def test_gemm_ex_row_statistics_and_folded_layernorm_chain(native_lib):
    """Producer GEMM (bias + residual) writes per-row (sum, sum of squares) partials; the consumer GEMM normalises its A rows
    from them with the LayerNorm gain folded into the weight: == layer_norm(x) @ W^T + b (+ GELU) on the same bf16 values."""
    from leopard_b200 import ops
    g = torch.Generator().manual_seed(3)
    M, K0, D, N, eps = 1400, 592, 1152, 4304, 1e-6
    a = torch.randn((M, K0), generator=g).to(torch.bfloat16).cuda()
    w0 = (torch.randn((D, K0), generator=g) * K0 ** -0.5).to(torch.bfloat16).cuda()
    b0 = (torch.randn(D, generator=g) * 0.1).cuda()
    res = torch.randn((M, D), generator=g).to(torch.bfloat16).cuda()
    x, stats = ops.gemm_ex(a, w0, bias=b0, residual=res, want_stats=True)
    xf = x.float()
    tot = stats.sum(1)
    assert torch.allclose(tot[:, 0], xf.sum(1), rtol=1e-4, atol=2e-3) and torch.allclose(tot[:, 1], (xf * xf).sum(1), rtol=1e-4)
    assert torch.equal(x, ops.gemm(a, w0, bias=b0, residual=res))          # statistics do not change the outputs
    w = (torch.randn((N, D), generator=g) * D ** -0.5).to(torch.bfloat16).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    b = (torch.randn(N, generator=g) * 0.1).cuda()
    want = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xf, (D,), gamma, beta, eps) @ w.float().t() + b,
                                    approximate="tanh")
    wfold = (w.float() * gamma[None]).to(torch.bfloat16)
    got = ops.gemm_ex(x, wfold, bias=w.float() @ beta + b, act=ops.ACT_GELU_TANH, norm_mode=ops.NORM_LAYER, stats_in=stats,
                      colsum=wfold.float().sum(1).contiguous(), eps=eps)
    assert float((got.float() - want).abs().max() / want.abs().max()) < 1.5e-2
    # lp_row_stats is the stand-alone producer: same consumer result up to summation order
    got2 = ops.gemm_ex(x, wfold, bias=w.float() @ beta + b, act=ops.ACT_GELU_TANH, norm_mode=ops.NORM_LAYER,
                       stats_in=ops.row_stats(x), colsum=wfold.float().sum(1).contiguous(), eps=eps)
    assert float((got2.float() - got.float()).abs().max()) < 2e-2 * float(want.abs().max())


def test_gemm_ex_rmsnorm_swiglu_and_rope(native_lib):
    from leopard_b200 import ops
    g = torch.Generator().manual_seed(4)
    M, D, F, eps = 700, 4096, 1024, 1e-5
    x = torch.randn((M, D), generator=g).to(torch.bfloat16).cuda()
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    xf = x.float()
    rs = torch.rsqrt((xf * xf).mean(-1, keepdim=True) + eps)
    stats = ops.row_stats(x)
    # gate/up with SwiGLU on 32-row granules
    wgu = (torch.randn((2 * F, D), generator=g) * D ** -0.5).to(torch.bfloat16).cuda()
    y = (xf * rs * gamma) @ wgu.float().t()
    yy = y.view(M, 2 * F // 64, 2, 32)
    want = (torch.nn.functional.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(M, F)
    got = ops.gemm_ex(x, (wgu.float() * gamma[None]).to(torch.bfloat16), act=ops.ACT_SWIGLU, norm_mode=ops.NORM_RMS, stats_in=stats,
                      eps=eps)
    assert float((got.float() - want).abs().max() / want.abs().max()) < 1.5e-2
    # QKV with RoPE on 6 of 8 heads
    heads, rot = 8, 6
    wq = (torch.randn((heads * 128, D), generator=g) * D ** -0.5).to(torch.bfloat16).cuda()
    pos = torch.arange(M, dtype=torch.int32).cuda()
    inv = (1.0 / (10000.0 ** (torch.arange(0, 128, 2).float() / 128))).cuda()
    cos, sin = ops.rope_table(pos, inv)
    yq = ((xf * rs * gamma) @ wq.float().t()).view(M, heads, 128)
    lo, hi = yq[..., :64], yq[..., 64:]
    rotd = torch.cat([lo * cos[:, None] - hi * sin[:, None], hi * cos[:, None] + lo * sin[:, None]], -1)
    want = torch.cat([rotd[:, :rot], yq[:, rot:]], 1).reshape(M, heads * 128)
    got = ops.gemm_ex(x, (wq.float() * gamma[None]).to(torch.bfloat16), norm_mode=ops.NORM_RMS, stats_in=stats, eps=eps,
                      rope=(cos, sin), rope_heads=rot)
    assert float((got.float() - want).abs().max() / want.abs().max()) < 1.5e-2
    # == the unfused kernels (norm -> GEMM -> in-place RoPE) within bf16 rounding of the intermediate
    h = ops.rmsnorm(x, gamma, eps)
    ref2 = ops.gemm(h, wq)
    ops.rope_inplace(ref2, 0, rot, 128, cos, sin)
    assert float((got.float() - ref2.float()).abs().max() / want.abs().max()) < 2e-2


def test_gemm_raster_balance_does_not_change_a_bit(native_lib):
    """lp_gemm_set_raster_balance only changes which persistent unit runs which tile (ViT fc2: W L2-resident, A not, ragged last
    column block): same bits either way, with and without the residual ring."""
    from leopard_b200 import _lib
    L = _lib.lib()
    assert _lib.ensure_streamk_workspace("cuda:0") is not None
    M, N, K = 18928, 1152, 4304
    g = torch.Generator(device="cuda").manual_seed(7)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    bias = torch.randn(N, device="cuda", generator=g) * 0.5
    r = _bf(torch.randn(M, N, device="cuda", generator=g))
    outs = []
    try:
        for bal in (0, 1):
            L.lp_gemm_set_raster_balance(bal)
            x = r.clone()
            outs.append(ops.gemm_ex(a, w, bias=bias, residual=x, act=0, out=x))
    finally:
        L.lp_gemm_set_raster_balance(1)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    rows = torch.arange(0, M, 97, device="cuda")
    _close(outs[1][rows], a[rows].float() @ w.float().t() + bias + r[rows].float(), BF, 1.0)


@pytest.mark.parametrize("M,N,K,res,act,stats", [
    (4796, 4096, 4096, True, 0, True),       # decoder o-proj at cfg3: 304 tiles on 74 CTA pairs -> stream-K tail, residual by TMA, row statistics
    (4796, 6144, 4096, False, 0, False),     # qkv: 456 tiles
    (2937, 4096, 14336, True, 0, False),     # down-proj at cfg2: 192 tiles (2.6 waves), K = 224 k-blocks
    (2937, 28672, 4096, False, 3, False),    # gate/up at cfg2 with SwiGLU: 1344 tiles (18.2 waves)
    (18928, 1152, 1152, True, 0, True),      # ViT out-proj: no tail (74 x 5 tiles), residual ring only
    (4796, 1536, 2048, True, 0, False),      # 114 tiles on 74 pairs: fewer than two waves, every tile belongs to the stream-K range
])
def test_gemm_stream_k_and_tma_residual_do_not_change_a_bit(native_lib, M, N, K, res, act, stats):
    """lp_gemm_set_streamk / lp_gemm_set_tma_residual are pure scheduling / data-path switches: every combination must give
    the bits of the plain schedule (and those must match fp32 torch)."""
    from leopard_b200 import _lib
    L = _lib.lib()
    assert _lib.ensure_streamk_workspace("cuda:0") is not None
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    bias = torch.randn(N, device="cuda", generator=g) * 0.5 if act == 0 else None
    n_out = N // 2 if act == 3 else N
    r = _bf(torch.randn(M, n_out, device="cuda", generator=g)) if res else None
    outs = {}
    try:
        for sk in (0, 1):
            for tma in (0, 1):
                L.lp_gemm_set_streamk(sk)
                L.lp_gemm_set_tma_residual(tma)
                x = r.clone() if res else None        # in place, as the layer loops use it
                o = ops.gemm_ex(a, w, bias=bias, residual=x, act=act, out=x, block_n=256, want_stats=stats)
                outs[(sk, tma)] = o if stats else (o, None)
    finally:
        L.lp_gemm_set_streamk(1)
        L.lp_gemm_set_tma_residual(1)
    torch.cuda.synchronize()
    base, base_st = outs[(0, 0)]
    ref = a.float() @ w.float().t()
    if bias is not None:
        ref = ref + bias
    if act == 3:
        yy = ref.view(M, N // 64, 2, 32)
        ref = (torch.nn.functional.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(M, N // 2)
    if res:
        ref = ref + r.float()
    _close(base, ref, BF, 1.0)
    for key, (o, st) in outs.items():
        assert torch.equal(o, base), f"streamk={key[0]} tma_residual={key[1]} changed the output"
        if stats:
            assert torch.equal(st, base_st), f"streamk={key[0]} tma_residual={key[1]} changed the row statistics"


@pytest.mark.parametrize("M,N,K,a_t,b_t,fp32", [
    (4096, 4096, 14336, False, True, False),     # dgrad of the down projection: dX = dY W_down, W given as [K, N]
    (4096, 14336, 4096, True, True, True),       # wgrad of gate / up (half of it): dW = dY^T X, fp32 output, both operands transposed
    (6144, 4096, 4104, True, True, True),        # wgrad of qkv, ragged contraction length (4104 tokens)
    (4796, 4096, 4096, False, True, False),      # ragged M
    (2048, 4096, 1000, True, False, False),      # A transposed only
    (4304, 1152, 14196, True, True, True),       # vision tower fc1 wgrad at 21 crops: ragged last column block, ragged contraction
    (14196, 4304, 1152, False, True, False),     # vision tower fc2 dgrad: N = 4304 (16.8 tiles)
])
def test_gemm_transposed_operands_match_explicit_transposes_bit_for_bit(native_lib, M, N, K, a_t, b_t, fp32):
    """lp_gemm_desc a_t / b_t (MN-major tensor-core operands straight from the row-major tensors, what the SFT step's dgrad and
    wgrad use) against fp32 torch, and bit for bit against the same kernel fed explicit transposed copies."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) * K ** -0.5)
    A = a.t().contiguous() if a_t else a
    W = w.t().contiguous() if b_t else w
    got = ops.gemm_ex(A, W, out_fp32=fp32, a_t=a_t, b_t=b_t)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if fp32:
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5
    else:
        _close(got, ref, BF, 1.0)
    if K % 8 == 0:      # the K-major kernel needs 16-byte row strides; the transposed operands have none along K
        assert torch.equal(got, ops.gemm(a, w, out_fp32=fp32, block_n=256))
