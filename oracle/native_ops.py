"""ORACLE (test infrastructure, never on the product path): CPU restatement of the CONTRACT of every native op that
leopard_b200/ops.py wraps — same signatures, same operand layouts (interleaved gate/up granules, fused qkv column
offsets, padded vocabulary, in-place / out= behaviour), arithmetic in fp32 with one rounding to the output dtype, as the
kernels do.  Two uses:
  * tests/test_train_glue.py swaps it in for `leopard_b200.ops` to run the SFT step's host sequencing (train.py) on the CPU
    and compares loss and gradients with autograd over the model oracle (oracle/train.py);
  * tools/gpu_check_train.py compares each CUDA kernel with the function of the same name here.
Only tests/, tools/gpu_check_*.py, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SWIGLU = 0, 1, 2, 3
BF = torch.bfloat16


def _store(y: torch.Tensor, out, dtype=BF):
    if out is None:
        return y.to(dtype)
    out.copy_(y)
    return out


def gemm(a, w, bias=None, residual=None, act=ACT_NONE, out=None, out_fp32=False, res_row_mod=0, block_n=0):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if act == ACT_GELU_TANH:
        y = F.gelu(y, approximate="tanh")
    elif act == ACT_GELU_ERF:
        y = F.gelu(y)
    elif act == ACT_SWIGLU:
        m, n = y.shape
        g = y.view(m, n // 64, 2, 32)
        y = (F.silu(g[:, :, 0]) * g[:, :, 1]).reshape(m, n // 2)
    if residual is not None:
        r = residual.float()
        if res_row_mod:
            r = r[torch.arange(y.shape[0]) % res_row_mod]
        y = y + r
    return _store(y, out, torch.float32 if out_fp32 else BF)


def gemm_ex(a, w, bias=None, out=None, out_fp32=False, a_t=False, b_t=False, **unsupported):
    """lp_gemm_bf16_ex's transposed-operand contract (what train.py uses it for): a_t -> `a` is At [K, M]; b_t -> `w` is Wt [K, N].
    The epilogue fusions of the prefill path are restated next to their tests (tests/test_gemm_kernel_emulated.py), not here."""
    assert not any(v for v in unsupported.values()), f"native_ops.gemm_ex: not restated here: {sorted(unsupported)}"
    return gemm(a.t() if a_t else a, w.t() if b_t else w, bias=bias, out=out, out_fp32=out_fp32)


def attention(qkv, B, T, n_heads, n_kv_heads, head_dim, head_stride, q_off, k_off, v_off, causal, scale=None, out=None,
              variant=0):
    scale = head_dim ** -0.5 if scale is None else scale
    x = qkv.float().view(B, T, -1)
    rep = n_heads // n_kv_heads
    o = torch.empty((B, T, n_heads * head_dim))
    mask = torch.ones(T, T, dtype=torch.bool).triu(1)
    for h in range(n_heads):
        g = h // rep
        q = x[:, :, q_off + h * head_stride: q_off + h * head_stride + head_dim]
        k = x[:, :, k_off + g * head_stride: k_off + g * head_stride + head_dim]
        v = x[:, :, v_off + g * head_stride: v_off + g * head_stride + head_dim]
        s = (q @ k.transpose(-1, -2)) * scale
        if causal:
            s = s.masked_fill(mask, float("-inf"))
        o[:, :, h * head_dim:(h + 1) * head_dim] = torch.softmax(s, -1) @ v
    return _store(o.view(B * T, -1), out)


def rmsnorm(x, gamma, eps, out=None):
    xf = x.float()
    return _store(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * gamma, out)


def _shuffle(y, side):
    m, d = y.shape
    half = side // 2
    g = y.view(m // (side * side), half, 2, half, 2, d)
    return g.permute(0, 1, 3, 2, 4, 5).reshape(m // 4, 4 * d)


def layernorm(x, gamma, beta, eps, shuffle_side=0, out=None):
    y = F.layer_norm(x.float(), (x.shape[1],), gamma, beta, eps)
    return _store(_shuffle(y, shuffle_side) if shuffle_side else y, out)


def layernorm_bwd(x, gamma, dy, eps, dgamma, dbeta, add=None, out=None):
    xf, df = x.float(), dy.float()
    mu = xf.mean(-1, keepdim=True)
    r = torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + eps)
    xh = (xf - mu) * r
    g = df * gamma
    dx = r * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if add is not None:
        dx = dx + add.float()
    dgamma.copy_((df * xh).sum(0))
    dbeta.copy_(df.sum(0))
    return _store(dx, out)


def pixel_unshuffle(d, side):
    m4, d4 = d.shape
    dd, half = d4 // 4, side // 2
    g = d.view(m4 // (half * half), half, half, 2, 2, dd)          # [b, oh, ow, dy, dx, e]
    return g.permute(0, 1, 3, 2, 4, 5).reshape(m4 * 4, dd).clone()  # [b, oh, dy, ow, dx, e]


def rowmod_sum(d, T, out):
    out.copy_(d.float().view(-1, T, d.shape[1]).sum(0))
    return out


def im2col_patch14(pixels, ldo=592):
    n, c, h, w = pixels.shape
    p = pixels.float().view(n, c, h // 14, 14, w // 14, 14).permute(0, 2, 4, 1, 3, 5).reshape(n * (h // 14) * (w // 14), c * 196)
    out = torch.zeros((p.shape[0], ldo), dtype=BF)
    out[:, : p.shape[1]] = p.to(BF)
    return out


def rope_table(pos, inv_freq):
    f = pos.float()[:, None] * inv_freq.float()[None, :]
    return f.cos(), f.sin()


def rope_inplace(x, col0, n_heads, head_dim, cs, sn):
    half = head_dim // 2
    v = x[:, col0: col0 + n_heads * head_dim].float().view(x.shape[0], n_heads, 2, half)
    x1, x2 = v[:, :, 0], v[:, :, 1]
    c, s = cs[:, None, :], sn[:, None, :]
    o = torch.stack([x1 * c - x2 * s, x2 * c + x1 * s], 2).reshape(x.shape[0], n_heads * head_dim)
    x[:, col0: col0 + n_heads * head_dim] = o.to(x.dtype)
    return x


def transpose(x, out=None):
    return _store(x.float().t(), out)


def rmsnorm_bwd(x, gamma, dy, eps, dgamma, add=None, out=None):
    xf, df = x.float(), dy.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    g = df * gamma
    dx = r * g - xf * r.pow(3) * (g * xf).mean(-1, keepdim=True)
    if add is not None:
        dx = dx + add.float()
    dgamma.copy_((df * xf * r).sum(0))
    return _store(dx, out)


def _split_gu(gu):
    m, n2 = gu.shape
    g = gu.float().view(m, n2 // 64, 2, 32)
    return g[:, :, 0].reshape(m, n2 // 2), g[:, :, 1].reshape(m, n2 // 2)


def swiglu_fwd(gu, out=None):
    g, u = _split_gu(gu)
    return _store(F.silu(g) * u, out)


def swiglu_bwd(gu, dm, out=None):
    g, u = _split_gu(gu)
    d = dm.float()
    s = torch.sigmoid(g)
    dg, du = d * u * s * (1 + g * (1 - s)), d * g * s
    m, f = dg.shape
    y = torch.stack([dg.view(m, f // 32, 32), du.view(m, f // 32, 32)], 2).reshape(m, 2 * f)
    return _store(y, out)


def gelu_bwd(z, dy, act, out=None):
    x, d = z.float(), dy.float()
    if act == ACT_GELU_TANH:
        c, a = math.sqrt(2.0 / math.pi), 0.044715
        t = torch.tanh(c * (x + a * x ** 3))
        g = 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * c * (1 + 3 * a * x * x)
    else:
        g = 0.5 * (1 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    return _store(d * g, out)


def colsum(dy, out):
    out.copy_(dy.float().sum(0))
    return out


def ce_fwd_bwd(logits, target, weight, vocab):
    x = logits.float()[:, :vocab]
    lse = torch.logsumexp(x, -1)
    valid = (target >= 0) & (target < vocab)
    picked = x.gather(-1, target.clamp(0, vocab - 1).unsqueeze(-1)).squeeze(-1)
    loss = torch.where(weight != 0, lse - torch.where(valid, picked, torch.zeros(())), torch.zeros(()))
    p = torch.softmax(x, -1)
    onehot = torch.zeros_like(p).scatter_(-1, target.clamp(0, vocab - 1).unsqueeze(-1), valid.float().unsqueeze(-1))
    d = torch.zeros(logits.shape)
    d[:, :vocab] = (p - onehot) * weight.unsqueeze(-1)
    return loss, d.to(BF)


def embed_bwd(d, ids, image_id, dE):
    keep = ids != image_id
    dE.index_add_(0, ids[keep].clamp(min=0, max=dE.shape[0] - 1), d.float()[keep])


def gather_rows(x, idx):
    return x[idx.long()].clone()


def softmax_bwd(scores, dp, causal, scale, P=None, dS=None):
    s = scores.float() * scale
    if causal:
        s = s.masked_fill(torch.ones(s.shape, dtype=torch.bool).triu(1), float("-inf"))
    p = torch.softmax(s, -1)
    g = scale * p * (dp.float() - (p * dp.float()).sum(-1, keepdim=True))
    return _store(p, P), _store(g, dS)


def accumulate(dst, src):
    dst.add_(src.view(dst.shape))
    return dst


def grad_norm(flat_grads, max_norm, partial=None, out2=None):
    norm = math.sqrt(sum(float(g.double().pow(2).sum()) for g in flat_grads))
    c = 1.0
    if max_norm > 0 and max_norm / (norm + 1e-6) < 1.0:
        c = max_norm / (norm + 1e-6)
    if out2 is None:
        out2 = torch.empty(2)
    out2.copy_(torch.tensor([norm, c]))
    return out2


def adamw_step(p, g, m, v, p16, step, lr, beta1, beta2, eps, weight_decay, norm_and_clip):
    if norm_and_clip is not None and not math.isfinite(float(norm_and_clip[0])):
        return                                       # non-finite gradient norm: the update is skipped
    gi = g * (float(norm_and_clip[1]) if norm_and_clip is not None else 1.0)
    m.mul_(beta1).add_(gi, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
    mh = m / (1 - beta1 ** step)
    vh = v / (1 - beta2 ** step)
    p.sub_(lr * (mh / (vh.sqrt() + eps) + weight_decay * p))
    if p16 is not None:
        p16.copy_(p.view(p16.shape))


def attention_bwd(qkv, out, dout, B, T, n_heads, n_kv_heads, head_dim, q_off, k_off, v_off, causal, scale=None, dqkv=None,
                  head_stride=None):
    """dqkv in the qkv column layout (head_stride columns per head, the first head_dim used; padding columns get zeros):
    closed form of oracle/train.py attention_bwd per sequence."""
    from . import train as OT
    scale = head_dim ** -0.5 if scale is None else scale
    hs = head_stride or head_dim
    if dqkv is None:
        dqkv = torch.zeros_like(qkv)
    x = qkv.float().view(B, T, -1)
    do = dout.float().view(B, T, n_heads, head_dim)

    def heads(xb, off, n):
        return xb[:, off: off + n * hs].reshape(T, n, hs)[:, :, :head_dim]

    def put(rows, off, n, g):
        full = torch.zeros((T, n, hs))
        full[:, :, :head_dim] = g
        dqkv[rows, off: off + n * hs] = full.reshape(T, -1).to(dqkv.dtype)

    for b in range(B):
        q, k, v = heads(x[b], q_off, n_heads), heads(x[b], k_off, n_kv_heads), heads(x[b], v_off, n_kv_heads)
        o, lse = OT.attention_fwd(q, k, v, causal, scale)
        dq, dk, dv = OT.attention_bwd(q, k, v, o, lse, do[b], causal, scale)
        rows = slice(b * T, (b + 1) * T)
        put(rows, q_off, n_heads, dq)
        put(rows, k_off, n_kv_heads, dk)
        put(rows, v_off, n_kv_heads, dv)
    return dqkv


# ---- parity mode (csrc/lp_parity.cu): contracts of the fp32 glue kernels ------------------------------------------------------
def _split3(x):
    h = x.to(BF)
    r = x - h.float()
    m = r.to(BF)
    return h, m, (r - m.float()).to(BF)


def parity_split(x, out, layout):
    M, K = x.shape
    h, m, l = _split3(x.float())
    blocks = {0: (h, m, l), 1: (h, h, h, m, m, l), 2: (h, m, l, h, m, h)}[layout]
    for j, b in enumerate(blocks):
        out[:M, j * K:(j + 1) * K] = b
    return out


def parity_norm(x, gamma, beta, eps, rms):
    if rms:
        return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * gamma
    return F.layer_norm(x, (x.shape[1],), gamma, beta, eps)


def parity_act(x, kind):
    if kind == ACT_SWIGLU:
        m, n = x.shape
        g = x.view(m, n // 64, 2, 32)
        return (F.silu(g[:, :, 0]) * g[:, :, 1]).reshape(m, n // 2)
    return F.gelu(x, approximate="tanh") if kind == ACT_GELU_TANH else F.gelu(x)


def parity_add_(y, x, row_mod=0):
    y += x[torch.arange(y.shape[0]) % row_mod] if row_mod else x
    return y


def parity_rope_(x, n_heads, head_dim, cs, sn):
    half = head_dim // 2
    v = x[:, : n_heads * head_dim].reshape(x.shape[0], n_heads, 2, half).clone()
    x1, x2 = v[:, :, 0], v[:, :, 1]
    c, s = cs[:, None, :], sn[:, None, :]
    x[:, : n_heads * head_dim] = torch.stack([x1 * c - x2 * s, x2 * c + x1 * s], 2).reshape(x.shape[0], n_heads * head_dim)
    return x


def parity_softmax_(scores, Tk, scale, causal):
    s = scores[:, :Tk] * scale
    if causal:
        s = s.masked_fill(torch.ones(s.shape, dtype=torch.bool).triu(1), float("-inf"))
    scores[:, :Tk] = torch.softmax(s, -1)
    return scores
