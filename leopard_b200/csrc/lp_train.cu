// leopard_b200 — HBM-bound kernels of the SFT step (backward of the glue ops, masked LM loss, optimizer).
//
// Specification: oracle/train.py (closed forms pinned to autograd on the CPU).  Reference call sites:
//   loss           Pai-Megatron-Patch/examples/llava/pretrain_megatron_llava.py:101-121 (shifted, weighted CE)
//   norms / MLP    megatron/legacy/model/rms_norm.py:26-31, megatron_patch/model/llava/transformer.py:160-176 (autograd there)
//   optimizer      Megatron-LM-240424/megatron/core/optimizer/optimizer.py:300-370, clip_grads.py:103-141, apex FusedAdam
// The dgrad / wgrad contractions run on lp_gemm_bf16 over operands re-laid out by transpose_kernel; everything here is
// element-wise or a row / column reduction: fp32 math, 16-byte accesses where the layout allows, deterministic
// reductions (per-block partials + a second pass) except the embedding-row scatter, which uses fp32 atomics.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

// ---- helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// 256-thread block reductions; `red` is 8 floats of shared memory, safe to reuse right after the call returns
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  __syncthreads();
  return t;
}
__device__ __forceinline__ float block_max_256(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]);
  __syncthreads();
  return t;
}

// ---- transpose: out[c, r] = in[r, c] (bf16), 64 x 64 tiles through shared memory --------------------------------
// Feeds lp_gemm_bf16 (which contracts the CONTIGUOUS dimension of both operands) with W^T for dgrad and with dY^T, X^T
// for wgrad.  Both sides move 4-byte pairs; 128-byte rows per warp on the load and on the store.
__global__ void __launch_bounds__(256)
transpose_kernel(const bf16* __restrict__ in, long long ld_in, int rows, int cols, bf16* __restrict__ out, long long ld_out) {
  __shared__ bf16 tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    bf16 a = __float2bfloat16(0.f), b = a;
    if (r < rows) {
      if (c + 1 < cols) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(in + (long long)r * ld_in + c);
        a = v.x;
        b = v.y;
      } else if (c < cols) {
        a = in[(long long)r * ld_in + c];
      }
    }
    tile[ty + 8 * i][2 * tx] = a;
    tile[ty + 8 * i][2 * tx + 1] = b;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int oc = ty + 8 * i;            // input column = output row
    const int c = c0 + oc, r = r0 + 2 * tx;
    if (c < cols) {
      if (r + 1 < rows) {
        __nv_bfloat162 v;
        v.x = tile[2 * tx][oc];
        v.y = tile[2 * tx + 1][oc];
        *reinterpret_cast<__nv_bfloat162*>(out + (long long)c * ld_out + r) = v;
      } else if (r < rows) {
        out[(long long)c * ld_out + r] = tile[2 * tx][oc];
      }
    }
  }
}

// ---- RMSNorm backward ---------------------------------------------------------------------------------------------
// y = x r g, r = rsqrt(mean(x^2) + eps):   dx = r (dy g) - x r^3 mean(dy g x) (+ add);   dg = sum_rows dy x r.
// One CTA walks rows blockIdx.x, + gridDim.x, ...; thread t owns the 8-column chunks t, t + 256, ... (J of them), so
// the per-column dg sums stay in registers across rows and leave as one fp32 partial row per CTA.
constexpr int NB_PARTIAL = 296;   // CTAs of the row-walking backward kernels = rows of their partial buffers

template <int J>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ gamma, const bf16* __restrict__ dy,
                   long long lddy, const bf16* __restrict__ add, long long ldadd, bf16* __restrict__ dx, long long lddx,
                   float* __restrict__ partial, int M, int D, float eps) {
  __shared__ float red[2][2][8];
  const int t = threadIdx.x;
  float gm[J][8], acc[J][8];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = (t + 256 * j) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gm[j][e] = (c < D) ? gamma[c + e] : 0.f;
      acc[j][e] = 0.f;
    }
  }
  int par = 0;
  for (int r = blockIdx.x; r < M; r += gridDim.x, par ^= 1) {
    uint4 xu[J], du[J];
    float ss = 0.f, sg = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = (t + 256 * j) * 8;
      xu[j] = make_uint4(0, 0, 0, 0);
      du[j] = make_uint4(0, 0, 0, 0);
      if (c < D) {
        xu[j] = *reinterpret_cast<const uint4*>(x + (long long)r * ldx + c);
        du[j] = *reinterpret_cast<const uint4*>(dy + (long long)r * lddy + c);
      }
      float xf[8], df[8];
      unpack8(xu[j], xf);
      unpack8(du[j], df);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        ss += xf[e] * xf[e];
        sg += df[e] * gm[j][e] * xf[e];
      }
    }
    ss = warp_sum(ss);
    sg = warp_sum(sg);
    if ((t & 31) == 0) {
      red[par][0][t >> 5] = ss;
      red[par][1][t >> 5] = sg;
    }
    __syncthreads();   // one barrier per row: the other parity's slots are free by the time they are rewritten
    float tss = 0.f, tsg = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      tss += red[par][0][w];
      tsg += red[par][1][w];
    }
    const float rinv = rsqrtf(tss / (float)D + eps);
    const float coef = rinv * rinv * rinv * tsg / (float)D;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = (t + 256 * j) * 8;
      if (c < D) {
        float xf[8], df[8], o[8];
        unpack8(xu[j], xf);
        unpack8(du[j], df);
        float af[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (add != nullptr) {
          const uint4 au = *reinterpret_cast<const uint4*>(add + (long long)r * ldadd + c);
          unpack8(au, af);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = rinv * df[e] * gm[j][e] - xf[e] * coef + af[e];
          acc[j][e] += df[e] * xf[e] * rinv;
        }
        *reinterpret_cast<uint4*>(dx + (long long)r * lddx + c) = pack8(o);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = (t + 256 * j) * 8;
    if (c < D) {
      float4* p = reinterpret_cast<float4*>(partial + (long long)blockIdx.x * D + c);
      p[0] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      p[1] = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
    }
  }
}

// ---- LayerNorm backward (vision tower) ------------------------------------------------------------------------------
// y = xh g + b, xh = (x - mu) r:  dx = r (gh - mean(gh) - xh mean(gh xh)) (+ add), gh = dy g;  dg = sum dy xh;  db = sum dy.
// Same walk as rmsnorm_bwd_kernel; statistics two-pass (mean, then centred variance) like the forward kernel.  The
// partial buffer holds [dg | db] rows of width 2 D per CTA.
template <int J>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ gamma, const bf16* __restrict__ dy,
                     long long lddy, const bf16* __restrict__ add, long long ldadd, bf16* __restrict__ dx, long long lddx,
                     float* __restrict__ partial, int M, int D, float eps) {
  __shared__ float red[8];
  const int t = threadIdx.x;
  float gm[J][8], accg[J][8], accb[J][8];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = (t + 256 * j) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gm[j][e] = (c < D) ? gamma[c + e] : 0.f;
      accg[j][e] = 0.f;
      accb[j][e] = 0.f;
    }
  }
  const float invD = 1.f / (float)D;
  for (int r = blockIdx.x; r < M; r += gridDim.x) {
    uint4 xu[J], du[J];
    float sx = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = (t + 256 * j) * 8;
      xu[j] = make_uint4(0, 0, 0, 0);
      du[j] = make_uint4(0, 0, 0, 0);
      if (c < D) {
        xu[j] = *reinterpret_cast<const uint4*>(x + (long long)r * ldx + c);
        du[j] = *reinterpret_cast<const uint4*>(dy + (long long)r * lddy + c);
        float xf[8];
        unpack8(xu[j], xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) sx += xf[e];
      }
    }
    const float mu = block_sum_256(sx, red) * invD;
    float sv = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = (t + 256 * j) * 8;
      if (c < D) {
        float xf[8];
        unpack8(xu[j], xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) sv += (xf[e] - mu) * (xf[e] - mu);
      }
    }
    const float rinv = rsqrtf(block_sum_256(sv, red) * invD + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = (t + 256 * j) * 8;
      if (c < D) {
        float xf[8], df[8];
        unpack8(xu[j], xf);
        unpack8(du[j], df);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float gh = df[e] * gm[j][e];
          sg += gh;
          sgx += gh * (xf[e] - mu) * rinv;
        }
      }
    }
    const float mg = block_sum_256(sg, red) * invD;
    const float mgx = block_sum_256(sgx, red) * invD;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = (t + 256 * j) * 8;
      if (c < D) {
        float xf[8], df[8], o[8];
        unpack8(xu[j], xf);
        unpack8(du[j], df);
        float af[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (add != nullptr) unpack8(*reinterpret_cast<const uint4*>(add + (long long)r * ldadd + c), af);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xf[e] - mu) * rinv;
          o[e] = rinv * (df[e] * gm[j][e] - mg - xh * mgx) + af[e];
          accg[j][e] += df[e] * xh;
          accb[j][e] += df[e];
        }
        *reinterpret_cast<uint4*>(dx + (long long)r * lddx + c) = pack8(o);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = (t + 256 * j) * 8;
    if (c < D) {
      float4* pg = reinterpret_cast<float4*>(partial + (long long)blockIdx.x * 2 * D + c);
      float4* pb = reinterpret_cast<float4*>(partial + (long long)blockIdx.x * 2 * D + D + c);
      pg[0] = make_float4(accg[j][0], accg[j][1], accg[j][2], accg[j][3]);
      pg[1] = make_float4(accg[j][4], accg[j][5], accg[j][6], accg[j][7]);
      pb[0] = make_float4(accb[j][0], accb[j][1], accb[j][2], accb[j][3]);
      pb[1] = make_float4(accb[j][4], accb[j][5], accb[j][6], accb[j][7]);
    }
  }
}

// ---- pixel-shuffle backward: x_grad[b, (2oh+dy)*side + 2ow+dx, :] = d[b, oh*half + ow, (dy*2+dx)*D : +D]; one CTA per row of x
__global__ void __launch_bounds__(128)
pixel_unshuffle_kernel(const bf16* __restrict__ d, long long ldd, bf16* __restrict__ out, long long ldo, int side, int D) {
  const int row = blockIdx.x;
  const int per = side * side, half = side >> 1;
  const int b = row / per, t = row % per, h = t / side, w = t % side;
  const long long srow = (long long)b * (half * half) + (h >> 1) * half + (w >> 1);
  const uint4* sp = reinterpret_cast<const uint4*>(d + srow * ldd + (((h & 1) << 1) | (w & 1)) * D);
  uint4* dp = reinterpret_cast<uint4*>(out + (long long)row * ldo);
  for (int c = threadIdx.x; c < (D >> 3); c += blockDim.x) dp[c] = sp[c];
}

// ---- position-embedding gradient: out[t, c] = sum_n d[n * T + t, c]  (fp32, fixed order) --------------------------------
__global__ void __launch_bounds__(256)
rowmod_sum_kernel(const bf16* __restrict__ d, long long ldd, int n_rep, int T, int D, float* __restrict__ out) {
  const int t = blockIdx.x;
  for (int c = threadIdx.x * 2; c < D; c += 512) {
    float2 s = make_float2(0.f, 0.f);
    for (int n = 0; n < n_rep; ++n) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(d + ((long long)n * T + t) * ldd + c));
      s.x += v.x;
      s.y += v.y;
    }
    out[(long long)t * D + c] = s.x;
    out[(long long)t * D + c + 1] = s.y;
  }
}

// out[c] = sum_b partial[b, c]   (fixed order: deterministic)
__global__ void __launch_bounds__(256)
partial_reduce_kernel(const float* __restrict__ partial, int nb, int D, float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= D) return;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += partial[(long long)b * D + c];
  out[c] = s;
}

// the LayerNorm variant: partial rows are [dg | db] (width 2 D)
__global__ void __launch_bounds__(256)
partial_reduce2_kernel(const float* __restrict__ partial, int nb, int D, float* __restrict__ dg, float* __restrict__ db) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= 2 * D) return;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += partial[(long long)b * 2 * D + c];
  if (c < D) dg[c] = s; else db[c - D] = s;
}

// ---- SwiGLU on the interleaved gate/up layout of the fused weight (32-column granules: gate_0..31, up_0..31, ...) ----
// forward (materialised for the backward's recompute): m = silu(g) u;  backward: dg = dm u s (1 + g (1 - s)), du = dm g s
template <bool BWD>
__global__ void __launch_bounds__(256)
swiglu_kernel(const bf16* __restrict__ gu, long long ldgu, const bf16* __restrict__ dm, long long lddm, bf16* __restrict__ out,
              long long ldo, int M, int F) {
  const long long chunks = (long long)M * (F >> 3);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < chunks; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / (F >> 3)), c = (int)(i % (F >> 3)) * 8;        // m columns c .. c + 7 (inside one granule)
    const int cg = (c >> 5) * 64 + (c & 31);
    const uint4 g4 = *reinterpret_cast<const uint4*>(gu + (long long)r * ldgu + cg);
    const uint4 u4 = *reinterpret_cast<const uint4*>(gu + (long long)r * ldgu + cg + 32);
    float g[8], u[8], o[8];
    unpack8(g4, g);
    unpack8(u4, u);
    if (!BWD) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = silu_f(g[e]) * u[e];
      *reinterpret_cast<uint4*>(out + (long long)r * ldo + c) = pack8(o);
    } else {
      float d[8], o2[8];
      unpack8(*reinterpret_cast<const uint4*>(dm + (long long)r * lddm + c), d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float s = 1.f / (1.f + __expf(-g[e]));
        o[e] = d[e] * u[e] * s * (1.f + g[e] * (1.f - s));
        o2[e] = d[e] * g[e] * s;
      }
      *reinterpret_cast<uint4*>(out + (long long)r * ldo + cg) = pack8(o);
      *reinterpret_cast<uint4*>(out + (long long)r * ldo + cg + 32) = pack8(o2);
    }
  }
}

// ---- GELU backward (projector): dz = dy gelu'(z); act 1 = tanh form, 2 = erf form -----------------------------------
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const bf16* __restrict__ z, long long ldz, const bf16* __restrict__ dy, long long lddy, bf16* __restrict__ dz,
                long long lddz, int M, int N, int act) {
  const long long chunks = (long long)M * (N >> 3);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < chunks; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / (N >> 3)), c = (int)(i % (N >> 3)) * 8;
    float x[8], d[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(z + (long long)r * ldz + c), x);
    unpack8(*reinterpret_cast<const uint4*>(dy + (long long)r * lddy + c), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float g;
      if (act == 1) {
        const float k0 = 0.7978845608028654f, k1 = 0.044715f;
        const float th = tanhf(k0 * (x[e] + k1 * x[e] * x[e] * x[e]));
        g = 0.5f * (1.f + th) + 0.5f * x[e] * (1.f - th * th) * k0 * (1.f + 3.f * k1 * x[e] * x[e]);
      } else {
        g = 0.5f * (1.f + erff(x[e] * 0.7071067811865476f)) + x[e] * 0.3989422804014327f * __expf(-0.5f * x[e] * x[e]);
      }
      o[e] = d[e] * g;
    }
    *reinterpret_cast<uint4*>(dz + (long long)r * lddz + c) = pack8(o);
  }
}

// ---- column sums (bias gradients): out[c] = sum_r dy[r, c]; one CTA per 64 columns, fixed summation order ----------
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ dy, long long ld, int M, int N, float* __restrict__ out) {
  __shared__ float2 red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + 2 * tx;
  float2 s = make_float2(0.f, 0.f);
  if (c < N) {
    for (int r = ty; r < M; r += 8) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + (long long)r * ld + c));
      s.x += v.x;
      s.y += v.y;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < N) {
    float2 t = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      t.x += red[w][tx].x;
      t.y += red[w][tx].y;
    }
    out[c] = t.x;
    if (c + 1 < N) out[c + 1] = t.y;
  }
}

// ---- masked LM loss, forward + gradient in one pass over the logits ------------------------------------------------
// row r: lse = logsumexp(logits[r, :V]); loss[r] = lse - logits[r, target[r]] (lse alone when the target is outside
// [0, V), as Megatron's vocab-parallel CE does); dlogits[r, i] = (exp(logits[r, i] - lse) - [i == target]) * weight[r].
// weight[r] already holds loss_mask / sum(loss_mask); rows with weight 0 get a zero gradient row without touching the
// logits (most rows of an SFT batch: prompt and image positions).  Columns V .. Vp-1 (vocabulary padding) get 0.
template <typename T>
__global__ void __launch_bounds__(256)
ce_fwd_bwd_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ target,
                  const float* __restrict__ weight, float* __restrict__ loss, bf16* __restrict__ dlogits, long long ldd, int V,
                  int Vp) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const T* row = logits + (long long)r * ld;
  bf16* drow = dlogits + (long long)r * ldd;
  const float w = weight[r];
  if (w == 0.f) {
    for (int i = threadIdx.x * 2; i < Vp; i += 512) *reinterpret_cast<uint32_t*>(drow + i) = 0u;
    if (threadIdx.x == 0) loss[r] = 0.f;
    return;
  }
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += 256) mx = fmaxf(mx, (float)row[i]);
  mx = block_max_256(mx, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += 256) s += __expf((float)row[i] - mx);
  s = block_sum_256(s, red);
  const long long tg = target[r];
  const bool valid = tg >= 0 && tg < V;
  if (threadIdx.x == 0) loss[r] = logf(s) + mx - (valid ? (float)row[tg] : 0.f);
  const float inv = w / s;
  for (int i = threadIdx.x * 2; i < Vp; i += 512) {   // Vp and ldd are even: 4-byte stores
    float a = 0.f, b = 0.f;
    if (i < V) a = __expf((float)row[i] - mx) * inv - ((valid && i == tg) ? w : 0.f);
    if (i + 1 < V) b = __expf((float)row[i + 1] - mx) * inv - ((valid && i + 1 == tg) ? w : 0.f);
    *reinterpret_cast<uint32_t*>(drow + i) = pack_bf16x2(a, b);
  }
}

// ---- embedding backward: dE[id] += d[r] for text rows; image slots (id == image_id) were overwritten by features and
// pass nothing; other negative ids (pads) read row 0 in the forward (clamp, vlm_model.py:226) and so feed row 0 --------
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const bf16* __restrict__ d, long long ldd, const long long* __restrict__ ids, long long image_id,
                 float* __restrict__ dE, long long ldE, int D, long long vocab) {
  const int r = blockIdx.x;
  long long id = ids[r];
  if (id == image_id) return;
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;   // the forward gather clamps the same way (merge_copy_kernel)
  for (int c = threadIdx.x * 2; c < D; c += 512) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(d + (long long)r * ldd + c));
    atomicAdd(dE + id * ldE + c, v.x);
    atomicAdd(dE + id * ldE + c + 1, v.y);
  }
}

// ---- row gather: out[i] = in[idx[i]]  (gradient of the image-feature scatter) ---------------------------------------
__global__ void __launch_bounds__(128)
gather_rows_kernel(const bf16* __restrict__ in, long long ld_in, const int* __restrict__ idx, int D, bf16* __restrict__ out,
                   long long ld_out) {
  const uint4* s = reinterpret_cast<const uint4*>(in + (long long)idx[blockIdx.x] * ld_in);
  uint4* o = reinterpret_cast<uint4*>(out + (long long)blockIdx.x * ld_out);
  for (int c = threadIdx.x; c < (D >> 3); c += 128) o[c] = s[c];
}

// ---- softmax backward over materialised score rows (first implementation of attention backward) ---------------------
// sc = raw q k^T (fp32), dp = dO v^T (fp32): P = softmax(scale sc, causal), dS = scale P (dp - sum_j P dp);
// both leave as bf16 with zeros beyond the causal limit (the consumers are full-width GEMMs)
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ sc, long long lds, const float* __restrict__ dp, long long lddp,
                   bf16* __restrict__ P, long long ldp, bf16* __restrict__ dS, long long ldds, int Sk, int causal, float scale) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const int n = causal ? min(r + 1, Sk) : Sk;
  const float* srow = sc + (long long)r * lds;
  const float* drow = dp + (long long)r * lddp;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, srow[i]);
  mx = block_max_256(mx, red);
  float s = 0.f, sd = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float e = __expf(scale * (srow[i] - mx));
    s += e;
    sd += e * drow[i];
  }
  s = block_sum_256(s, red);
  sd = block_sum_256(sd, red);
  const float inv = 1.f / s, delta = sd * inv;
  for (int i = threadIdx.x; i < Sk; i += 256) {
    float p = 0.f, g = 0.f;
    if (i < n) {
      p = __expf(scale * (srow[i] - mx)) * inv;
      g = scale * p * (drow[i] - delta);
    }
    P[(long long)r * ldp + i] = __float2bfloat16(p);
    dS[(long long)r * ldds + i] = __float2bfloat16(g);
  }
}

// ---- gradient norm + clip coefficient + AdamW ------------------------------------------------------------------------
constexpr int NORM_BLOCKS = 1024;

// partial[b] += sum of squares of this CTA's grid-stride share of x (launches are stream-ordered: no atomics needed)
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = x4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 256) s += x[i] * x[i];
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] += s;
}

// out[0] = global L2 norm, out[1] = clip coefficient (clip_grads.py:133-139: max / (norm + 1e-6) when that is < 1, else 1)
__global__ void __launch_bounds__(256)
grad_norm_finish_kernel(const float* __restrict__ partial, int nb, float max_norm, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += (double)partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(red[0]);
    float c = 1.f;
    if (max_norm > 0.f) {
      const float k = max_norm / (norm + 1.0e-6f);
      if (k < 1.f) c = k;
    }
    out[0] = norm;
    out[1] = c;
  }
}

// dst += src (fp32): gradient accumulation over micro-batches (the wgrad GEMM overwrites its output)
__global__ void __launch_bounds__(256)
accumulate_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  const long long n4 = n >> 2;
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 a = d4[i];
    const float4 b = s4[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    d4[i] = a;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 256) dst[i] += src[i];
}

// FusedAdam, adam_w_mode: m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; p -= lr (m / bc1 / (sqrt(v / bc2) + eps) + wd p)
// g is scaled by the clip coefficient read from device memory; the bf16 model copy is refreshed in the same pass.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p16, long long n, float lr, float b1, float b2, float eps, float wd, float inv_bc1,
             float inv_bc2, const float* __restrict__ clip) {
  // a non-finite gradient norm skips the whole update (Megatron's found-inf path: optimizer.py:313-340 returns without
  // stepping); NaN fails the comparison as well
  if (clip != nullptr && !(clip[0] < INFINITY)) return;
  const float cs = clip != nullptr ? clip[1] : 1.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float gi = g[i] * cs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    float pi = p[i];
    pi -= lr * (mi * inv_bc1 / (sqrtf(vi * inv_bc2) + eps) + wd * pi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (p16 != nullptr) p16[i] = __float2bfloat16(pi);
  }
}

// same update, four parameters per thread with 16-byte accesses (30 B of HBM traffic per parameter: the scalar kernel above reached
// 4.3 TB/s of the 6.5 TB/s the part copies at); element-wise arithmetic identical to adamw_kernel
__global__ void __launch_bounds__(256)
adamw_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                  uint2* __restrict__ p16, long long n4, float lr, float b1, float b2, float eps, float wd, float inv_bc1,
                  float inv_bc2, const float* __restrict__ clip) {
  if (clip != nullptr && !(clip[0] < INFINITY)) return;
  const float cs = clip != nullptr ? clip[1] : 1.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 g4 = g[i];
    float4 m4 = m[i], v4 = v[i], p4 = p[i];
#define LP_ADAMW_ONE(C)                                                                   \
    {                                                                                     \
      const float gi = g4.C * cs;                                                         \
      m4.C = b1 * m4.C + (1.f - b1) * gi;                                                 \
      v4.C = b2 * v4.C + (1.f - b2) * gi * gi;                                            \
      p4.C -= lr * (m4.C * inv_bc1 / (sqrtf(v4.C * inv_bc2) + eps) + wd * p4.C);          \
    }
    LP_ADAMW_ONE(x) LP_ADAMW_ONE(y) LP_ADAMW_ONE(z) LP_ADAMW_ONE(w)
#undef LP_ADAMW_ONE
    m[i] = m4;
    v[i] = v4;
    p[i] = p4;
    if (p16 != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      p16[i] = u;
    }
  }
}

}  // namespace lp

using namespace lp;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool al4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }
static unsigned grid_for(long long work_items, int per_block = 256) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

extern "C" int lp_transpose_bf16(const void* in, long long ld_in, int rows, int cols, void* out, long long ld_out,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(in && out && rows > 0 && cols > 0, "lp_transpose_bf16: bad arguments rows=%d cols=%d", rows, cols);
  LP_REQUIRE(al4(in) && al4(out) && ld_in % 2 == 0 && ld_out % 2 == 0 && ld_in >= cols && ld_out >= rows,
             "lp_transpose_bf16: bases must be 4-byte aligned and leading dimensions even (ld_in=%lld ld_out=%lld)", ld_in,
             ld_out);
  note_launch();
  transpose_kernel<<<dim3((cols + 63) / 64, (rows + 63) / 64), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(in), ld_in, rows, cols, reinterpret_cast<bf16*>(out), ld_out);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t lp_rmsnorm_bwd_workspace_bytes(int D) { return (size_t)NB_PARTIAL * (size_t)D * sizeof(float); }

extern "C" int lp_rmsnorm_bwd(const void* x, long long ldx, const float* gamma, const void* dy, long long lddy,
                              const void* add, long long ldadd, void* dx, long long lddx, float* dgamma, int M, int D,
                              float eps, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && gamma && dy && dx && dgamma && workspace && M > 0, "lp_rmsnorm_bwd: null argument or M=%d", M);
  LP_REQUIRE(D % 8 == 0 && D >= 8 && D <= 8192, "lp_rmsnorm_bwd: D=%d must be a multiple of 8, at most 8192", D);
  LP_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && (add == nullptr || ldadd % 8 == 0) && al16(x) && al16(dy) &&
                 al16(dx) && (add == nullptr || al16(add)) && al16(workspace),
             "lp_rmsnorm_bwd: operands must be 16-byte aligned with leading dimensions that are multiples of 8");
  LP_REQUIRE(workspace_bytes >= lp_rmsnorm_bwd_workspace_bytes(D), "lp_rmsnorm_bwd: workspace too small");
  const int nb = M < NB_PARTIAL ? M : NB_PARTIAL;
  float* partial = reinterpret_cast<float*>(workspace);
  const bf16 *xb = reinterpret_cast<const bf16*>(x), *dyb = reinterpret_cast<const bf16*>(dy),
             *ab = reinterpret_cast<const bf16*>(add);
  bf16* dxb = reinterpret_cast<bf16*>(dx);
  note_launch(2);
  const int J = (D + 2047) / 2048;
  if (J == 1)
    rmsnorm_bwd_kernel<1><<<nb, 256, 0, stream>>>(xb, ldx, gamma, dyb, lddy, ab, ldadd, dxb, lddx, partial, M, D, eps);
  else if (J == 2)
    rmsnorm_bwd_kernel<2><<<nb, 256, 0, stream>>>(xb, ldx, gamma, dyb, lddy, ab, ldadd, dxb, lddx, partial, M, D, eps);
  else
    rmsnorm_bwd_kernel<4><<<nb, 256, 0, stream>>>(xb, ldx, gamma, dyb, lddy, ab, ldadd, dxb, lddx, partial, M, D, eps);
  LP_CHECK_CUDA(cudaGetLastError());
  partial_reduce_kernel<<<(D + 255) / 256, 256, 0, stream>>>(partial, nb, D, dgamma);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t lp_layernorm_bwd_workspace_bytes(int D) { return (size_t)NB_PARTIAL * 2 * (size_t)D * sizeof(float); }

extern "C" int lp_layernorm_bwd(const void* x, long long ldx, const float* gamma, const void* dy, long long lddy,
                                const void* add, long long ldadd, void* dx, long long lddx, float* dgamma, float* dbeta, int M,
                                int D, float eps, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && gamma && dy && dx && dgamma && dbeta && workspace && M > 0, "lp_layernorm_bwd: null argument or M=%d", M);
  LP_REQUIRE(D % 8 == 0 && D >= 8 && D <= 4096, "lp_layernorm_bwd: D=%d must be a multiple of 8, at most 4096", D);
  LP_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && (add == nullptr || ldadd % 8 == 0) && al16(x) && al16(dy) &&
                 al16(dx) && (add == nullptr || al16(add)) && al16(workspace),
             "lp_layernorm_bwd: operands must be 16-byte aligned with leading dimensions that are multiples of 8");
  LP_REQUIRE(workspace_bytes >= lp_layernorm_bwd_workspace_bytes(D), "lp_layernorm_bwd: workspace too small");
  const int nb = M < NB_PARTIAL ? M : NB_PARTIAL;
  float* partial = reinterpret_cast<float*>(workspace);
  const bf16 *xb = reinterpret_cast<const bf16*>(x), *dyb = reinterpret_cast<const bf16*>(dy),
             *ab = reinterpret_cast<const bf16*>(add);
  bf16* dxb = reinterpret_cast<bf16*>(dx);
  note_launch(3);
  if (D <= 2048)
    layernorm_bwd_kernel<1><<<nb, 256, 0, stream>>>(xb, ldx, gamma, dyb, lddy, ab, ldadd, dxb, lddx, partial, M, D, eps);
  else
    layernorm_bwd_kernel<2><<<nb, 256, 0, stream>>>(xb, ldx, gamma, dyb, lddy, ab, ldadd, dxb, lddx, partial, M, D, eps);
  LP_CHECK_CUDA(cudaGetLastError());
  // partial rows are [dg | db] of width 2 D: reduce both halves with one launch over 2 D columns, then split
  partial_reduce2_kernel<<<(2 * D + 255) / 256, 256, 0, stream>>>(partial, nb, D, dgamma, dbeta);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_pixel_unshuffle(const void* d, long long ldd, void* out, long long ldo, int M, int D, int side,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(d && out && M > 0 && D > 0 && D % 8 == 0 && side > 0 && side % 2 == 0 && M % (side * side) == 0,
             "lp_pixel_unshuffle: bad arguments M=%d D=%d side=%d", M, D, side);
  LP_REQUIRE(al16(d) && al16(out) && ldd % 8 == 0 && ldo % 8 == 0 && ldd >= 4LL * D && ldo >= D,
             "lp_pixel_unshuffle: operands must be 16-byte aligned with leading dimensions that are multiples of 8");
  note_launch();
  pixel_unshuffle_kernel<<<M, 128, 0, stream>>>(reinterpret_cast<const bf16*>(d), ldd, reinterpret_cast<bf16*>(out), ldo, side,
                                                D);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_rowmod_sum(const void* d, long long ldd, int n_rep, int T, int D, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(d && out && n_rep > 0 && T > 0 && D > 0 && D % 2 == 0 && ldd % 2 == 0 && al4(d),
             "lp_rowmod_sum: bad arguments n_rep=%d T=%d D=%d", n_rep, T, D);
  note_launch();
  rowmod_sum_kernel<<<T, 256, 0, stream>>>(reinterpret_cast<const bf16*>(d), ldd, n_rep, T, D, out);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int swiglu_checks(const char* who, const void* gu, long long ldgu, const void* other, long long ldo, int M, int F) {
  LP_REQUIRE(gu && other && M > 0 && F > 0 && F % 32 == 0, "%s: bad arguments M=%d F=%d (F must be a multiple of 32)", who, M, F);
  LP_REQUIRE(al16(gu) && al16(other) && ldgu % 8 == 0 && ldo % 8 == 0 && ldgu >= 2LL * F,
             "%s: operands must be 16-byte aligned with leading dimensions that are multiples of 8", who);
  return 0;
}

extern "C" int lp_swiglu_fwd(const void* gu, long long ldgu, void* out, long long ldo, int M, int F, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = swiglu_checks("lp_swiglu_fwd", gu, ldgu, out, ldo, M, F)) return rc;
  note_launch();
  swiglu_kernel<false><<<grid_for((long long)M * (F >> 3)), 256, 0, stream>>>(reinterpret_cast<const bf16*>(gu), ldgu, nullptr, 0,
                                                                             reinterpret_cast<bf16*>(out), ldo, M, F);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_swiglu_bwd(const void* gu, long long ldgu, const void* dm, long long lddm, void* dgu, long long lddgu,
                             int M, int F, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = swiglu_checks("lp_swiglu_bwd", gu, ldgu, dgu, lddgu, M, F)) return rc;
  LP_REQUIRE(dm && al16(dm) && lddm % 8 == 0 && lddgu >= 2LL * F, "lp_swiglu_bwd: bad dm / dgu layout");
  note_launch();
  swiglu_kernel<true><<<grid_for((long long)M * (F >> 3)), 256, 0, stream>>>(reinterpret_cast<const bf16*>(gu), ldgu,
                                                                            reinterpret_cast<const bf16*>(dm), lddm,
                                                                            reinterpret_cast<bf16*>(dgu), lddgu, M, F);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_gelu_bwd(const void* z, long long ldz, const void* dy, long long lddy, void* dz, long long lddz, int M,
                           int N, int act, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(z && dy && dz && M > 0 && N > 0 && N % 8 == 0, "lp_gelu_bwd: bad arguments M=%d N=%d", M, N);
  LP_REQUIRE(act == LP_ACT_GELU_TANH || act == LP_ACT_GELU_ERF, "lp_gelu_bwd: act must be LP_ACT_GELU_TANH or LP_ACT_GELU_ERF");
  LP_REQUIRE(al16(z) && al16(dy) && al16(dz) && ldz % 8 == 0 && lddy % 8 == 0 && lddz % 8 == 0,
             "lp_gelu_bwd: operands must be 16-byte aligned with leading dimensions that are multiples of 8");
  note_launch();
  gelu_bwd_kernel<<<grid_for((long long)M * (N >> 3)), 256, 0, stream>>>(reinterpret_cast<const bf16*>(z), ldz,
                                                                        reinterpret_cast<const bf16*>(dy), lddy,
                                                                        reinterpret_cast<bf16*>(dz), lddz, M, N, act);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_colsum_bf16(const void* dy, long long ld, int M, int N, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(dy && out && M > 0 && N > 0 && N % 2 == 0 && ld % 2 == 0 && al4(dy), "lp_colsum_bf16: bad arguments M=%d N=%d", M, N);
  note_launch();
  colsum_kernel<<<(N + 63) / 64, 256, 0, stream>>>(reinterpret_cast<const bf16*>(dy), ld, M, N, out);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_ce_fwd_bwd(const void* logits, int logits_fp32, long long ld, const long long* target, const float* weight,
                             float* loss, void* dlogits, long long ldd, int rows, int V, int Vp, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(logits && target && weight && loss && dlogits && rows > 0 && V > 0, "lp_ce_fwd_bwd: bad arguments rows=%d V=%d", rows, V);
  LP_REQUIRE(Vp >= V && Vp % 2 == 0 && ldd % 2 == 0 && ldd >= Vp && ld >= V && al4(dlogits),
             "lp_ce_fwd_bwd: Vp=%d must be even and >= V=%d, ldd=%lld even and >= Vp", Vp, V, ldd);
  note_launch();
  if (logits_fp32)
    ce_fwd_bwd_kernel<float><<<rows, 256, 0, stream>>>(reinterpret_cast<const float*>(logits), ld, target, weight, loss,
                                                       reinterpret_cast<bf16*>(dlogits), ldd, V, Vp);
  else
    ce_fwd_bwd_kernel<bf16><<<rows, 256, 0, stream>>>(reinterpret_cast<const bf16*>(logits), ld, target, weight, loss,
                                                      reinterpret_cast<bf16*>(dlogits), ldd, V, Vp);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_embed_bwd(const void* d, long long ldd, const long long* ids, int S, long long image_id, float* dE,
                            long long ldE, int D, long long vocab, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(d && ids && dE && S > 0 && D > 0 && D % 2 == 0 && ldd % 2 == 0 && al4(d), "lp_embed_bwd: bad arguments S=%d D=%d", S, D);
  LP_REQUIRE(vocab > 0, "lp_embed_bwd: vocab=%lld", vocab);
  note_launch();
  embed_bwd_kernel<<<S, 256, 0, stream>>>(reinterpret_cast<const bf16*>(d), ldd, ids, image_id, dE, ldE, D, vocab);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_gather_rows_bf16(const void* in, long long ld_in, const int* idx, int n, int D, void* out, long long ld_out,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(in && idx && out && n > 0 && D > 0 && D % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && al16(in) && al16(out),
             "lp_gather_rows_bf16: bad arguments n=%d D=%d", n, D);
  note_launch();
  gather_rows_kernel<<<n, 128, 0, stream>>>(reinterpret_cast<const bf16*>(in), ld_in, idx, D, reinterpret_cast<bf16*>(out), ld_out);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_softmax_bwd(const float* scores, long long lds, const float* dp, long long lddp, void* P, long long ldp,
                              void* dS, long long ldds, int Sq, int Sk, int causal, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(scores && dp && P && dS && Sq > 0 && Sk > 0, "lp_softmax_bwd: bad arguments Sq=%d Sk=%d", Sq, Sk);
  note_launch();
  softmax_bwd_kernel<<<Sq, 256, 0, stream>>>(scores, lds, dp, lddp, reinterpret_cast<bf16*>(P), ldp, reinterpret_cast<bf16*>(dS),
                                             ldds, Sk, causal, scale);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_accumulate_f32(float* dst, const float* src, long long n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(dst && src && n > 0 && al16(dst) && al16(src), "lp_accumulate_f32: bad arguments n=%lld", n);
  note_launch();
  accumulate_kernel<<<grid_for((n + 3) / 4), 256, 0, stream>>>(dst, src, n);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t lp_grad_norm_workspace_bytes(void) { return (size_t)NORM_BLOCKS * sizeof(float); }

extern "C" int lp_sumsq_accum(const float* x, long long n, float* partial, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && partial && n > 0 && al16(x), "lp_sumsq_accum: bad arguments n=%lld", n);
  note_launch();
  sumsq_kernel<<<NORM_BLOCKS, 256, 0, stream>>>(x, n, partial);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_grad_norm_finish(const float* partial, float max_norm, float* out2, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(partial && out2, "lp_grad_norm_finish: null argument");
  note_launch();
  grad_norm_finish_kernel<<<1, 256, 0, stream>>>(partial, NORM_BLOCKS, max_norm, out2);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, const float* norm_and_clip, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(p && g && m && v && n > 0 && step >= 1, "lp_adamw_step: bad arguments n=%lld step=%d", n, step);
  const float inv_bc1 = (float)(1.0 / (1.0 - pow((double)beta1, (double)step)));
  const float inv_bc2 = (float)(1.0 / (1.0 - pow((double)beta2, (double)step)));
  note_launch();
  if (n % 4 == 0 && al16(p) && al16(g) && al16(m) && al16(v) && (p_bf16 == nullptr || (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0)) {
    adamw_vec4_kernel<<<grid_for(n / 4), 256, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                                            reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
                                                            reinterpret_cast<uint2*>(p_bf16), n / 4, lr, beta1, beta2, eps,
                                                            weight_decay, inv_bc1, inv_bc2, norm_and_clip);
    LP_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  adamw_kernel<<<grid_for(n), 256, 0, stream>>>(p, g, m, v, reinterpret_cast<bf16*>(p_bf16), n, lr, beta1, beta2, eps,
                                                weight_decay, inv_bc1, inv_bc2, norm_and_clip);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
