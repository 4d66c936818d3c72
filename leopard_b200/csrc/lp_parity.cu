// leopard_b200 — PARITY MODE glue: fp32 element-wise kernels for the fp32-equivalent forward (leopard_b200/parity.py).
//
// The reference's HF evaluation runs the whole model in fp32 (evaluations/models/llava_multiimg_siglip_anyres.py:373); BASELINE.json
// asks for logits within 1e-3 max-abs of it — below bf16 resolution through 59 layers.  Parity mode keeps every
// activation in fp32 and runs every contraction on the SAME tcgen05 GEMM kernel with the fp32 operand split into three bf16
// terms (x = hi + mid + lo, 8 + 8 + 8 mantissa bits): products of bf16 terms are exact in the fp32 TMEM accumulator, so
// A W^T = (hi + mid + lo) W^T is computed to fp32 accuracy as ONE bf16 GEMM over a 3x longer contraction (weights are
// bf16-exact; attention, where both operands are fp32, uses the six significant cross terms).  What remains between the GEMMs —
// norms, activations, softmax, RoPE, residual adds — are the small fp32 kernels below, written for accuracy (tanhf / erff / expf,
// true division), not speed.  It is a verification path: the fast path (bf16 activations) is what bench.py measures.
#include "lp_host.h"
#include "../../include/leopard_b200.h"
#include <cuda_bf16.h>
#include <math.h>

using namespace lp;

namespace {

typedef __nv_bfloat16 bf16;

// x -> (hi, mid, lo) bf16 with hi + mid + lo == x to ~2^-24 relative
__device__ __forceinline__ void split3(float x, bf16& h, bf16& m, bf16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

// layout 0: [h | m | l] (3K)   pairs with a weight tiled 3x
// layout 1: [h | h | h | m | m | l] (6K)   "A side" of a two-sided product
// layout 2: [h | m | l | h | m | h] (6K)   "B side": column blocks pair up as hh, hm, hl, mh, mm, lh
__global__ void split_kernel(const float* __restrict__ x, long long ldx, int M, int K, bf16* __restrict__ out, long long ldo,
                             int layout) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * K) return;
  const int r = (int)(i / K), c = (int)(i % K);
  bf16 h, m, l;
  split3(x[(long long)r * ldx + c], h, m, l);
  bf16* o = out + (long long)r * ldo + c;
  if (layout == 0) { o[0] = h; o[K] = m; o[2LL * K] = l; }
  else if (layout == 1) { o[0] = h; o[K] = h; o[2LL * K] = h; o[3LL * K] = m; o[4LL * K] = m; o[5LL * K] = l; }
  else { o[0] = h; o[K] = m; o[2LL * K] = l; o[3LL * K] = h; o[4LL * K] = m; o[5LL * K] = h; }
}

// LayerNorm (beta != null) / RMSNorm (rms != 0) over fp32 rows, two-pass statistics in fp32 with a block reduction
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < THREADS / 32; ++i) t += red[i];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256)
norm_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ out, long long ldo, int D, float eps, int rms) {
  __shared__ float red[8];
  const float* xr = x + (long long)blockIdx.x * ldx;
  float* orow = out + (long long)blockIdx.x * ldo;
  float mu = 0.f;
  if (!rms) {
    float s = 0.f;
    for (int c = threadIdx.x; c < D; c += 256) s += xr[c];
    mu = block_sum<256>(s, red) / (float)D;
  }
  float s2 = 0.f;
  for (int c = threadIdx.x; c < D; c += 256) {
    const float d = xr[c] - mu;
    s2 += d * d;
  }
  const float rstd = 1.0f / sqrtf(block_sum<256>(s2, red) / (float)D + eps);
  for (int c = threadIdx.x; c < D; c += 256) {
    const float v = (xr[c] - mu) * rstd * gamma[c];
    orow[c] = beta ? v + beta[c] : v;
  }
}

// act 1: GELU tanh, 2: GELU erf (in place);  act 3: SwiGLU on 32-column gate / up granules -> out [M, N/2]
__global__ void act_f32_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out, long long ldo, int M, int N,
                               int act) {
  const int n_out = act == LP_ACT_SWIGLU ? N / 2 : N;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * n_out) return;
  const int r = (int)(i / n_out), c = (int)(i % n_out);
  const float* xr = x + (long long)r * ldx;
  float y;
  if (act == LP_ACT_SWIGLU) {
    const int gcol = (c >> 5) * 64 + (c & 31);
    const float g = xr[gcol], u = xr[gcol + 32];
    y = g / (1.0f + expf(-g)) * u;
  } else {
    const float v = xr[c];
    if (act == LP_ACT_GELU_TANH) y = 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));
    else y = 0.5f * v * (1.0f + erff(v * 0.7071067811865476f));
  }
  out[(long long)r * ldo + c] = y;
}

__global__ void add_f32_kernel(float* __restrict__ y, long long ldy, const float* __restrict__ x, long long ldx, int M, int N,
                               int x_row_mod) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int r = (int)(i / N), c = (int)(i % N);
  const int xr = x_row_mod > 0 ? r % x_row_mod : r;
  y[(long long)r * ldy + c] += x[(long long)xr * ldx + c];
}

// rotate-half RoPE in place on heads [0, n_heads) of x [S, ld] (head_dim columns each), cos / sin fp32 [S, head_dim / 2]
__global__ void rope_f32_kernel(float* __restrict__ x, long long ld, int S, int n_heads, int head_dim,
                                const float* __restrict__ cs, const float* __restrict__ sn) {
  const int half = head_dim / 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)S * n_heads * half) return;
  const int j = (int)(i % half), h = (int)((i / half) % n_heads), r = (int)(i / ((long long)half * n_heads));
  float* p = x + (long long)r * ld + h * head_dim + j;
  const float c = cs[(long long)r * half + j], s = sn[(long long)r * half + j];
  const float a = p[0], b = p[half];
  p[0] = a * c - b * s;
  p[half] = b * c + a * s;
}

// softmax(scale * scores) per row, optional causal mask (key j <= query row + offset), in place, fp32
__global__ void __launch_bounds__(256)
softmax_f32_kernel(float* __restrict__ sc, long long ld, int Tk, float scale, int causal) {
  __shared__ float red[8];
  float* row = sc + (long long)blockIdx.x * ld;
  const int limit = causal ? min(Tk, (int)blockIdx.x + 1) : Tk;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < limit; c += 256) mx = fmaxf(mx, row[c] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < limit; c += 256) {
    const float e = expf(row[c] * scale - mx);
    row[c] = e;
    s += e;
  }
  const float inv = 1.0f / block_sum<256>(s, red);
  for (int c = threadIdx.x; c < Tk; c += 256) row[c] = c < limit ? row[c] * inv : 0.f;
}

inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" int lp_parity_split(const float* x, long long ldx, int M, int K, void* out, long long ldo, int layout, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && out && M > 0 && K > 0 && layout >= 0 && layout <= 2, "lp_parity_split: bad arguments M=%d K=%d layout=%d", M, K, layout);
  LP_REQUIRE(ldo >= (long long)(layout == 0 ? 3 : 6) * K, "lp_parity_split: ldo=%lld too small", ldo);
  note_launch();
  split_kernel<<<blocks_for((long long)M * K), 256, 0, stream>>>(x, ldx, M, K, reinterpret_cast<bf16*>(out), ldo, layout);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_parity_norm(const float* x, long long ldx, const float* gamma, const float* beta, float* out, long long ldo, int M,
                              int D, float eps, int rms, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && gamma && out && M > 0 && D > 0, "lp_parity_norm: bad arguments M=%d D=%d", M, D);
  note_launch();
  norm_f32_kernel<<<M, 256, 0, stream>>>(x, ldx, gamma, rms ? nullptr : beta, out, ldo, D, eps, rms);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_parity_act(const float* x, long long ldx, float* out, long long ldo, int M, int N, int act, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && out && M > 0 && N > 0 && act >= LP_ACT_GELU_TANH && act <= LP_ACT_SWIGLU && (act != LP_ACT_SWIGLU || N % 64 == 0),
             "lp_parity_act: bad arguments M=%d N=%d act=%d", M, N, act);
  note_launch();
  act_f32_kernel<<<blocks_for((long long)M * (act == LP_ACT_SWIGLU ? N / 2 : N)), 256, 0, stream>>>(x, ldx, out, ldo, M, N, act);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_parity_add(float* y, long long ldy, const float* x, long long ldx, int M, int N, int x_row_mod, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && y && M > 0 && N > 0, "lp_parity_add: bad arguments M=%d N=%d", M, N);
  note_launch();
  add_f32_kernel<<<blocks_for((long long)M * N), 256, 0, stream>>>(y, ldy, x, ldx, M, N, x_row_mod);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_parity_rope(float* x, long long ld, int S, int n_heads, int head_dim, const float* cs, const float* sn,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && cs && sn && S > 0 && n_heads > 0 && head_dim > 0 && head_dim % 2 == 0, "lp_parity_rope: bad arguments");
  note_launch();
  rope_f32_kernel<<<blocks_for((long long)S * n_heads * (head_dim / 2)), 256, 0, stream>>>(x, ld, S, n_heads, head_dim, cs, sn);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_parity_softmax(float* scores, long long ld, int Tq, int Tk, float scale, int causal, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(scores && Tq > 0 && Tk > 0, "lp_parity_softmax: bad arguments Tq=%d Tk=%d", Tq, Tk);
  note_launch();
  softmax_f32_kernel<<<Tq, 256, 0, stream>>>(scores, ld, Tk, scale, causal);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
