// leopard_b200 — single-token decode step over the prefill's KV cache (greedy generation after the prefill).
//
// The cached-decode branch of myLlavaForConditionalGeneration.forward
// (evaluations/models/llava_multiimg_siglip_anyres.py:291-320, driven by `generate(..., max_new_tokens=128,
// use_cache=True)` at :448-452): with B = 1 and no padding the rebuilt mask is all ones and the position is the
// number of cached rows.
//
// The step is HBM-bound: every decoder weight (16.06 GB for Llama-3.1-8B) is read exactly once per token, so the
// floor is weights / HBM bandwidth (~2.5 ms on a B200) and the design is about keeping the memory system saturated
// across ~200 short kernels:
//   gemv_kernel<MODE>   y = W x for ONE activation row: a warp owns two weight rows at a time and streams them with
//                       8 x 16-byte non-allocating loads in flight per lane; x is staged once per CTA in shared memory.
//                       Fused prologue: RMSNorm of the residual stream (MLM/legacy/model/rms_norm.py:26-31).
//                       Fused epilogues: QKV -> rotate-half RoPE on (i, i+64) row pairs + append K/V to the cache
//                       (MP/model/llava/transformer.py:853-856); gate/up -> silu(gate)*up (:129-139); o/down -> residual
//                       add in place; lm_head -> fp32 logits.
//   decode_attn_*       split-K GQA attention over the bf16 KV cache (heads x 16 key ranges) + combine.
//   argmax / embed      greedy token selection and the next embedding row, device-resident, so that a whole step
//                       (lp_decode_greedy_step) reads its position and token from device memory and can be captured
//                       once into a CUDA graph and replayed per token with no host work.
// All kernels start with griddepcontrol.wait and may be launched with programmatic dependent launch (mode bit 1): the
// next kernel's CTAs become resident and prefetch their first weight rows into L2 while the previous kernel drains.
// Mode bit 0 = 0 selects the first implementation (M = 1 rows through the tcgen05 GEMM), kept for A/B measurements.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

int rope_table_base(int pos_base, const float* inv_freq, float* cs, float* sn, int S, int half, cudaStream_t stream);

static int g_decode_mode = 1;  // bit 0: GEMV kernels, bit 1: programmatic dependent launch

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ int decode_position(const int* pos_dev, int pos_host) {
  return pos_dev ? *pos_dev : pos_host;
}

// ---------------------------------------------------------------------------------------------------- GEMV
constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_UNROLL = 4;  // 16-byte loads in flight per lane and weight row
enum GemvMode { GV_PLAIN = 0, GV_QKV = 1, GV_SWIGLU = 2, GV_LOGITS = 3 };

struct GemvArgs {
  const bf16* W;       // [N, K] row-major (nn.Linear layout)
  long long ldw;
  int N, K;
  const bf16* x;       // activation row [K]
  const float* gamma;  // non-null: x <- bf16(x * rsqrt(mean(x^2) + eps) * gamma) first (K == hidden)
  float eps;
  bf16* y;             // GV_PLAIN: y[N] (+= in place when residual), GV_SWIGLU: y[N/2]
  int residual;
  float* y32;          // GV_LOGITS
  // GV_QKV: rows = q heads | k heads | v heads, head_dim 128
  bf16* q_out;         // [heads * 128]
  bf16* kc;            // this layer's K cache [kv_capacity, kv_heads * 128]
  bf16* vc;
  int heads, kv_heads;
  long long kv_capacity;
  const float* inv_freq;  // [64]
  const int* pos_dev;
  int pos_host;
};

__device__ __forceinline__ float dot8(const uint4& w, const float* xf, float acc) {
  const float2 a = unpack_bf16x2(w.x), b = unpack_bf16x2(w.y), c = unpack_bf16x2(w.z), d = unpack_bf16x2(w.w);
  acc = fmaf(a.x, xf[0], acc); acc = fmaf(a.y, xf[1], acc);
  acc = fmaf(b.x, xf[2], acc); acc = fmaf(b.y, xf[3], acc);
  acc = fmaf(c.x, xf[4], acc); acc = fmaf(c.y, xf[5], acc);
  acc = fmaf(d.x, xf[6], acc); acc = fmaf(d.y, xf[7], acc);
  return acc;
}

// group g of a launch -> the two weight rows one warp reduces together
template <int MODE>
__device__ __forceinline__ void gemv_group_rows(const GemvArgs& a, int g, int& r0, int& r1) {
  if (MODE == GV_QKV) {
    const int n_rot = (a.heads + a.kv_heads) * 64;  // rotate-half partners (i, i + 64) of every q and k head
    if (g < n_rot) { r0 = (g >> 6) * 128 + (g & 63); r1 = r0 + 64; }
    else { r0 = (a.heads + a.kv_heads) * 128 + 2 * (g - n_rot); r1 = r0 + 1; }
  } else if (MODE == GV_SWIGLU) {  // 32-row granules [gate | up]: gate row i pairs with up row i + 32
    r0 = (g >> 5) * 64 + (g & 31); r1 = r0 + 32;
  } else {
    r0 = 2 * g; r1 = r0 + 1;
  }
}

template <int MODE>
__global__ void __launch_bounds__(GV_THREADS, 4)
gemv_kernel(const GemvArgs a) {
  extern __shared__ uint4 xs[];  // K / 8 chunks of 8 bf16
  __shared__ float red[GV_WARPS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K8 = a.K >> 3;
  const int n_groups = a.N >> 1;
  const int warps_total = gridDim.x * GV_WARPS;
  const int wg = blockIdx.x * GV_WARPS + warp;

  // weights do not depend on the previous kernel: warm L2 with this warp's first rows, then let the next kernel in
  if (wg < n_groups) {
    int r0, r1;
    gemv_group_rows<MODE>(a, wg, r0, r1);
    const char* p0 = reinterpret_cast<const char*>(a.W + (long long)r0 * a.ldw);
    const char* p1 = reinterpret_cast<const char*>(a.W + (long long)r1 * a.ldw);
    for (int off = lane * 128; off < a.K * 2; off += 4096) { prefetch_l2(p0 + off); prefetch_l2(p1 + off); }
  }
  pdl_launch_dependents();
  pdl_wait();

  // ---- stage the activation row (optionally RMS-normalised) in shared memory
  const uint4* xg = reinterpret_cast<const uint4*>(a.x);
  if (a.gamma != nullptr) {
    float ss = 0.f;
    for (int c = tid; c < K8; c += GV_THREADS) {
      const uint4 u = xg[c];
      xs[c] = u;
      const float2 p = unpack_bf16x2(u.x), q = unpack_bf16x2(u.y), r = unpack_bf16x2(u.z), s = unpack_bf16x2(u.w);
      ss += p.x * p.x + p.y * p.y + q.x * q.x + q.y * q.y + r.x * r.x + r.y * r.y + s.x * s.x + s.y * s.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < GV_WARPS; ++w) tot += red[w];
    const float rstd = rsqrtf(tot / (float)a.K + a.eps);
    for (int c = tid; c < K8; c += GV_THREADS) {  // each thread re-reads the chunks it wrote itself
      const uint4 u = xs[c];
      const float4 g0 = *reinterpret_cast<const float4*>(a.gamma + c * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(a.gamma + c * 8 + 4);
      const float2 p = unpack_bf16x2(u.x), q = unpack_bf16x2(u.y), r = unpack_bf16x2(u.z), s = unpack_bf16x2(u.w);
      uint4 y;
      y.x = pack_bf16x2(p.x * rstd * g0.x, p.y * rstd * g0.y);
      y.y = pack_bf16x2(q.x * rstd * g0.z, q.y * rstd * g0.w);
      y.z = pack_bf16x2(r.x * rstd * g1.x, r.y * rstd * g1.y);
      y.w = pack_bf16x2(s.x * rstd * g1.z, s.y * rstd * g1.w);
      xs[c] = y;
    }
  } else {
    for (int c = tid; c < K8; c += GV_THREADS) xs[c] = xg[c];
  }
  __syncthreads();

  const int pos = (MODE == GV_QKV) ? decode_position(a.pos_dev, a.pos_host) : 0;

  for (int g = wg; g < n_groups; g += warps_total) {
    int r0, r1;
    gemv_group_rows<MODE>(a, g, r0, r1);
    const uint4* w0 = reinterpret_cast<const uint4*>(a.W + (long long)r0 * a.ldw);
    const uint4* w1 = reinterpret_cast<const uint4*>(a.W + (long long)r1 * a.ldw);
    float acc0 = 0.f, acc1 = 0.f;
    for (int c = lane; c < K8; c += 32 * GV_UNROLL) {
      uint4 u0[GV_UNROLL], u1[GV_UNROLL];
#pragma unroll
      for (int u = 0; u < GV_UNROLL; ++u) {
        const int cc = c + u * 32;
        if (cc < K8) { u0[u] = ld_stream16(w0 + cc); u1[u] = ld_stream16(w1 + cc); }
        else { u0[u] = make_uint4(0u, 0u, 0u, 0u); u1[u] = u0[u]; }
      }
#pragma unroll
      for (int u = 0; u < GV_UNROLL; ++u) {
        const int cc = c + u * 32;
        if (cc < K8) {
          const uint4 xv = xs[cc];
          const float2 p = unpack_bf16x2(xv.x), q = unpack_bf16x2(xv.y), r = unpack_bf16x2(xv.z),
                       s = unpack_bf16x2(xv.w);
          const float xf[8] = {p.x, p.y, q.x, q.y, r.x, r.y, s.x, s.y};
          acc0 = dot8(u0[u], xf, acc0);
          acc1 = dot8(u1[u], xf, acc1);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
    }
    if (lane != 0) continue;
    if (MODE == GV_PLAIN) {
      if (a.residual) { acc0 += __bfloat162float(a.y[r0]); acc1 += __bfloat162float(a.y[r1]); }
      a.y[r0] = __float2bfloat16(acc0);
      a.y[r1] = __float2bfloat16(acc1);
    } else if (MODE == GV_LOGITS) {
      a.y32[r0] = acc0;
      a.y32[r1] = acc1;
    } else if (MODE == GV_SWIGLU) {
      a.y[g] = __float2bfloat16(silu_f(acc0) * acc1);  // output column 32*(g>>5) + (g&31) == g
    } else {  // GV_QKV
      const int q_w = a.heads * 128, kv_w = a.kv_heads * 128;
      if (pos < 0 || pos >= a.kv_capacity) continue;  // host validates; never write outside the cache
      if (r0 < q_w + kv_w) {
        // same rounding points as the prefill: GEMM output -> bf16, then t*cos + rotate_half(t)*sin -> bf16
        const float t1 = __bfloat162float(__float2bfloat16(acc0)), t2 = __bfloat162float(__float2bfloat16(acc1));
        const float ang = (float)pos * a.inv_freq[r0 & 63];
        const float cs = cosf(ang), sn = sinf(ang);
        const bf16 o1 = __float2bfloat16(t1 * cs - t2 * sn), o2 = __float2bfloat16(t2 * cs + t1 * sn);
        if (r0 < q_w) { a.q_out[r0] = o1; a.q_out[r1] = o2; }
        else {
          bf16* kp = a.kc + (long long)pos * kv_w + (r0 - q_w);
          kp[0] = o1; kp[64] = o2;
        }
      } else {
        bf16* vp = a.vc + (long long)pos * kv_w + (r0 - q_w - kv_w);
        vp[0] = __float2bfloat16(acc0);
        vp[1] = __float2bfloat16(acc1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- attention
constexpr int DEC_WARPS = 8;
constexpr int DEC_SPLITS = 16;  // key ranges per head -> heads * DEC_SPLITS CTAs (>= 1 wave on 148 SMs)

// Split-K decode attention, pass 1: CTA (head h, split s) scans keys [s*chunk, (s+1)*chunk) of kv head h/group.
// A warp handles FOUR keys per step (8 lanes x 16 dims each, two 16-byte loads per lane, 3-step shuffle reduce) so that
// four independent load->dot->exp chains are in flight; partial (m, l, acc[128]) go to `part` for the combine pass.
// Attends to rows 0 .. position (the row at `position` was appended by this step's QKV kernel).
__global__ void __launch_bounds__(DEC_WARPS * 32)
decode_attn_partial_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kc, const bf16* __restrict__ vc,
                           const int* __restrict__ pos_dev, int pos_host, int max_ctx, int kv_w, int group,
                           float scale_log2, float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  const int n_ctx = min(decode_position(pos_dev, pos_host) + 1, max_ctx);  // never read past the cache
  const int h = blockIdx.x, split = blockIdx.y, g = h / group;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, l8 = lane & 7;  // key slot within the warp step, 16-dim slice of the head
  const unsigned gmask = 0xFFu << (sub * 8);
  const int chunk = (n_ctx + DEC_SPLITS - 1) / DEC_SPLITS;
  const int t_begin = split * chunk, t_end = min(n_ctx, t_begin + chunk);
  __shared__ float s_m[DEC_WARPS * 4], s_l[DEC_WARPS * 4], s_acc[DEC_WARPS * 4][128];
  float qv[16];
  {
    const uint4 u0 = *reinterpret_cast<const uint4*>(q + h * 128 + l8 * 16);
    const uint4 u1 = *reinterpret_cast<const uint4*>(q + h * 128 + l8 * 16 + 8);
    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = unpack_bf16x2(w[i]);
      qv[2 * i] = f.x * scale_log2;
      qv[2 * i + 1] = f.y * scale_log2;
    }
  }
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int t = t_begin + warp * 4 + sub; t < t_end; t += DEC_WARPS * 4) {
    const bf16* kp = kc + (long long)t * kv_w + g * 128 + l8 * 16;
    const bf16* vp = vc + (long long)t * kv_w + g * 128 + l8 * 16;
    const uint4 k0 = *reinterpret_cast<const uint4*>(kp), k1 = *reinterpret_cast<const uint4*>(kp + 8);
    const uint4 v0 = *reinterpret_cast<const uint4*>(vp), v1 = *reinterpret_cast<const uint4*>(vp + 8);
    const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = unpack_bf16x2(kw[i]);
      s += qv[2 * i] * f.x + qv[2 * i + 1] * f.y;
    }
    // only the 8 lanes that share this key take part: key slots of one warp run different trip counts at the tail
    s += __shfl_xor_sync(gmask, s, 4);
    s += __shfl_xor_sync(gmask, s, 2);
    s += __shfl_xor_sync(gmask, s, 1);
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    l = l * corr + p;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = unpack_bf16x2(vw[i]);
      acc[2 * i] = acc[2 * i] * corr + p * f.x;
      acc[2 * i + 1] = acc[2 * i + 1] * corr + p * f.y;
    }
    m = mn;
  }
  // note: a lane whose key slots were all beyond t_end keeps m = -inf, l = 0 (handled by the combine weights)
  const int slot = warp * 4 + sub;
  if (l8 == 0) { s_m[slot] = m; s_l[slot] = l; }
#pragma unroll
  for (int i = 0; i < 16; ++i) s_acc[slot][l8 * 16 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int d = threadIdx.x;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < DEC_WARPS * 4; ++w) mm = fmaxf(mm, s_m[w]);
    float ll = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS * 4; ++w) {
      const float c = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - mm);
      ll += s_l[w] * c;
      o += s_acc[w][d] * c;
    }
    float* pp = part + ((long long)h * DEC_SPLITS + split) * 130;
    pp[2 + d] = o;
    if (d == 0) { pp[0] = mm; pp[1] = ll; }
  }
}

// pass 2: merge the DEC_SPLITS partials of each head
__global__ void __launch_bounds__(128)
decode_attn_combine_kernel(const float* __restrict__ part, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, d = threadIdx.x;
  const float* pp = part + (long long)h * DEC_SPLITS * 130;
  float mm = -INFINITY;
#pragma unroll
  for (int s = 0; s < DEC_SPLITS; ++s) mm = fmaxf(mm, pp[s * 130]);
  float ll = 0.f, o = 0.f;
#pragma unroll
  for (int s = 0; s < DEC_SPLITS; ++s) {
    const float m = pp[s * 130];
    const float c = (m == -INFINITY) ? 0.f : exp2f(m - mm);
    ll += pp[s * 130 + 1] * c;
    o += pp[s * 130 + 2 + d] * c;
  }
  out[h * 128 + d] = __float2bfloat16(o / ll);
}

// ---------------------------------------------------------------------------------------------------- token plumbing
// x = table[clamp(token)]  (Embedding.forward, MP/model/llava/vlm_model.py:224-254; HF get_input_embeddings, EV :263)
__global__ void __launch_bounds__(256)
decode_embed_kernel(const long long* __restrict__ token, const bf16* __restrict__ table, long long ldt, long long vocab,
                    bf16* __restrict__ x, int H) {
  pdl_launch_dependents();
  pdl_wait();
  long long id = *token;
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const uint4* sp = reinterpret_cast<const uint4*>(table + id * ldt);
  uint4* dp = reinterpret_cast<uint4*>(x);
  for (int c = threadIdx.x; c < (H >> 3); c += blockDim.x) dp[c] = sp[c];
}

// greedy selection: token = argmax(logits[0:V]) (lowest index among ties, like torch.argmax), position += 1
__global__ void __launch_bounds__(1024)
decode_argmax_kernel(const float* __restrict__ logits, int V, long long* __restrict__ token_io,
                     int* __restrict__ position_io) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 1024) {
    const float v = logits[i];
    if (v > best) { best = v; bi = i; }  // ascending i per thread: the first maximum is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = s_v[threadIdx.x]; bi = s_i[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
      *token_io = (bi == 0x7fffffff) ? 0 : (long long)bi;  // all-NaN / -inf rows fall back to id 0
      *position_io += 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------- launch helpers
template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_decode_mode & 2) ? 1 : 0;
  LP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
  note_launch();
  return 0;
}

template <int MODE>
static int launch_gemv(const GemvArgs& a, cudaStream_t stream) {
  LP_REQUIRE(a.N % 2 == 0 && a.K % 8 == 0 && a.ldw % 8 == 0 && a.K * 2 <= 48 * 1024,
             "decode gemv: unsupported shape N=%d K=%d ldw=%lld", a.N, a.K, a.ldw);
  const int groups = a.N / 2;
  int grid = (groups + GV_WARPS - 1) / GV_WARPS;
  const int cap = sm_count() * 4;
  if (grid > cap) grid = cap;
  ProfScope prof(10 + MODE, a.N, a.K, 0, stream);
  return launch_pdl(gemv_kernel<MODE>, dim3(grid), dim3(GV_THREADS), (size_t)a.K * 2, stream, a);
}

}  // namespace lp

using namespace lp;

static inline size_t bf(size_t n) { return n * 2; }

extern "C" int lp_decode_set_mode(int mode) {
  const int old = g_decode_mode;
  g_decode_mode = mode & 3;
  return old;
}

extern "C" size_t lp_decode_workspace_bytes(const lp_dec_dims* d) {
  const size_t qkv_w = (size_t)(d->heads + 2 * d->kv_heads) * d->head_dim;
  return 256 * 9 + bf(2 * (size_t)d->hidden) * 2 + bf(qkv_w) + bf((size_t)d->ffn) + (size_t)d->head_dim * 4 + 64 * 4 +
         (size_t)d->heads * DEC_SPLITS * 130 * 4;
}

#define LP_TRY(call) do { rc = (call); if (rc) return rc; } while (0)

// first implementation: every projection is an M = 1 problem of the tcgen05 GEMM (kept for A/B measurements)
static int decode_step_v1(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table, long long ld_embed,
                          const long long* token_id, int position, void* kv_cache, long long kv_capacity, void* logits,
                          void* workspace, cudaStream_t stream) {
  const int H = d->hidden, hd = d->head_dim;
  const int q_w = d->heads * hd, kv_w = d->kv_heads * hd, qkv_w = q_w + 2 * kv_w;
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t n) { void* r = p; p += (n + 255) & ~size_t(255); return r; };
  __nv_bfloat16* x = (__nv_bfloat16*)take(bf(H));
  __nv_bfloat16* h = (__nv_bfloat16*)take(bf(H));
  __nv_bfloat16* qkv = (__nv_bfloat16*)take(bf(qkv_w));
  __nv_bfloat16* att = (__nv_bfloat16*)take(bf(H));
  __nv_bfloat16* act = (__nv_bfloat16*)take(bf(d->ffn));
  float* cs = (float*)take(hd / 2 * 4);
  float* sn = (float*)take(hd / 2 * 4);
  int* plan = (int*)take(16);
  float* part = (float*)take((size_t)d->heads * DEC_SPLITS * 130 * 4);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  int rc;
  // new token's embedding row (ids stay on the device: no host round trip of the sampled token is needed here)
  LP_TRY(lp_embed_merge(token_id, 1, -(1LL << 62), 1, embed_table, ld_embed, d->vocab, nullptr, 0, 0, x, H, 1, H, plan,
                        stream));
  LP_TRY(rope_table_base(position, w->inv_freq, cs, sn, 1, hd / 2, stream));
  for (int l = 0; l < d->layers; ++l) {
    const lp_dec_layer_weights* L = &w->layers[l];
    __nv_bfloat16* kc = (__nv_bfloat16*)kv_cache + (size_t)l * 2 * kv_capacity * kv_w;
    __nv_bfloat16* vc = kc + (size_t)kv_capacity * kv_w;
    LP_TRY(lp_rmsnorm(x, H, L->ln1_g, h, H, 1, H, d->eps, stream));
    LP_TRY(lp_gemm_bf16(h, H, L->qkv_w, H, qkv, qkv_w, 1, qkv_w, H, nullptr, nullptr, 0, 0, LP_ACT_NONE, 0, 128,
                        stream));
    LP_TRY(lp_rope_inplace(qkv, qkv_w, 1, 0, d->heads + d->kv_heads, hd, cs, sn, stream));
    LP_CHECK_CUDA(cudaMemcpyAsync(kc + (size_t)position * kv_w, qkv + q_w, bf(kv_w), cudaMemcpyDeviceToDevice, stream));
    LP_CHECK_CUDA(cudaMemcpyAsync(vc + (size_t)position * kv_w, qkv + q_w + kv_w, bf(kv_w), cudaMemcpyDeviceToDevice,
                                  stream));
    {
      ProfScope prof(8, position + 1, d->heads, 0, stream);
      note_launch(2);
      decode_attn_partial_kernel<<<dim3(d->heads, DEC_SPLITS), DEC_WARPS * 32, 0, stream>>>(
          qkv, kc, vc, nullptr, position, (int)kv_capacity, kv_w, d->heads / d->kv_heads, scale_log2, part);
      decode_attn_combine_kernel<<<d->heads, 128, 0, stream>>>(part, att);
      LP_CHECK_CUDA(cudaGetLastError());
    }
    LP_TRY(lp_gemm_bf16(att, H, L->o_w, H, x, H, 1, H, H, nullptr, x, H, 0, LP_ACT_NONE, 0, 128, stream));
    LP_TRY(lp_rmsnorm(x, H, L->ln2_g, h, H, 1, H, d->eps, stream));
    LP_TRY(lp_gemm_bf16(h, H, L->gateup_w, H, act, d->ffn, 1, 2 * d->ffn, H, nullptr, nullptr, 0, 0, LP_ACT_SWIGLU, 0,
                        128, stream));
    LP_TRY(lp_gemm_bf16(act, d->ffn, L->down_w, d->ffn, x, H, 1, H, d->ffn, nullptr, x, H, 0, LP_ACT_NONE, 0, 128,
                        stream));
  }
  LP_TRY(lp_rmsnorm(x, H, w->final_g, h, H, 1, H, d->eps, stream));
  LP_TRY(lp_gemm_bf16(h, H, w->lm_head, H, logits, d->vocab, 1, d->vocab, H, nullptr, nullptr, 0, 0, LP_ACT_NONE, 1, 128,
                      stream));
  return 0;
}

// GEMV implementation; position comes from the device (pos_dev) or the host (pos_dev == NULL)
static int decode_step_v2(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table, long long ld_embed,
                          const long long* token_id, const int* pos_dev, int pos_host, void* kv_cache,
                          long long kv_capacity, void* logits, void* workspace, cudaStream_t stream) {
  const int H = d->hidden, hd = d->head_dim;
  const int q_w = d->heads * hd, kv_w = d->kv_heads * hd, qkv_w = q_w + 2 * kv_w;
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t n) { void* r = p; p += (n + 255) & ~size_t(255); return r; };
  bf16* x = (bf16*)take(bf(H));
  bf16* qb = (bf16*)take(bf(q_w));
  bf16* att = (bf16*)take(bf(H));
  bf16* act = (bf16*)take(bf(d->ffn));
  float* part = (float*)take((size_t)d->heads * DEC_SPLITS * 130 * 4);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  int rc;
  LP_TRY(launch_pdl(decode_embed_kernel, dim3(1), dim3(256), 0, stream, token_id, (const bf16*)embed_table, ld_embed,
                    (long long)d->vocab, x, H));
  GemvArgs base = {};
  base.eps = d->eps;
  base.heads = d->heads; base.kv_heads = d->kv_heads; base.kv_capacity = kv_capacity;
  base.inv_freq = w->inv_freq; base.pos_dev = pos_dev; base.pos_host = pos_host;
  for (int l = 0; l < d->layers; ++l) {
    const lp_dec_layer_weights* L = &w->layers[l];
    bf16* kc = (bf16*)kv_cache + (size_t)l * 2 * kv_capacity * kv_w;
    bf16* vc = kc + (size_t)kv_capacity * kv_w;
    GemvArgs a = base;  // RMSNorm(ln1) -> QKV -> RoPE -> cache append
    a.W = (const bf16*)L->qkv_w; a.ldw = H; a.N = qkv_w; a.K = H; a.x = x; a.gamma = L->ln1_g;
    a.q_out = qb; a.kc = kc; a.vc = vc;
    LP_TRY(launch_gemv<GV_QKV>(a, stream));
    {
      ProfScope prof(8, pos_host + 1, d->heads, 0, stream);
      LP_TRY(launch_pdl(decode_attn_partial_kernel, dim3(d->heads, DEC_SPLITS), dim3(DEC_WARPS * 32), 0, stream,
                        (const bf16*)qb, (const bf16*)kc, (const bf16*)vc, pos_dev, pos_host, (int)kv_capacity, kv_w,
                        d->heads / d->kv_heads, scale_log2, part));
      LP_TRY(launch_pdl(decode_attn_combine_kernel, dim3(d->heads), dim3(128), 0, stream, (const float*)part, att));
    }
    a = base;  // x += o_proj(att)
    a.W = (const bf16*)L->o_w; a.ldw = H; a.N = H; a.K = H; a.x = att; a.y = x; a.residual = 1;
    LP_TRY(launch_gemv<GV_PLAIN>(a, stream));
    a = base;  // act = silu(gate) * up of RMSNorm(ln2)(x)
    a.W = (const bf16*)L->gateup_w; a.ldw = H; a.N = 2 * d->ffn; a.K = H; a.x = x; a.gamma = L->ln2_g; a.y = act;
    LP_TRY(launch_gemv<GV_SWIGLU>(a, stream));
    a = base;  // x += down(act)
    a.W = (const bf16*)L->down_w; a.ldw = d->ffn; a.N = H; a.K = d->ffn; a.x = act; a.y = x; a.residual = 1;
    LP_TRY(launch_gemv<GV_PLAIN>(a, stream));
  }
  GemvArgs a = base;  // final RMSNorm -> lm_head
  a.W = (const bf16*)w->lm_head; a.ldw = H; a.N = d->vocab; a.K = H; a.x = x; a.gamma = w->final_g;
  a.y32 = (float*)logits;
  LP_TRY(launch_gemv<GV_LOGITS>(a, stream));
  return 0;
}

static int decode_check(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table, const void* token_id,
                        const void* kv_cache, const void* logits, const void* workspace, size_t ws_bytes) {
  LP_REQUIRE(d && w && embed_table && token_id && kv_cache && logits && workspace, "lp_decode_step: null argument");
  LP_REQUIRE(d->head_dim == 128, "lp_decode_step: head_dim must be 128");
  LP_REQUIRE(d->vocab % 2 == 0 && d->hidden % 8 == 0 && d->ffn % 32 == 0 && d->ffn * 2 <= 48 * 1024,
             "lp_decode_step: unsupported widths hidden=%d ffn=%d vocab=%d", d->hidden, d->ffn, d->vocab);
  LP_REQUIRE(ws_bytes >= lp_decode_workspace_bytes(d), "lp_decode_step: workspace too small");
  return 0;
}

extern "C" int lp_decode_step(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table, long long ld_embed,
                              const long long* token_id, int position, void* kv_cache, long long kv_capacity,
                              void* logits, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc;
  LP_TRY(decode_check(d, w, embed_table, token_id, kv_cache, logits, workspace, ws_bytes));
  LP_REQUIRE(position >= 0 && position < kv_capacity, "lp_decode_step: position %d outside the cache capacity %lld",
             position, kv_capacity);
  if (!(g_decode_mode & 1))
    return decode_step_v1(d, w, embed_table, ld_embed, token_id, position, kv_cache, kv_capacity, logits, workspace,
                          stream);
  return decode_step_v2(d, w, embed_table, ld_embed, token_id, nullptr, position, kv_cache, kv_capacity, logits,
                        workspace, stream);
}

extern "C" int lp_decode_greedy_step(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table,
                                     long long ld_embed, long long* token_io, int* position_io, void* kv_cache,
                                     long long kv_capacity, void* logits, int vocab_valid, void* workspace,
                                     size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc;
  LP_TRY(decode_check(d, w, embed_table, token_io, kv_cache, logits, workspace, ws_bytes));
  LP_REQUIRE(position_io != nullptr && vocab_valid > 0 && vocab_valid <= d->vocab,
             "lp_decode_greedy_step: bad position pointer / vocab_valid=%d", vocab_valid);
  LP_TRY(decode_step_v2(d, w, embed_table, ld_embed, token_io, position_io, 0, kv_cache, kv_capacity, logits, workspace,
                        stream));
  LP_TRY(launch_pdl(decode_argmax_kernel, dim3(1), dim3(1024), 0, stream, (const float*)logits, vocab_valid, token_io,
                    position_io));
  return 0;
}

// ---------------------------------------------------------------------------------------------------- CUDA graph
// One greedy step = ~200 short kernels; replaying them as a graph removes the per-launch host cost (and lets
// programmatic dependent launch edges survive, mode bit 1).  Capture happens on a private non-blocking stream because
// the caller's stream may be the legacy default stream, which cannot be captured.
struct lp_decode_graph {
  cudaGraphExec_t exec;
  long long launches_per_replay;
};

extern "C" int lp_decode_graph_create(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table,
                                      long long ld_embed, long long* token_io, int* position_io, void* kv_cache,
                                      long long kv_capacity, void* logits, int vocab_valid, void* workspace,
                                      size_t ws_bytes, lp_decode_graph** out) {
  LP_REQUIRE(out != nullptr, "lp_decode_graph_create: null output handle");
  *out = nullptr;
  cudaStream_t cs;
  LP_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  const long long l0 = lp_launch_count();
  cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  int rc = 0;
  if (e == cudaSuccess) {
    rc = lp_decode_greedy_step(d, w, embed_table, ld_embed, token_io, position_io, kv_cache, kv_capacity, logits,
                               vocab_valid, workspace, ws_bytes, cs);
    e = cudaStreamEndCapture(cs, &graph);  // always end the capture, even after a failed enqueue
  }
  const long long per_replay = lp_launch_count() - l0;
  note_launch((int)-per_replay);  // captured, not executed: replays are counted by lp_decode_graph_launch
  cudaGraphExec_t exec = nullptr;
  if (rc == 0 && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  cudaStreamDestroy(cs);
  if (rc) return rc;
  if (e != cudaSuccess) {
    set_error("lp_decode_graph_create: capture/instantiate failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return LP_ERR_CUDA;
  }
  lp_decode_graph* g = new lp_decode_graph;
  g->exec = exec;
  g->launches_per_replay = per_replay;
  *out = g;
  return 0;
}

extern "C" int lp_decode_graph_launch(lp_decode_graph* g, void* stream) {
  LP_REQUIRE(g != nullptr && g->exec != nullptr, "lp_decode_graph_launch: null graph");
  LP_CHECK_CUDA(cudaGraphLaunch(g->exec, reinterpret_cast<cudaStream_t>(stream)));
  note_launch((int)g->launches_per_replay);
  return 0;
}

extern "C" void lp_decode_graph_destroy(lp_decode_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  delete g;
}
