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

static int g_decode_mode = 5;  // bit 0: GEMV kernels, bit 1: programmatic dependent launch, bit 2: block-cooperative GEMV, bit 3: half-SM footprints

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
constexpr int DEC_SPLITS = 48;   // key ranges per kv head -> kv_heads * DEC_SPLITS CTAs, all resident in one wave
constexpr int DEC_GROUP = 4;     // q heads per kv head handled by one CTA (Llama-3.1-8B / Mistral-7B: 32 / 8)
constexpr int DEC_STREAMS = DEC_WARPS / DEC_GROUP;  // warps per q head, interleaved over the CTA's key range

// Split-K GQA decode attention: CTA (kv head g, split s) scans keys [s*chunk, (s+1)*chunk) for the FOUR q heads that
// share kv head g.  Warp w serves q head 4g + (w & 3) on key stream (w >> 2); the four head-warps of a stream walk the
// same K/V rows in lock-step, so the rows come from HBM/L2 once per CTA and the other three reads hit L1 (the
// per-q-head version re-read every row four times from L2 and was L2-bandwidth bound: 16 us for 20 MB).
// A warp handles FOUR keys per step (8 lanes x 16 dims each, two 16-byte loads per lane, 3-step shuffle reduce) and
// requests the next step's rows before consuming the current ones.  Per-(head, split) partials (m, l, acc[128]) go to
// `part`; the LAST CTA of a kv head to finish (atomic ticket) merges its heads' partials and writes the bf16 output
// rows, so there is no separate combine launch.  `tickets` [kv_heads] must be zero on entry (decode_embed_kernel clears
// it every step) and is left zero.  Attends to rows 0 .. position (the row at `position` was appended by this step's
// QKV kernel).
__global__ void __launch_bounds__(DEC_WARPS * 32, 3)
decode_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kc, const bf16* __restrict__ vc,
                   const int* __restrict__ pos_dev, int pos_host, int max_ctx, int kv_w, float scale_log2,
                   float* __restrict__ part, unsigned* __restrict__ tickets, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int n_ctx = min(decode_position(pos_dev, pos_host) + 1, max_ctx);  // never read past the cache
  const int g = blockIdx.x, split = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hh = warp & (DEC_GROUP - 1), stream = warp / DEC_GROUP;
  const int h = g * DEC_GROUP + hh;
  const int sub = lane >> 3, l8 = lane & 7;  // key slot within the warp step, 16-dim slice of the head
  const unsigned gmask = 0xFFu << (sub * 8);
  const int chunk = (n_ctx + DEC_SPLITS - 1) / DEC_SPLITS;
  const int t_begin = split * chunk, t_end = min(n_ctx, t_begin + chunk);
  constexpr int SLOTS = DEC_STREAMS * 4;  // partial states per q head inside the CTA
  __shared__ float s_m[DEC_GROUP][SLOTS], s_l[DEC_GROUP][SLOTS], s_acc[DEC_GROUP][SLOTS][128];
  __shared__ float s_w[DEC_GROUP][DEC_SPLITS], s_inv[DEC_GROUP];
  __shared__ unsigned s_ticket;
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
  const long long col = (long long)g * 128 + l8 * 16;
  constexpr int STEP = DEC_STREAMS * 4;  // keys per CTA step
  int t = t_begin + stream * 4 + sub;
  uint4 k0 = make_uint4(0u, 0u, 0u, 0u), k1 = k0, v0 = k0, v1 = k0;
  if (t < t_end) {
    const bf16* kp = kc + (long long)t * kv_w + col;
    const bf16* vp = vc + (long long)t * kv_w + col;
    k0 = *reinterpret_cast<const uint4*>(kp); k1 = *reinterpret_cast<const uint4*>(kp + 8);
    v0 = *reinterpret_cast<const uint4*>(vp); v1 = *reinterpret_cast<const uint4*>(vp + 8);
  }
  while (t < t_end) {
    const int tn = t + STEP;
    uint4 nk0 = k0, nk1 = k1, nv0 = v0, nv1 = v1;
    if (tn < t_end) {  // request the next key/value rows before consuming the current ones
      const bf16* kp = kc + (long long)tn * kv_w + col;
      const bf16* vp = vc + (long long)tn * kv_w + col;
      nk0 = *reinterpret_cast<const uint4*>(kp); nk1 = *reinterpret_cast<const uint4*>(kp + 8);
      nv0 = *reinterpret_cast<const uint4*>(vp); nv1 = *reinterpret_cast<const uint4*>(vp + 8);
    }
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
    k0 = nk0; k1 = nk1; v0 = nv0; v1 = nv1;
    t = tn;
  }
  // note: a lane whose key slots were all beyond t_end keeps m = -inf, l = 0 (handled by the merge weights)
  const int slot = stream * 4 + sub;
  if (l8 == 0) { s_m[hh][slot] = m; s_l[hh][slot] = l; }
#pragma unroll
  for (int i = 0; i < 16; ++i) s_acc[hh][slot][l8 * 16 + i] = acc[i];
  __syncthreads();
  {  // merge the CTA's slots: thread -> (head = tid / 64, two adjacent dims)
    const int mh = threadIdx.x >> 6, d2 = (threadIdx.x & 63) * 2;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < SLOTS; ++w) mm = fmaxf(mm, s_m[mh][w]);
    float ll = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < SLOTS; ++w) {
      const float c = (s_m[mh][w] == -INFINITY) ? 0.f : exp2f(s_m[mh][w] - mm);
      ll += s_l[mh][w] * c;
      o0 += s_acc[mh][w][d2] * c;
      o1 += s_acc[mh][w][d2 + 1] * c;
    }
    float* pp = part + ((long long)(g * DEC_GROUP + mh) * DEC_SPLITS + split) * 130;
    *reinterpret_cast<float2*>(pp + 2 + d2) = make_float2(o0, o1);
    if (d2 == 0) { pp[0] = mm; pp[1] = ll; }
  }
  // ---- last CTA of this kv head merges the partials (release: fence + ticket; acquire: ticket + fence)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[g], 1u);
  __syncthreads();
  if (s_ticket != DEC_SPLITS - 1) return;
  __threadfence();
  const float* pg = part + (long long)g * DEC_GROUP * DEC_SPLITS * 130;
  if (threadIdx.x < DEC_GROUP * DEC_SPLITS) s_w[threadIdx.x / DEC_SPLITS][threadIdx.x % DEC_SPLITS] =
      __ldcg(pg + (long long)threadIdx.x * 130);  // m of (head, split)
  __syncthreads();
  if (threadIdx.x < DEC_GROUP) {  // weights exp2(m_s - max) and 1 / sum(l_s w_s) per head
    const int mh = threadIdx.x;
    float mm = -INFINITY;
    for (int s2 = 0; s2 < DEC_SPLITS; ++s2) mm = fmaxf(mm, s_w[mh][s2]);
    float ll = 0.f;
    for (int s2 = 0; s2 < DEC_SPLITS; ++s2) {
      const float ms = s_w[mh][s2];
      const float c = (ms == -INFINITY) ? 0.f : exp2f(ms - mm);
      s_w[mh][s2] = c;
      ll += __ldcg(pg + ((long long)mh * DEC_SPLITS + s2) * 130 + 1) * c;
    }
    s_inv[mh] = 1.0f / ll;
  }
  __syncthreads();
  {
    const int mh = threadIdx.x >> 6, d2 = (threadIdx.x & 63) * 2;
    const float* po = pg + (long long)mh * DEC_SPLITS * 130 + 2 + d2;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int s0 = 0; s0 < DEC_SPLITS; s0 += 16) {  // 16 independent 8-byte loads in flight per round
      float2 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldcg(reinterpret_cast<const float2*>(po + (long long)(s0 + i) * 130));
#pragma unroll
      for (int i = 0; i < 16; ++i) { o0 += v[i].x * s_w[mh][s0 + i]; o1 += v[i].y * s_w[mh][s0 + i]; }
    }
    const float inv = s_inv[mh];
    *reinterpret_cast<uint32_t*>(out + (g * DEC_GROUP + mh) * 128 + d2) = pack_bf16x2(o0 * inv, o1 * inv);
  }
  if (threadIdx.x == 0) tickets[g] = 0u;  // ready for the next layer / step
}

// ---------------------------------------------------------------------------------------------------- token plumbing
// x = table[clamp(token)]  (Embedding.forward, MP/model/llava/vlm_model.py:224-254; HF get_input_embeddings, EV :263)
__global__ void __launch_bounds__(256)
decode_embed_kernel(const long long* __restrict__ token, const bf16* __restrict__ table, long long ldt, long long vocab,
                    bf16* __restrict__ x, int H, unsigned* __restrict__ tickets, int n_tickets) {
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < n_tickets; i += blockDim.x) tickets[i] = 0u;  // attention merge tickets
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

// ---------------------------------------------------------------------------------------------------- GEMV, block-cooperative
// The warp-per-row-pair kernel above walks a row in four dependent load rounds, which costs ~10 us of latency chain
// per launch on the small projections (ncu: 2.7-3.0 TB/s on the 34-50 MB matrices, 6.5 TB/s only on lm_head).  Here a
// whole CTA reduces R rows at once with the K dimension split over its 256 threads: thread t owns the 16-byte chunks
// t, t+256, ... of EVERY row, so (a) all R x KCH loads of a block of rows are in flight together (one DRAM round trip
// per block, 64 KB per CTA), (b) the activation chunks a thread needs never change and live in registers (the RMSNorm
// prologue is thread-local apart from one block reduction), (c) the first block's weights are requested before the
// activation row is even read (and before griddepcontrol.wait).  The next block's loads are issued as soon as the
// dot products have consumed the registers, so they fly during the cross-warp reduction and the epilogue.
// Row blocks pair each "primary" row with a partner at distance P (64: rotate-half RoPE, 32: gate/up, else adjacent).
template <int MODE, int R>
__device__ __forceinline__ void gemv3_block_rows(const GemvArgs& a, int blk, int& r0, int& P) {
  constexpr int H = R / 2;
  if (MODE == GV_QKV) {
    const int per_head = 64 / H;  // blocks per q/k head
    const int n_rot = (a.heads + a.kv_heads) * per_head;
    if (blk < n_rot) { r0 = (blk / per_head) * 128 + (blk % per_head) * H; P = 64; }
    else { r0 = (a.heads + a.kv_heads) * 128 + R * (blk - n_rot); P = H; }
  } else if (MODE == GV_SWIGLU) {
    const int per_gran = 32 / H;
    r0 = (blk / per_gran) * 64 + (blk % per_gran) * H; P = 32;
  } else {
    r0 = R * blk; P = H;
  }
}

// XS: keep the activation row in shared memory instead of registers (long rows: K = 14336 would cost 28 registers)
// THREADS / MINB: CTA size and CTAs per SM the register budget is compiled for.  The <256, 4> and <512, 2> shapes
// use at most half of an SM's registers and threads at one wave, so that with programmatic dependent launch the NEXT
// kernel's CTAs are resident (and have their first weight rows in flight) while this one is still streaming.
template <int MODE, int KCH, int R, bool XS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
gemv3_kernel(const GemvArgs a) {
  constexpr int GV_THREADS = THREADS;
  constexpr int GV_WARPS = THREADS / 32;
  constexpr int H = R / 2;
  extern __shared__ uint4 xs3[];  // XS only: K / 8 chunks
  __shared__ float s_red[2][GV_WARPS][R];
  __shared__ float s_ss[GV_WARPS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K8 = a.K >> 3;
  const int n_blocks = a.N / R;
  uint4 wreg[R][KCH];

  auto load_block = [&](int blk) {
    int r0, P;
    gemv3_block_rows<MODE, R>(a, blk, r0, P);
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int row = r0 + (rr < H ? rr : P + rr - H);
      const uint4* wp = reinterpret_cast<const uint4*>(a.W + (long long)row * a.ldw);
#pragma unroll
      for (int j = 0; j < KCH; ++j) {
        const int c = tid + j * GV_THREADS;
        wreg[rr][j] = (c < K8) ? ld_stream16(wp + c) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };

  int blk = blockIdx.x;
  if (blk < n_blocks) load_block(blk);  // weights do not depend on the previous kernel
  pdl_launch_dependents();
  pdl_wait();

  // ---- this thread's activation chunks (optionally RMS-normalised), kept as packed bf16
  uint4 xr[XS ? 1 : KCH];
  if (XS) {  // no RMSNorm on this path (down projection): plain copy, each thread later reads back its own chunks
    const uint4* xg = reinterpret_cast<const uint4*>(a.x);
    for (int c = tid; c < K8; c += GV_THREADS) xs3[c] = xg[c];
    __syncthreads();
  } else {
    const uint4* xg = reinterpret_cast<const uint4*>(a.x);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      const int c = tid + j * GV_THREADS;
      xr[j] = (c < K8) ? xg[c] : make_uint4(0u, 0u, 0u, 0u);
      const float2 p = unpack_bf16x2(xr[j].x), q = unpack_bf16x2(xr[j].y), r = unpack_bf16x2(xr[j].z),
                   s = unpack_bf16x2(xr[j].w);
      ss += p.x * p.x + p.y * p.y + q.x * q.x + q.y * q.y + r.x * r.x + r.y * r.y + s.x * s.x + s.y * s.y;
    }
    if (a.gamma != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) s_ss[warp] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) tot += s_ss[w];
      const float rstd = rsqrtf(tot / (float)a.K + a.eps);
#pragma unroll
      for (int j = 0; j < KCH; ++j) {
        const int c = tid + j * GV_THREADS;
        if (c < K8) {
          const float4 g0 = *reinterpret_cast<const float4*>(a.gamma + c * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(a.gamma + c * 8 + 4);
          const float2 p = unpack_bf16x2(xr[j].x), q = unpack_bf16x2(xr[j].y), r = unpack_bf16x2(xr[j].z),
                       s = unpack_bf16x2(xr[j].w);
          xr[j].x = pack_bf16x2(p.x * rstd * g0.x, p.y * rstd * g0.y);
          xr[j].y = pack_bf16x2(q.x * rstd * g0.z, q.y * rstd * g0.w);
          xr[j].z = pack_bf16x2(r.x * rstd * g1.x, r.y * rstd * g1.y);
          xr[j].w = pack_bf16x2(s.x * rstd * g1.z, s.y * rstd * g1.w);
        }
      }
    }
  }
  const int pos = (MODE == GV_QKV) ? decode_position(a.pos_dev, a.pos_host) : 0;

  int it = 0;
  for (; blk < n_blocks; blk += gridDim.x, ++it) {
    float acc[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) acc[rr] = 0.f;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      uint4 xv;
      if (XS) { const int c = tid + j * GV_THREADS; xv = (c < K8) ? xs3[c] : make_uint4(0u, 0u, 0u, 0u); }
      else xv = xr[j];
      const float2 p = unpack_bf16x2(xv.x), q = unpack_bf16x2(xv.y), r = unpack_bf16x2(xv.z), s = unpack_bf16x2(xv.w);
      const float xf[8] = {p.x, p.y, q.x, q.y, r.x, r.y, s.x, s.y};
#pragma unroll
      for (int rr = 0; rr < R; ++rr) acc[rr] = dot8(wreg[rr][j], xf, acc[rr]);
    }
    const int next = blk + gridDim.x;
    if (next < n_blocks) load_block(next);  // in flight during the reduction and the epilogue below
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[rr] += __shfl_xor_sync(0xffffffffu, acc[rr], o);
    }
    float(*red)[R] = s_red[it & 1];
    if (lane == 0) {
#pragma unroll
      for (int rr = 0; rr < R; ++rr) red[warp][rr] = acc[rr];
    }
    __syncthreads();  // one barrier per block: the partials are double-buffered by iteration parity
    if (tid < H) {
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) { v0 += red[w][tid]; v1 += red[w][tid + H]; }
      int r0, P;
      gemv3_block_rows<MODE, R>(a, blk, r0, P);
      const int ra = r0 + tid, rb = r0 + P + tid;
      if (MODE == GV_PLAIN) {
        if (a.residual) { v0 += __bfloat162float(a.y[ra]); v1 += __bfloat162float(a.y[rb]); }
        a.y[ra] = __float2bfloat16(v0);
        a.y[rb] = __float2bfloat16(v1);
      } else if (MODE == GV_LOGITS) {
        a.y32[ra] = v0;
        a.y32[rb] = v1;
      } else if (MODE == GV_SWIGLU) {
        a.y[(ra >> 6) * 32 + (ra & 31)] = __float2bfloat16(silu_f(v0) * v1);
      } else {  // GV_QKV
        const int q_w = a.heads * 128, kv_w = a.kv_heads * 128;
        if (pos >= 0 && pos < a.kv_capacity) {  // host validates; never write outside the cache
          if (ra < q_w + kv_w) {
            const float t1 = __bfloat162float(__float2bfloat16(v0)), t2 = __bfloat162float(__float2bfloat16(v1));
            const float ang = (float)pos * a.inv_freq[ra & 63];
            const float cs = cosf(ang), sn = sinf(ang);
            const bf16 o1 = __float2bfloat16(t1 * cs - t2 * sn), o2 = __float2bfloat16(t2 * cs + t1 * sn);
            if (ra < q_w) { a.q_out[ra] = o1; a.q_out[rb] = o2; }
            else {
              bf16* kp = a.kc + (long long)pos * kv_w + (ra - q_w);
              kp[0] = o1; kp[64] = o2;
            }
          } else {
            bf16* vp = a.vc + (long long)pos * kv_w + (ra - q_w - kv_w);
            vp[0] = __float2bfloat16(v0);
            vp[P] = __float2bfloat16(v1);
          }
        }
      }
    }
  }
}

template <int MODE, int KCH, int R, bool XS, int THREADS, int MINB, int CTAS_PER_SM>
static int launch_gemv3(const GemvArgs& a, cudaStream_t stream) {
  const int n_blocks = a.N / R;
  int grid = n_blocks;
  const int cap = sm_count() * CTAS_PER_SM;
  if (grid > cap) grid = cap;
  ProfScope prof(10 + MODE, a.N, a.K, 3, stream);
  return launch_pdl(gemv3_kernel<MODE, KCH, R, XS, THREADS, MINB>, dim3(grid), dim3(THREADS),
                    XS ? (size_t)a.K * 2 : 0, stream, a);
}

template <int MODE>
static int launch_gemv(const GemvArgs& a, cudaStream_t stream) {
  if (g_decode_mode & 4) {  // block-cooperative kernels for the Llama-3.1-8B / Mistral-7B widths
    const bool n_ok = (MODE == GV_SWIGLU) ? (a.N % 64 == 0) : (a.N % 8 == 0);
    const bool half_sm = (g_decode_mode & 8) != 0;  // footprints that leave room for the dependent kernel
    if (a.K == 4096 && n_ok && a.ldw % 8 == 0)
      return half_sm ? launch_gemv3<MODE, 2, 4, false, 256, 4, 2>(a, stream)
                     : launch_gemv3<MODE, 2, 8, false, 256, 2, 2>(a, stream);
    if constexpr (MODE == GV_PLAIN) {
      if (a.K == 14336 && a.N % 2 == 0 && a.ldw % 8 == 0 && a.gamma == nullptr)
        return half_sm ? launch_gemv3<MODE, 4, 2, true, 512, 2, 1>(a, stream)
                       : launch_gemv3<MODE, 7, 2, true, 256, 2, 2>(a, stream);
    }
  }

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
  g_decode_mode = mode & 15;
  return old;
}

extern "C" size_t lp_decode_workspace_bytes(const lp_dec_dims* d) {
  const size_t qkv_w = (size_t)(d->heads + 2 * d->kv_heads) * d->head_dim;
  return 256 * 9 + bf(2 * (size_t)d->hidden) * 2 + bf(qkv_w) + bf((size_t)d->ffn) + (size_t)d->head_dim * 4 + 64 * 4 +
         (size_t)d->heads * DEC_SPLITS * 130 * 4 + (size_t)d->heads * 4 + 256;
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
  unsigned* tickets = (unsigned*)take((size_t)d->heads * 4);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  int rc;
  LP_CHECK_CUDA(cudaMemsetAsync(tickets, 0, (size_t)d->heads * 4, stream));
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
      note_launch();
      decode_attn_kernel<<<dim3(d->kv_heads, DEC_SPLITS), DEC_WARPS * 32, 0, stream>>>(
          qkv, kc, vc, nullptr, position, (int)kv_capacity, kv_w, scale_log2, part, tickets, att);
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
  unsigned* tickets = (unsigned*)take((size_t)d->heads * 4);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  int rc;
  LP_TRY(launch_pdl(decode_embed_kernel, dim3(1), dim3(256), 0, stream, token_id, (const bf16*)embed_table, ld_embed,
                    (long long)d->vocab, x, H, tickets, d->heads));
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
      LP_TRY(launch_pdl(decode_attn_kernel, dim3(d->kv_heads, DEC_SPLITS), dim3(DEC_WARPS * 32), 0, stream,
                        (const bf16*)qb, (const bf16*)kc, (const bf16*)vc, pos_dev, pos_host, (int)kv_capacity, kv_w,
                        scale_log2, part, tickets, att));
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
  LP_REQUIRE(d->head_dim == 128 && d->heads == DEC_GROUP * d->kv_heads,
             "lp_decode_step: head_dim must be 128 with 4 query heads per kv head (got %d / %d heads)", d->heads,
             d->kv_heads);
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
