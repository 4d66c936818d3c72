// leopard_b200 — single-token decode step over the prefill's KV cache (greedy generation after the prefill).
//
// The cached-decode branch of myLlavaForConditionalGeneration.forward
// (evaluations/models/llava_multiimg_siglip_anyres.py:291-320, driven by `generate(..., max_new_tokens=128,
// use_cache=True)` at :448-452): with B = 1 and no padding the rebuilt mask is all ones and the position is the
// number of cached rows.
//
// The step is HBM-bound: every decoder weight (15.0 GB for Llama-3.1-8B incl. lm_head) is read exactly once per token,
// so the floor is weights / HBM bandwidth (~2.4 ms on a B200) and the design is about keeping the memory system
// saturated across ~160 short kernels (measured: 3.2 ms/token at S = 4796 = 75 % of the HBM copy peak):
//   gemv3_kernel<MODE,..> y = W x for ONE activation row, block-cooperative: a CTA reduces 8 weight rows at a time with
//                       K split over its 256 threads — 16 x 16-byte non-allocating loads in flight per thread, the
//                       activation chunks in registers (default, mode bit 2).
//   gemv_kernel<MODE>   the first GEMV: a warp owns two weight rows and walks them in four load rounds (mode bit 2 = 0).
//                       Both: fused prologue RMSNorm of the residual stream (MLM/legacy/model/rms_norm.py:26-31);
//                       fused epilogues QKV -> rotate-half RoPE on (i, i+64) row pairs + append K/V to the cache
//                       (MP/model/llava/transformer.py:853-856); gate/up -> silu(gate)*up (:129-139); o/down -> residual
//                       add in place; lm_head -> fp32 logits.
//   decode_attn_kernel  split-K GQA attention over the bf16 KV cache: CTA = (kv head, 1 of 37 key ranges) for the four
//                       q heads of the group, cp.async K/V pipeline, last CTA per kv head merges the partials.
//   argmax / embed      greedy token selection and the next embedding row, device-resident, so that a whole step
//                       (lp_decode_greedy_step) reads its position and token from device memory and can be captured
//                       once into a CUDA graph (lp_decode_graph_*) and replayed per token with no host work.
//   decode_mega_kernel  the same step as ONE persistent cooperative kernel with grid barriers (mode bit 4; slower).
// All kernels start with griddepcontrol.wait and may be launched with programmatic dependent launch (mode bit 1): the
// next kernel's CTAs request their first weight rows while the previous kernel drains.
// Mode bit 0 = 0 selects the first implementation (M = 1 rows through the tcgen05 GEMM), kept for A/B measurements.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

// MiB of the o_proj / gate-up weights the decode attention kernel pulls into L2 while it runs (lp_decode_set_prefetch)
static int g_dec_prefetch_mb[2] = {32, 24};   // profiles/r02p_decode_prefetch.jsonl: 3.14 ms/token vs 3.18-3.25 without (noise +-0.1)

int rope_table_base(int pos_base, const float* inv_freq, float* cs, float* sn, int S, int half, cudaStream_t stream);

static int g_decode_mode = 7;  // bit 0: GEMV kernels, bit 1: programmatic dependent launch, bit 2: block-cooperative GEMV, bit 4: one persistent kernel per token


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
constexpr int DEC_SPLITS = 37;   // key ranges per kv head: 8 x 37 = 296 CTAs = one wave at 2 CTAs per SM (82 KB smem each)
constexpr int DEC_GROUP = 4;     // q heads per kv head handled by one CTA (Llama-3.1-8B / Mistral-7B: 32 / 8)
constexpr int DA_TILE = 32;      // keys per shared-memory tile
constexpr int DA_STAGES = 4;     // cp.async pipeline depth: 4 x (8 KB K + 8 KB V) in flight per CTA
constexpr int DA_STAGE_BYTES = 2 * DA_TILE * 256;
constexpr int DA_SMEM_BYTES = DA_STAGES * DA_STAGE_BYTES;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint4 ld_cg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float dec_ex2(float x) {  // MUFU.EX2 (rel. error 2^-22): inputs are already in log2 units
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int SPLITS>
struct DecAttnSmem {
  float m[DEC_GROUP][8], l[DEC_GROUP][8], acc[DEC_GROUP][8][128];
  float w[DEC_GROUP][SPLITS], l2[DEC_GROUP][SPLITS], inv[DEC_GROUP];
  unsigned ticket;
};

// Split-K GQA decode attention for work item (kv head g, key range `split` of SPLITS): keys [s*chunk, (s+1)*chunk) for the
// FOUR q heads that share kv head g.
//  * K/V rows are staged in shared memory by a 4-deep cp.async pipeline of 32-key tiles, so a CTA has up to 64 KB of
//    UNIQUE cache bytes in flight (the register-prefetch version had the four head-warps request the same rows: 8 KB of
//    unique bytes in flight per CTA, ~1.6 TB/s chip-wide, 12-16 dependent DRAM round trips per launch).
//  * Warp w serves q head 4g + (w & 3) on key stream (w >> 2): 16 keys of every tile, FOUR keys per step (8 lanes x 16
//    dims each: 16-byte chunks l8 and l8 + 8 of the 256-byte row, conflict-free), 3-step shuffle reduce.
//  * Per-(head, split) partials (m, l, acc[128]) go to `part`; the LAST CTA of a kv head to finish (atomic ticket) merges
//    its heads' partials and writes the bf16 output rows — no separate combine launch.  `tickets` [kv_heads] must be
//    zero on entry and is left zero.
// q / part / tickets are read L2-coherently (they may have been written earlier in the same kernel, see the
// single-kernel step).  Attends to rows 0 .. n_ctx - 1.
template <int SPLITS>
__device__ __forceinline__ void decode_attn_item(const bf16* q, const bf16* __restrict__ kc, const bf16* __restrict__ vc,
                                                 int n_ctx, int kv_w, int g, int split, float scale_log2, float* part,
                                                 unsigned* tickets, bf16* out, uint8_t* dsm, DecAttnSmem<SPLITS>& sm) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hh = warp & (DEC_GROUP - 1), stream = warp / DEC_GROUP;
  const int h = g * DEC_GROUP + hh;
  const int sub = lane >> 3, l8 = lane & 7;  // key slot within the warp step, dims [8 l8, +8) and [64 + 8 l8, +8)
  const unsigned gmask = 0xFFu << (sub * 8);
  const int chunk = (n_ctx + SPLITS - 1) / SPLITS;
  const int t_begin = split * chunk, t_end = min(n_ctx, t_begin + chunk);
  const int n_tiles = t_end > t_begin ? (t_end - t_begin + DA_TILE - 1) / DA_TILE : 0;
  auto issue = [&](int i) {  // tile i -> stage i % DA_STAGES; always commits (possibly empty) to keep group counts uniform
    if (i < n_tiles) {
      const int t0 = t_begin + i * DA_TILE;
      uint8_t* st = dsm + (i % DA_STAGES) * DA_STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = tid + k * 256, row = c >> 4, ch = c & 15;
        const int t = t0 + row;
        if (t < t_end) {
          const long long off = (long long)t * kv_w + g * 128 + ch * 8;
          cp_async16(st + row * 256 + ch * 16, kc + off);
          cp_async16(st + DA_TILE * 256 + row * 256 + ch * 16, vc + off);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < DA_STAGES - 1; ++i) issue(i);
  float qv[16];
  {
    const uint4 u0 = ld_cg16(q + h * 128 + l8 * 8), u1 = ld_cg16(q + h * 128 + 64 + l8 * 8);
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
  for (int i = 0; i < n_tiles; ++i) {
    cp_async_wait<DA_STAGES - 2>();
    __syncthreads();               // tile i has landed for every thread; everyone is done with tile i - 1
    issue(i + DA_STAGES - 1);      // refill the stage tile i - 1 occupied
    const uint8_t* st = dsm + (i % DA_STAGES) * DA_STAGE_BYTES;
    const int t0 = t_begin + i * DA_TILE;
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      const int row = stream * 16 + step * 4 + sub;
      if (t0 + row < t_end) {  // uniform over the 8 lanes that share this key
        const uint8_t* kp = st + row * 256 + l8 * 16;
        const uint8_t* vp = kp + DA_TILE * 256;
        const uint4 k0 = *reinterpret_cast<const uint4*>(kp), k1 = *reinterpret_cast<const uint4*>(kp + 128);
        const uint4 v0 = *reinterpret_cast<const uint4*>(vp), v1 = *reinterpret_cast<const uint4*>(vp + 128);
        const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 f = unpack_bf16x2(kw[e]);
          s += qv[2 * e] * f.x + qv[2 * e + 1] * f.y;
        }
        s += __shfl_xor_sync(gmask, s, 4);
        s += __shfl_xor_sync(gmask, s, 2);
        s += __shfl_xor_sync(gmask, s, 1);
        const float mn = fmaxf(m, s);
        const float corr = dec_ex2(m - mn), p = dec_ex2(s - mn);
        const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        l = l * corr + p;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 f = unpack_bf16x2(vw[e]);
          acc[2 * e] = acc[2 * e] * corr + p * f.x;
          acc[2 * e + 1] = acc[2 * e + 1] * corr + p * f.y;
        }
        m = mn;
      }
    }
  }
  cp_async_wait<0>();
  // note: a lane whose key slots were all beyond t_end keeps m = -inf, l = 0 (handled by the merge weights)
  const int slot = stream * 4 + sub;
  if (l8 == 0) { sm.m[hh][slot] = m; sm.l[hh][slot] = l; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sm.acc[hh][slot][l8 * 8 + i] = acc[i];
    sm.acc[hh][slot][64 + l8 * 8 + i] = acc[8 + i];
  }
  __syncthreads();
  {  // merge the CTA's 8 slots per head: thread -> (head = tid / 64, two adjacent dims)
    const int mh = tid >> 6, d2 = (tid & 63) * 2;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) mm = fmaxf(mm, sm.m[mh][w]);
    float ll = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float c = (sm.m[mh][w] == -INFINITY) ? 0.f : dec_ex2(sm.m[mh][w] - mm);
      ll += sm.l[mh][w] * c;
      o0 += sm.acc[mh][w][d2] * c;
      o1 += sm.acc[mh][w][d2 + 1] * c;
    }
    float* pp = part + ((long long)(g * DEC_GROUP + mh) * SPLITS + split) * 130;
    *reinterpret_cast<float2*>(pp + 2 + d2) = make_float2(o0, o1);
    if (d2 == 0) { pp[0] = mm; pp[1] = ll; }
  }
  // ---- last CTA of this kv head merges the partials (release: fence + ticket; acquire: ticket + fence)
  __threadfence();
  __syncthreads();
  if (tid == 0) sm.ticket = atomicAdd(&tickets[g], 1u);
  __syncthreads();
  if (sm.ticket != SPLITS - 1) return;
  __threadfence();
  const float* pg = part + (long long)g * DEC_GROUP * SPLITS * 130;
  if (tid < DEC_GROUP * SPLITS) {  // (m, l) of every (head, split): one parallel round of 8-byte loads
    const float2 ml = __ldcg(reinterpret_cast<const float2*>(pg + (long long)tid * 130));
    sm.w[tid / SPLITS][tid % SPLITS] = ml.x;
    sm.l2[tid / SPLITS][tid % SPLITS] = ml.y;
  }
  __syncthreads();
  if (tid < DEC_GROUP) {  // weights exp2(m_s - max) and 1 / sum(l_s w_s) per head (shared memory only)
    float mm = -INFINITY;
    for (int s2 = 0; s2 < SPLITS; ++s2) mm = fmaxf(mm, sm.w[tid][s2]);
    float ll = 0.f;
    for (int s2 = 0; s2 < SPLITS; ++s2) {
      const float ms = sm.w[tid][s2];
      const float c = (ms == -INFINITY) ? 0.f : dec_ex2(ms - mm);
      sm.w[tid][s2] = c;
      ll += sm.l2[tid][s2] * c;
    }
    sm.inv[tid] = 1.0f / ll;
  }
  __syncthreads();
  {
    const int mh = tid >> 6, d2 = (tid & 63) * 2;
    const float* po = pg + (long long)mh * SPLITS * 130 + 2 + d2;
    float o0 = 0.f, o1 = 0.f;
    float2 v[SPLITS];  // every partial of this (head, dim pair) requested at once: one L2 round trip instead of five
#pragma unroll
    for (int s2 = 0; s2 < SPLITS; ++s2) v[s2] = __ldcg(reinterpret_cast<const float2*>(po + (long long)s2 * 130));
#pragma unroll
    for (int s2 = 0; s2 < SPLITS; ++s2) {
      o0 += v[s2].x * sm.w[mh][s2];
      o1 += v[s2].y * sm.w[mh][s2];
    }
    const float inv = sm.inv[mh];
    *reinterpret_cast<uint32_t*>(out + (g * DEC_GROUP + mh) * 128 + d2) = pack_bf16x2(o0 * inv, o1 * inv);
  }
  if (tid == 0) tickets[g] = 0u;  // ready for the next layer / step
}

// multi-kernel step: one CTA per (kv head, key range).
// pf_a / pf_b: weights of the NEXT projections (o_proj, the head of gate/up) to pull into L2 while this kernel runs: attention reads
// only the KV cache (20 MB per layer at S = 4796, latency-bound split-K), so the HBM is nearly idle for its ~19 us — time in which
// ~100 MB of weights can cross it.  The lines are requested before griddepcontrol.wait (weights do not depend on earlier kernels).
__global__ void __launch_bounds__(DEC_WARPS * 32, 2)
decode_attn_kernel(const bf16* q, const bf16* __restrict__ kc, const bf16* __restrict__ vc,
                   const int* __restrict__ pos_dev, int pos_host, int max_ctx, int kv_w, float scale_log2,
                   float* part, unsigned* tickets, bf16* out, const uint8_t* pf_a, long long pf_a_bytes, const uint8_t* pf_b,
                   long long pf_b_bytes) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  __shared__ DecAttnSmem<DEC_SPLITS> sm;
  pdl_launch_dependents();
  {
    const long long nthr = (long long)gridDim.x * gridDim.y * blockDim.x;
    const long long t = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    for (long long i = t * 128; i < pf_a_bytes; i += nthr * 128) prefetch_l2(pf_a + i);
    for (long long i = t * 128; i < pf_b_bytes; i += nthr * 128) prefetch_l2(pf_b + i);
  }
  pdl_wait();
  const int n_ctx = min(decode_position(pos_dev, pos_host) + 1, max_ctx);  // never read past the cache
  decode_attn_item<DEC_SPLITS>(q, kc, vc, n_ctx, kv_w, blockIdx.x, blockIdx.y, scale_log2, part, tickets, out, da_smem,
                               sm);
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
    // (half-SM footprints — R = 4 at 64 registers / 512-thread CTAs, so that the dependent kernel is co-resident under
    //  programmatic dependent launch — were measured slower: 3.61 / 4.18 ms per token without / with PDL vs 3.37 / 3.28)
    if (a.K == 4096 && n_ok && a.ldw % 8 == 0) return launch_gemv3<MODE, 2, 8, false, 256, 2, 2>(a, stream);
    if constexpr (MODE == GV_PLAIN) {
      if (a.K == 14336 && a.N % 2 == 0 && a.ldw % 8 == 0 && a.gamma == nullptr)
        return launch_gemv3<MODE, 7, 2, true, 256, 2, 2>(a, stream);
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

// ---------------------------------------------------------------------------------------------------- one kernel per token
// The multi-kernel step spends ~25 % of its time in launch gaps, grid ramp-up and tails (5 dependent kernels per
// layer, each moving 20-235 MB).  decode_mega_kernel runs the WHOLE greedy step as one persistent cooperative grid
// (2 CTAs per SM): the same phases — RMSNorm+QKV+RoPE+cache append, GQA attention, o-proj, gate/up+SwiGLU, down, ...,
// lm_head, arg-max — separated by grid-wide barriers (one atomic arrival per CTA on a monotonic counter, ~2 us), and
// every GEMV phase's first block of weight rows is requested BEFORE the barrier that precedes it, so HBM keeps
// streaming while the grid synchronises.  Activations written by other CTAs inside this kernel (x, q, att, act,
// partials) are read with ld.global.cg (L2-coherent; L1 may hold stale lines from an earlier phase); weights use
// non-allocating read-only loads; K/V rows are read once per launch, so their L1 lines cannot be stale.
constexpr int MEGA_THREADS = 256;
constexpr int MEGA_WARPS = MEGA_THREADS / 32;
constexpr int MEGA_MAX_LAYERS = 32;

struct MegaLayer {
  const bf16 *qkv_w, *o_w, *gateup_w, *down_w;
  const float *ln1_g, *ln2_g;
};
struct MegaArgs {
  MegaLayer layers[MEGA_MAX_LAYERS];
  int n_layers, hidden, ffn, heads, kv_heads, vocab, vocab_valid;
  float eps, scale_log2;
  const bf16* lm_head;
  const float *final_g, *inv_freq;
  const bf16* embed;
  long long ld_embed, embed_rows;
  bf16* kv_cache;
  long long kv_capacity;
  long long* token_io;
  int* position_io;
  float* logits;
  bf16 *x, *qb, *att, *act;  // workspace
  float* part;
  unsigned *tickets, *barrier;
  float* cand_v;
  int* cand_i;
  unsigned long long* dbg;  // optional: globaltimer stamps of CTA 0 around every grid barrier (lp_decode_debug_stamps)
};

__device__ __forceinline__ unsigned long long mega_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid-wide barrier on a monotonic counter (zeroed by the host before the launch); all CTAs are co-resident
// (cooperative launch).  Bounded spin: a wedged grid traps instead of hanging the GPU.
__device__ __forceinline__ void mega_grid_sync(unsigned* counter, unsigned& epoch, unsigned long long* dbg = nullptr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    if (dbg != nullptr && blockIdx.x == 0) dbg[2 * epoch] = mega_globaltimer();
    const unsigned target = epoch * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    if (ld_acquire_u32(counter) < target) {
      const long long t0 = clock64();
      while (ld_acquire_u32(counter) < target) {
        if (clock64() - t0 > LP_WATCHDOG_CYCLES) {
          printf("lp: decode grid barrier watchdog: block %d epoch %u\n", (int)blockIdx.x, epoch);
          __trap();
        }
      }
    }
    __threadfence();
    if (dbg != nullptr && blockIdx.x == 0) dbg[2 * epoch + 1] = mega_globaltimer();
  }
  __syncthreads();
}

struct MegaGemv {  // one GEMV phase (same meaning as GemvArgs)
  const bf16* W;
  long long ldw;
  int N, K;
  const bf16* x;
  const float* gamma;
  bf16* y;
  int residual;
  float* y32;
  bf16 *kc, *vc;
};

// request the weight rows of this CTA's first row block of a phase (flat register tile wreg[rr * KCH + j])
template <int MODE, int KCH, int R>
__device__ __forceinline__ void mega_issue(const MegaArgs& m, const MegaGemv& g, int blk, uint4* wreg) {
  constexpr int H = R / 2;
  if (blk >= g.N / R) return;
  GemvArgs a;
  a.heads = m.heads; a.kv_heads = m.kv_heads;
  int r0, P;
  gemv3_block_rows<MODE, R>(a, blk, r0, P);
  const int K8 = g.K >> 3;
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = r0 + (rr < H ? rr : P + rr - H);
    const uint4* wp = reinterpret_cast<const uint4*>(g.W + (long long)row * g.ldw);
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      const int c = threadIdx.x + j * MEGA_THREADS;
      wreg[rr * KCH + j] = (c < K8) ? ld_stream16(wp + c) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// run one GEMV phase; wreg already holds block blockIdx.x (mega_issue was called before the preceding grid barrier).
// best_v / best_i: running arg-max of the logits rows this thread finalises (GV_LOGITS only).
template <int MODE, int KCH, int R, bool XS>
__device__ __forceinline__ void mega_gemv(const MegaArgs& m, const MegaGemv& g, int pos, uint4* wreg, uint4* xs3,
                                          float (*s_red)[MEGA_WARPS][8], float* s_ss, float& best_v, int& best_i) {
  constexpr int H = R / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K8 = g.K >> 3;
  const int n_blocks = g.N / R;
  GemvArgs a;
  a.heads = m.heads; a.kv_heads = m.kv_heads;
  uint4 xr[XS ? 1 : KCH];
  if (XS) {
    for (int c = tid; c < K8; c += MEGA_THREADS) xs3[c] = ld_cg16(reinterpret_cast<const uint4*>(g.x) + c);
    __syncthreads();
  } else {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      const int c = tid + j * MEGA_THREADS;
      xr[j] = (c < K8) ? ld_cg16(reinterpret_cast<const uint4*>(g.x) + c) : make_uint4(0u, 0u, 0u, 0u);
      const float2 p = unpack_bf16x2(xr[j].x), q = unpack_bf16x2(xr[j].y), r = unpack_bf16x2(xr[j].z),
                   s = unpack_bf16x2(xr[j].w);
      ss += p.x * p.x + p.y * p.y + q.x * q.x + q.y * q.y + r.x * r.x + r.y * r.y + s.x * s.x + s.y * s.y;
    }
    if (g.gamma != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      __syncthreads();  // s_ss may still be read by a slow warp of the previous phase
      if (lane == 0) s_ss[warp] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < MEGA_WARPS; ++w) tot += s_ss[w];
      const float rstd = rsqrtf(tot / (float)g.K + m.eps);
#pragma unroll
      for (int j = 0; j < KCH; ++j) {
        const int c = tid + j * MEGA_THREADS;
        if (c < K8) {
          const float4 g0 = *reinterpret_cast<const float4*>(g.gamma + c * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(g.gamma + c * 8 + 4);
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
  int it = 0;
  for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
    float acc[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) acc[rr] = 0.f;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      uint4 xv;
      if (XS) { const int c = tid + j * MEGA_THREADS; xv = (c < K8) ? xs3[c] : make_uint4(0u, 0u, 0u, 0u); }
      else xv = xr[j];
      const float2 p = unpack_bf16x2(xv.x), q = unpack_bf16x2(xv.y), r = unpack_bf16x2(xv.z), s = unpack_bf16x2(xv.w);
      const float xf[8] = {p.x, p.y, q.x, q.y, r.x, r.y, s.x, s.y};
#pragma unroll
      for (int rr = 0; rr < R; ++rr) acc[rr] = dot8(wreg[rr * KCH + j], xf, acc[rr]);
    }
    mega_issue<MODE, KCH, R>(m, g, blk + gridDim.x, wreg);  // next block: in flight during the reduction / epilogue
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[rr] += __shfl_xor_sync(0xffffffffu, acc[rr], o);
    }
    float(*red)[8] = s_red[it & 1];
    if (lane == 0) {
#pragma unroll
      for (int rr = 0; rr < R; ++rr) red[warp][rr] = acc[rr];
    }
    __syncthreads();
    if (tid < H) {
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < MEGA_WARPS; ++w) { v0 += red[w][tid]; v1 += red[w][tid + H]; }
      int r0, P;
      gemv3_block_rows<MODE, R>(a, blk, r0, P);
      const int ra = r0 + tid, rb = r0 + P + tid;
      if (MODE == GV_PLAIN) {
        if (g.residual) {  // L2-coherent reads: another CTA updated these rows in an earlier phase of this kernel
          v0 += __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(g.y + ra))));
          v1 += __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(g.y + rb))));
        }
        g.y[ra] = __float2bfloat16(v0);
        g.y[rb] = __float2bfloat16(v1);
      } else if (MODE == GV_LOGITS) {
        g.y32[ra] = v0;
        g.y32[rb] = v1;
        if (ra < m.vocab_valid && (v0 > best_v || (v0 == best_v && ra < best_i))) { best_v = v0; best_i = ra; }
        if (rb < m.vocab_valid && (v1 > best_v || (v1 == best_v && rb < best_i))) { best_v = v1; best_i = rb; }
      } else if (MODE == GV_SWIGLU) {
        g.y[(ra >> 6) * 32 + (ra & 31)] = __float2bfloat16(silu_f(v0) * v1);
      } else {  // GV_QKV
        const int q_w = m.heads * 128, kv_w = m.kv_heads * 128;
        if (pos >= 0 && pos < m.kv_capacity) {
          if (ra < q_w + kv_w) {
            const float t1 = __bfloat162float(__float2bfloat16(v0)), t2 = __bfloat162float(__float2bfloat16(v1));
            const float ang = (float)pos * m.inv_freq[ra & 63];
            const float cs = cosf(ang), sn = sinf(ang);
            const bf16 o1 = __float2bfloat16(t1 * cs - t2 * sn), o2 = __float2bfloat16(t2 * cs + t1 * sn);
            if (ra < q_w) { m.qb[ra] = o1; m.qb[rb] = o2; }
            else {
              bf16* kp = g.kc + (long long)pos * kv_w + (ra - q_w);
              kp[0] = o1; kp[64] = o2;
            }
          } else {
            bf16* vp = g.vc + (long long)pos * kv_w + (ra - q_w - kv_w);
            vp[0] = __float2bfloat16(v0);
            vp[P] = __float2bfloat16(v1);
          }
        }
      }
    }
  }
  __syncthreads();  // s_red / xs3 are reused by the next phase
}

constexpr int MEGA_SPLITS = 37;  // 8 kv heads x 37 key ranges = 296 = 2 CTAs x 148 SMs

__global__ void __launch_bounds__(MEGA_THREADS, 2)
decode_mega_kernel(const __grid_constant__ MegaArgs m) {
  extern __shared__ uint4 xs3[];  // attention K/V pipeline (64 KB) / activation row of the down projection
  __shared__ float s_red[2][MEGA_WARPS][8];
  __shared__ float s_ss[MEGA_WARPS];
  __shared__ DecAttnSmem<MEGA_SPLITS> s_att;
  __shared__ float s_bv[MEGA_WARPS];
  __shared__ int s_bi[MEGA_WARPS];
  unsigned epoch = 0;
  const int tid = threadIdx.x;
  const int pos = *m.position_io;
  const int H = m.hidden, kv_w = m.kv_heads * 128, qkv_w = (m.heads + 2 * m.kv_heads) * 128;
  uint4 wreg[16];
  float best_v = -INFINITY;
  int best_i = 0x7fffffff;

  MegaGemv g = {};
  g.W = m.layers[0].qkv_w; g.ldw = H; g.N = qkv_w; g.K = H; g.x = m.x; g.gamma = m.layers[0].ln1_g;
  mega_issue<GV_QKV, 2, 8>(m, g, blockIdx.x, wreg);
  if (blockIdx.x == 0) {  // embedding row of the previous token -> residual stream; clear the attention tickets
    long long id = *m.token_io;
    if (id < 0) id = 0;
    if (id >= m.embed_rows) id = m.embed_rows - 1;
    const uint4* sp = reinterpret_cast<const uint4*>(m.embed + id * m.ld_embed);
    uint4* dp = reinterpret_cast<uint4*>(m.x);
    for (int c = tid; c < (H >> 3); c += MEGA_THREADS) dp[c] = sp[c];
    for (int i = tid; i < m.kv_heads; i += MEGA_THREADS) m.tickets[i] = 0u;
  }
  mega_grid_sync(m.barrier, epoch, m.dbg);

  for (int l = 0; l < m.n_layers; ++l) {
    const MegaLayer& L = m.layers[l];
    bf16* kc = m.kv_cache + (size_t)l * 2 * m.kv_capacity * kv_w;
    bf16* vc = kc + (size_t)m.kv_capacity * kv_w;
    // ---- RMSNorm(ln1) -> QKV -> RoPE -> cache append
    g = MegaGemv{};
    g.W = L.qkv_w; g.ldw = H; g.N = qkv_w; g.K = H; g.x = m.x; g.gamma = L.ln1_g; g.kc = kc; g.vc = vc;
    mega_gemv<GV_QKV, 2, 8, false>(m, g, pos, wreg, xs3, s_red, s_ss, best_v, best_i);
    mega_grid_sync(m.barrier, epoch, m.dbg);
    // ---- attention over rows 0 .. pos
    if ((int)blockIdx.x < m.kv_heads * MEGA_SPLITS)  // uniform per CTA: exactly one work item per CTA of the grid
      decode_attn_item<MEGA_SPLITS>(m.qb, kc, vc, min(pos + 1, (int)m.kv_capacity), kv_w, (int)blockIdx.x / MEGA_SPLITS,
                                    (int)blockIdx.x % MEGA_SPLITS, m.scale_log2, m.part, m.tickets, m.att,
                                    reinterpret_cast<uint8_t*>(xs3), s_att);
    __syncthreads();
    g = MegaGemv{};
    g.W = L.o_w; g.ldw = H; g.N = H; g.K = H; g.x = m.att; g.y = m.x; g.residual = 1;
    mega_issue<GV_PLAIN, 2, 8>(m, g, blockIdx.x, wreg);
    mega_grid_sync(m.barrier, epoch, m.dbg);
    // ---- x += o_proj(att)
    mega_gemv<GV_PLAIN, 2, 8, false>(m, g, pos, wreg, xs3, s_red, s_ss, best_v, best_i);
    g = MegaGemv{};
    g.W = L.gateup_w; g.ldw = H; g.N = 2 * m.ffn; g.K = H; g.x = m.x; g.gamma = L.ln2_g; g.y = m.act;
    mega_issue<GV_SWIGLU, 2, 8>(m, g, blockIdx.x, wreg);
    mega_grid_sync(m.barrier, epoch, m.dbg);
    // ---- act = silu(gate) * up of RMSNorm(ln2)(x)
    mega_gemv<GV_SWIGLU, 2, 8, false>(m, g, pos, wreg, xs3, s_red, s_ss, best_v, best_i);
    g = MegaGemv{};
    g.W = L.down_w; g.ldw = m.ffn; g.N = H; g.K = m.ffn; g.x = m.act; g.y = m.x; g.residual = 1;
    mega_issue<GV_PLAIN, 7, 2>(m, g, blockIdx.x, wreg);
    mega_grid_sync(m.barrier, epoch, m.dbg);
    // ---- x += down(act)
    mega_gemv<GV_PLAIN, 7, 2, true>(m, g, pos, wreg, xs3, s_red, s_ss, best_v, best_i);
    g = MegaGemv{};
    if (l + 1 < m.n_layers) {
      g.W = m.layers[l + 1].qkv_w; g.ldw = H; g.N = qkv_w; g.K = H;
      mega_issue<GV_QKV, 2, 8>(m, g, blockIdx.x, wreg);
    } else {
      g.W = m.lm_head; g.ldw = H; g.N = m.vocab; g.K = H;
      mega_issue<GV_LOGITS, 2, 8>(m, g, blockIdx.x, wreg);
    }
    mega_grid_sync(m.barrier, epoch, m.dbg);
  }
  // ---- final RMSNorm -> lm_head -> arg-max
  g = MegaGemv{};
  g.W = m.lm_head; g.ldw = H; g.N = m.vocab; g.K = H; g.x = m.x; g.gamma = m.final_g; g.y32 = m.logits;
  mega_gemv<GV_LOGITS, 2, 8, false>(m, g, pos, wreg, xs3, s_red, s_ss, best_v, best_i);
  {  // CTA candidate (threads 0..3 hold the rows this CTA finalised)
    if (tid < 32) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
      }
      if (tid == 0) { m.cand_v[blockIdx.x] = best_v; m.cand_i[blockIdx.x] = best_i; }
    }
  }
  mega_grid_sync(m.barrier, epoch, m.dbg);
  if (blockIdx.x == 0) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < (int)gridDim.x; i += MEGA_THREADS) {
      const float v = __ldcg(m.cand_v + i);
      const int ci = __ldcg(m.cand_i + i);
      if (v > bv || (v == bv && ci < bi)) { bv = v; bi = ci; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_bv[tid >> 5] = bv; s_bi[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < MEGA_WARPS; ++w)
        if (s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi)) { bv = s_bv[w]; bi = s_bi[w]; }
      *m.token_io = (bi == 0x7fffffff) ? 0 : (long long)bi;
      *m.position_io = pos + 1;
    }
  }
}

}  // namespace lp

using namespace lp;

static inline size_t bf(size_t n) { return n * 2; }

extern "C" int lp_decode_set_prefetch(int o_proj_mb, int gateup_mb) {
  g_dec_prefetch_mb[0] = o_proj_mb < 0 ? 0 : o_proj_mb;
  g_dec_prefetch_mb[1] = gateup_mb < 0 ? 0 : gateup_mb;
  return 0;
}

extern "C" int lp_decode_set_mode(int mode) {
  const int old = g_decode_mode;
  g_decode_mode = mode & 23;
  return old;
}

extern "C" size_t lp_decode_workspace_bytes(const lp_dec_dims* d) {
  const size_t qkv_w = (size_t)(d->heads + 2 * d->kv_heads) * d->head_dim;
  return 256 * 9 + bf(2 * (size_t)d->hidden) * 2 + bf(qkv_w) + bf((size_t)d->ffn) + (size_t)d->head_dim * 4 + 64 * 4 +
         (size_t)d->heads * DEC_SPLITS * 130 * 4 + (size_t)d->heads * 4 + 256 + 3 * 256 + 2 * 4 * 2 * 256;
}

#define LP_TRY(call) do { rc = (call); if (rc) return rc; } while (0)

static int decode_attn_configure() {  // opt in to > 48 KB of dynamic shared memory once
  static bool done = false;
  if (!done) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DA_SMEM_BYTES));
    LP_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DA_SMEM_BYTES));
    done = true;
  }
  return 0;
}

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
      LP_TRY(decode_attn_configure());
      decode_attn_kernel<<<dim3(d->kv_heads, DEC_SPLITS), DEC_WARPS * 32, DA_SMEM_BYTES, stream>>>(
          qkv, kc, vc, nullptr, position, (int)kv_capacity, kv_w, scale_log2, part, tickets, att, nullptr, 0LL, nullptr, 0LL);
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
      LP_TRY(decode_attn_configure());
      LP_TRY(launch_pdl(decode_attn_kernel, dim3(d->kv_heads, DEC_SPLITS), dim3(DEC_WARPS * 32), DA_SMEM_BYTES, stream,
                        (const bf16*)qb, (const bf16*)kc, (const bf16*)vc, pos_dev, pos_host, (int)kv_capacity, kv_w,
                        scale_log2, part, tickets, att, (const uint8_t*)L->o_w,
                        min((long long)g_dec_prefetch_mb[0] << 20, (long long)H * H * 2), (const uint8_t*)L->gateup_w,
                        min((long long)g_dec_prefetch_mb[1] << 20, (long long)2 * d->ffn * H * 2)));
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

static unsigned long long* g_mega_dbg = nullptr;
extern "C" int lp_decode_debug_stamps(void* device_u64_buffer) {  // >= 2 * (5 * layers + 3) + 2 entries, or NULL
  g_mega_dbg = reinterpret_cast<unsigned long long*>(device_u64_buffer);
  return 0;
}

// one persistent cooperative kernel per token (mode bit 4)
static int decode_step_mega(const lp_dec_dims* d, const lp_dec_weights* w, const void* embed_table, long long ld_embed,
                            long long* token_io, int* position_io, void* kv_cache, long long kv_capacity, void* logits,
                            int vocab_valid, void* workspace, cudaStream_t stream) {
  const int H = d->hidden, hd = d->head_dim;
  const int q_w = d->heads * hd;
  LP_REQUIRE(d->layers <= MEGA_MAX_LAYERS && H % 8 == 0 && H <= 4096 && d->ffn <= 14336 && d->ffn % 32 == 0 &&
                 d->vocab % 8 == 0 && d->heads == 4 * d->kv_heads,
             "lp_decode_greedy_step: the single-kernel step supports hidden <= 4096, ffn <= 14336, <= 32 layers");
  const int grid = sm_count() * 2;
  LP_REQUIRE(grid >= d->kv_heads * MEGA_SPLITS, "lp_decode_greedy_step: single-kernel step needs %d resident CTAs",
             d->kv_heads * MEGA_SPLITS);
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  auto take = [&](size_t n) { void* r = p; p += (n + 255) & ~size_t(255); return r; };
  MegaArgs m = {};
  m.x = (bf16*)take(bf(H));
  m.qb = (bf16*)take(bf(q_w));
  m.att = (bf16*)take(bf(H));
  m.act = (bf16*)take(bf(d->ffn));
  m.part = (float*)take((size_t)d->heads * DEC_SPLITS * 130 * 4);
  m.tickets = (unsigned*)take((size_t)d->heads * 4);
  m.barrier = (unsigned*)take(4);
  m.cand_v = (float*)take((size_t)grid * 4);
  m.cand_i = (int*)take((size_t)grid * 4);
  for (int l = 0; l < d->layers; ++l) {
    const lp_dec_layer_weights* L = &w->layers[l];
    m.layers[l] = MegaLayer{(const bf16*)L->qkv_w, (const bf16*)L->o_w, (const bf16*)L->gateup_w, (const bf16*)L->down_w,
                            L->ln1_g, L->ln2_g};
  }
  m.n_layers = d->layers; m.hidden = H; m.ffn = d->ffn; m.heads = d->heads; m.kv_heads = d->kv_heads;
  m.vocab = d->vocab; m.vocab_valid = vocab_valid;
  m.eps = d->eps; m.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  m.lm_head = (const bf16*)w->lm_head; m.final_g = w->final_g; m.inv_freq = w->inv_freq;
  m.embed = (const bf16*)embed_table; m.ld_embed = ld_embed; m.embed_rows = d->vocab;
  m.kv_cache = (bf16*)kv_cache; m.kv_capacity = kv_capacity;
  m.token_io = token_io; m.position_io = position_io; m.logits = (float*)logits;
  m.dbg = g_mega_dbg;
  LP_CHECK_CUDA(cudaMemsetAsync(m.barrier, 0, 4, stream));
  const size_t smem = DA_SMEM_BYTES;  // >= ffn * 2 (activation row of the down projection)
  LP_REQUIRE((size_t)d->ffn * 2 <= smem, "lp_decode_greedy_step: ffn too wide for the staging buffer");
  int rc;
  LP_TRY(decode_attn_configure());
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(MEGA_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident, or the launch fails (never a deadlock)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(15, d->layers, H, 0, stream);
  LP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_mega_kernel, m));
  note_launch();
  return 0;
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
  if (g_decode_mode & 16)
    return decode_step_mega(d, w, embed_table, ld_embed, token_io, position_io, kv_cache, kv_capacity, logits,
                            vocab_valid, workspace, stream);
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
