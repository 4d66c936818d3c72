// leopard_b200 — HBM-bound glue kernels of the prefill path (128-bit coalesced loads/stores, fp32 math).
//
//   lp_layernorm     SigLIP LayerNorm (eps 1e-6, affine); optional pixel-shuffle store fuses Leopard's 4:1 visual
//                    token compressor (evaluations/models/llava_multiimg_siglip_anyres.py:165-176 ==
//                    megatron_patch/model/llava/vlm_model.py:456-466) into ln_post's write addressing
//   lp_rmsnorm       Llama/Mistral RMSNorm (Megatron-LM-240603/megatron/legacy/model/rms_norm.py:26-31)
//   lp_rope_table / lp_rope_inplace   rotate-half RoPE on the q/k column slices of the fused QKV buffer
//                    (megatron/core/models/common/embeddings/rotary_pos_embedding.py:197-239)
//   lp_embed_merge   token embedding + image-feature merge, HF-expand and Megatron-scatter conventions
//                    (HF _merge_input_ids_with_image_features; vlm_model.py:516-533)
//   lp_im2col_patch14  14x14/stride-14 patch extraction feeding the patch-embed GEMM
//                    (megatron_patch/model/idefics2/idefics_vision_tower.py:57-64,118-122)
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

constexpr int NORM_THREADS = 128;   // 4 warps = 4 rows per CTA
constexpr int NORM_MAX_D = 4096;

// One WARP per row (no shared memory, shuffle reductions only), the row lives in registers between the statistics
// pass and the normalise/store pass: 2 x rows x D x 2 bytes of HBM traffic, 16-byte accesses.
// RMS = false: LayerNorm (two-pass mean / variance in fp32), RMS = true: RMSNorm.  CH = 16-byte chunks per lane.
template <bool RMS, int CH>
__global__ void __launch_bounds__(NORM_THREADS)
norm_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
            bf16* __restrict__ out, long long ldo, int M, int D, float eps, int shuffle_side) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (NORM_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const bf16* xr = x + (long long)row * ldx;
  float v[CH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + c * 8);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
      v[i][4] = cc.x; v[i][5] = cc.y; v[i][6] = d.x; v[i][7] = d.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += RMS ? v[i][j] * v[i][j] : v[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(s / (float)D + eps);
  } else {
    mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + i * 32;
      if (c < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          q += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd = rsqrtf(q / (float)D + eps);
  }
  // destination row / column offset (pixel-shuffle store: 2x2 neighbourhoods are concatenated along channels)
  long long orow = row;
  int ocol = 0;
  if (shuffle_side > 0) {
    const int per = shuffle_side * shuffle_side;
    const int b = row / per, t = row % per;
    const int h = t / shuffle_side, w = t % shuffle_side;
    const int half = shuffle_side >> 1;
    orow = (long long)b * (half * half) + (h >> 1) * half + (w >> 1);
    ocol = (((h & 1) << 1) | (w & 1)) * D;
  }
  bf16* orp = out + orow * ldo + ocol;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + c * 8 + 4);
      float y[8];
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (v[i][j] - mean) * rstd * gg[j];
      if (!RMS && beta != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(beta + c * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(beta + c * 8 + 4);
        y[0] += b0.x; y[1] += b0.y; y[2] += b0.z; y[3] += b0.w;
        y[4] += b1.x; y[5] += b1.y; y[6] += b1.z; y[7] += b1.w;
      }
      uint4 u;
      u.x = pack_bf16x2(y[0], y[1]); u.y = pack_bf16x2(y[2], y[3]);
      u.z = pack_bf16x2(y[4], y[5]); u.w = pack_bf16x2(y[6], y[7]);
      *reinterpret_cast<uint4*>(orp + c * 8) = u;
    }
  }
}

// RMSNorm for wide rows (D > 1280): the row is NOT kept in registers — pass 1 accumulates sum(x^2), pass 2 re-reads
// the row (it was just fetched, so this hits L1/L2, HBM traffic stays 2 x rows x D x 2 B) and writes x * rstd * gamma.
// Few registers -> 8 rows per CTA at full occupancy, which the register-resident variant (16 chunks/lane) cannot reach.
__global__ void __launch_bounds__(256)
rmsnorm_reread_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                      bf16* __restrict__ out, long long ldo, int M, int D, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  float s = 0.f;
#pragma unroll 4
  for (int c = lane; c < nchunk; c += 32) {
    const uint4 u = xr[c];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    s += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float rstd = rsqrtf(s / (float)D + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + (long long)row * ldo);
#pragma unroll 4
  for (int c = lane; c < nchunk; c += 32) {
    const uint4 u = xr[c];
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(gamma + c * 8 + 4);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    uint4 y;
    y.x = pack_bf16x2(a.x * rstd * g0.x, a.y * rstd * g0.y);
    y.y = pack_bf16x2(b.x * rstd * g0.z, b.y * rstd * g0.w);
    y.z = pack_bf16x2(cc.x * rstd * g1.x, cc.y * rstd * g1.y);
    y.w = pack_bf16x2(d.x * rstd * g1.z, d.y * rstd * g1.w);
    orow[c] = y;
  }
}

// RMSNorm for wide rows, register-resident and PACKED: a lane keeps its 16-byte chunks as bf16 (4 registers each, 64 for
// D = 4096) instead of 8 floats, so the whole row is requested in one batch (16 independent loads per lane in flight,
// 8 KB per warp), read from memory exactly once, and occupancy stays at 16 warps per SM.  The statistics / normalise
// arithmetic and rounding points are identical to norm_kernel<true, *> and rmsnorm_reread_kernel.
template <int CH>
__global__ void __launch_bounds__(256, 3)  // <= 85 registers: keeps the row packed instead of 128 unpacked floats
rmsnorm_packed_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                      bf16* __restrict__ out, long long ldo, int M, int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nchunk = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  uint4 u[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + i * 32;
    u[i] = (c < nchunk) ? xr[c] : make_uint4(0u, 0u, 0u, 0u);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const float2 a = unpack_bf16x2(u[i].x), b = unpack_bf16x2(u[i].y), cc = unpack_bf16x2(u[i].z), d = unpack_bf16x2(u[i].w);
    s += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float rstd = rsqrtf(s / (float)D + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + (long long)row * ldo);
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    // opaque to the optimiser: the floats unpacked for the statistics pass must not be kept alive (128 registers);
    // they are re-derived from the packed registers here (2 ALU ops per pair)
    asm volatile("" : "+r"(u[i].x), "+r"(u[i].y), "+r"(u[i].z), "+r"(u[i].w));
    const int c = lane + i * 32;
    if (c < nchunk) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + c * 8 + 4);
      const float2 a = unpack_bf16x2(u[i].x), b = unpack_bf16x2(u[i].y), cc = unpack_bf16x2(u[i].z), d = unpack_bf16x2(u[i].w);
      uint4 y;
      y.x = pack_bf16x2(a.x * rstd * g0.x, a.y * rstd * g0.y);
      y.y = pack_bf16x2(b.x * rstd * g0.z, b.y * rstd * g0.w);
      y.z = pack_bf16x2(cc.x * rstd * g1.x, cc.y * rstd * g1.y);
      y.w = pack_bf16x2(d.x * rstd * g1.z, d.y * rstd * g1.w);
      orow[c] = y;
    }
  }
}

// (sum, sum of squares) of each row: the `stats_in` producer for rows that no GEMM epilogue wrote (lp_gemm_bf16_ex).
// One warp per row, 16-byte loads, fixed shuffle-tree order (deterministic).
__global__ void __launch_bounds__(256)
row_stats_kernel(const bf16* __restrict__ x, long long ldx, int M, int D, float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < (D >> 3); c += 32) {
    const uint4 u = xr[c];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    s1 += ((a.x + a.y) + (b.x + b.y)) + ((cc.x + cc.y) + (d.x + d.y));
    s2 += ((a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y)) + ((cc.x * cc.x + cc.y * cc.y) + (d.x * d.x + d.y * d.y));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) stats[row] = make_float2(s1, s2);
}

static int g_norm_variant = 1;  // 1: packed register-resident RMSNorm for wide rows, 0: re-read variant

template <bool RMS>
static void launch_norm(const bf16* x, long long ldx, const float* g, const float* b, bf16* out, long long ldo, int M,
                        int D, float eps, int shuffle_side, cudaStream_t stream) {
  const int grid = (M + NORM_THREADS / 32 - 1) / (NORM_THREADS / 32);
  if (D <= 32 * 8 * 5)
    launch_ex(norm_kernel<RMS, 5>, dim3(grid), dim3(NORM_THREADS), 0, stream, 1, prefill_pdl(), x, ldx, g, b, out, ldo, M, D,
              eps, shuffle_side);
  else if (RMS && shuffle_side == 0 && g_norm_variant == 1)
    launch_ex(rmsnorm_packed_kernel<16>, dim3((M + 7) / 8), dim3(256), 0, stream, 1, prefill_pdl(), x, ldx, g, out, ldo, M, D,
              eps);
  else if (RMS && shuffle_side == 0 && x != out)
    rmsnorm_reread_kernel<<<(M + 7) / 8, 256, 0, stream>>>(x, ldx, g, out, ldo, M, D, eps);
  else
    norm_kernel<RMS, 16><<<grid, NORM_THREADS, 0, stream>>>(x, ldx, g, b, out, ldo, M, D, eps, shuffle_side);
}

// cos/sin tables [S, half] from integer positions and fp32 inverse frequencies
__global__ void rope_table_kernel(const int* __restrict__ pos, int pos_base, const float* __restrict__ inv_freq,
                                  float* __restrict__ cs, float* __restrict__ sn, int S, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * half) return;
  const int s = i / half, j = i % half;
  const float a = (float)(pos ? pos[s] : pos_base + s) * inv_freq[j];
  cs[i] = cosf(a);
  sn[i] = sinf(a);
}

// rotate-half RoPE in place on `n_heads` consecutive heads of width 2*half starting at column col0.
// one thread per (row, head, 8-wide chunk pair): x1 = cols [c, c+8), x2 = cols [c + half, c + half + 8)
__global__ void rope_kernel(bf16* __restrict__ x, long long ld, int S, int col0, int n_heads, int half,
                            const float* __restrict__ cs, const float* __restrict__ sn) {
  pdl_launch_dependents();
  pdl_wait();
  const int cpr = half >> 3;  // chunk pairs per head
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S * n_heads * cpr;
  if (i >= total) return;
  const int c = (int)(i % cpr);
  const int h = (int)((i / cpr) % n_heads);
  const int s = (int)(i / ((long long)cpr * n_heads));
  bf16* p1 = x + (long long)s * ld + col0 + h * 2 * half + c * 8;
  bf16* p2 = p1 + half;
  const uint4 u1 = *reinterpret_cast<const uint4*>(p1), u2 = *reinterpret_cast<const uint4*>(p2);
  const float* cp = cs + (long long)s * half + c * 8;
  const float* sp = sn + (long long)s * half + c * 8;
  const uint32_t a1[4] = {u1.x, u1.y, u1.z, u1.w}, a2[4] = {u2.x, u2.y, u2.z, u2.w};
  uint32_t o1[4], o2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f1 = unpack_bf16x2(a1[j]), f2 = unpack_bf16x2(a2[j]);
    const float c0 = cp[2 * j], c1 = cp[2 * j + 1], s0 = sp[2 * j], s1 = sp[2 * j + 1];
    // t*cos + rotate_half(t)*sin, rotate_half(x) = cat(-x2, x1)
    o1[j] = pack_bf16x2(f1.x * c0 - f2.x * s0, f1.y * c1 - f2.y * s1);
    o2[j] = pack_bf16x2(f2.x * c0 + f1.x * s0, f2.y * c1 + f1.y * s1);
  }
  *reinterpret_cast<uint4*>(p1) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
  *reinterpret_cast<uint4*>(p2) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
}

// RoPE on the q and k heads of the fused QKV buffer AND the KV-cache append of the same rows in one pass (prefill with
// a cache): k heads are written rotated both in place (attention reads them there) and to kc[row], v heads are copied
// to vc[row].  Replaces rope_kernel + two strided cudaMemcpy2DAsync per layer (64 copy launches per prefill).
// Layout of x: [n_q heads | n_kv k heads | n_kv v heads] x 2*half columns; cache rows have n_kv * 2*half columns.
__global__ void rope_kv_kernel(bf16* __restrict__ x, long long ld, int S, int n_q, int n_kv, int half,
                               const float* __restrict__ cs, const float* __restrict__ sn, bf16* __restrict__ kc,
                               bf16* __restrict__ vc, long long ld_cache) {
  pdl_launch_dependents();
  pdl_wait();
  const int cpr = half >> 3;  // chunk pairs per head
  const int n_heads = n_q + 2 * n_kv;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)S * n_heads * cpr;
  if (i >= total) return;
  const int c = (int)(i % cpr);
  const int h = (int)((i / cpr) % n_heads);
  const int s = (int)(i / ((long long)cpr * n_heads));
  bf16* p1 = x + (long long)s * ld + h * 2 * half + c * 8;
  bf16* p2 = p1 + half;
  if (cs == nullptr && h < n_q) return;  // append-only mode (RoPE already applied by the QKV GEMM's epilogue): q heads untouched
  const uint4 u1 = *reinterpret_cast<const uint4*>(p1), u2 = *reinterpret_cast<const uint4*>(p2);
  if (cs == nullptr) {  // k head, already rotated: copy into the cache
    if (h < n_q + n_kv) {
      bf16* d = kc + (long long)s * ld_cache + (h - n_q) * 2 * half + c * 8;
      *reinterpret_cast<uint4*>(d) = u1;
      *reinterpret_cast<uint4*>(d + half) = u2;
      return;
    }
  }
  if (h >= n_q + n_kv) {  // v head: plain copy into the cache
    bf16* d = vc + (long long)s * ld_cache + (h - n_q - n_kv) * 2 * half + c * 8;
    *reinterpret_cast<uint4*>(d) = u1;
    *reinterpret_cast<uint4*>(d + half) = u2;
    return;
  }
  const float* cp = cs + (long long)s * half + c * 8;
  const float* sp = sn + (long long)s * half + c * 8;
  const uint32_t a1[4] = {u1.x, u1.y, u1.z, u1.w}, a2[4] = {u2.x, u2.y, u2.z, u2.w};
  uint32_t o1[4], o2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f1 = unpack_bf16x2(a1[j]), f2 = unpack_bf16x2(a2[j]);
    const float c0 = cp[2 * j], c1 = cp[2 * j + 1], s0 = sp[2 * j], s1 = sp[2 * j + 1];
    o1[j] = pack_bf16x2(f1.x * c0 - f2.x * s0, f1.y * c1 - f2.y * s1);   // same arithmetic as rope_kernel
    o2[j] = pack_bf16x2(f2.x * c0 + f1.x * s0, f2.y * c1 + f1.y * s1);
  }
  const uint4 r1 = make_uint4(o1[0], o1[1], o1[2], o1[3]), r2 = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  *reinterpret_cast<uint4*>(p1) = r1;
  *reinterpret_cast<uint4*>(p2) = r2;
  if (h >= n_q) {  // k head: also into the cache (K is cached after RoPE)
    bf16* d = kc + (long long)s * ld_cache + (h - n_q) * 2 * half + c * 8;
    *reinterpret_cast<uint4*>(d) = r1;
    *reinterpret_cast<uint4*>(d + half) = r2;
  }
}

// ---- embedding + merge -------------------------------------------------------------------------------------------
// pass 1 (single CTA): per input token t, dst[t] = first output row, src[t] = first feature row (or -1 for text)
__global__ void __launch_bounds__(1024)
merge_plan_kernel(const long long* __restrict__ ids, int L, long long image_id, int rows_per_slot,
                  int* __restrict__ dst, int* __restrict__ src, int* __restrict__ totals) {
  __shared__ int s_cnt[1024];
  const int per = (L + 1023) / 1024;
  const int t0 = threadIdx.x * per, t1 = min(L, t0 + per);
  int cnt = 0;
  for (int t = t0; t < t1; ++t) cnt += (ids[t] == image_id);
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  // inclusive scan (Hillis-Steele)
  for (int o = 1; o < 1024; o <<= 1) {
    int v = (threadIdx.x >= o) ? s_cnt[threadIdx.x - o] : 0;
    __syncthreads();
    s_cnt[threadIdx.x] += v;
    __syncthreads();
  }
  int before = s_cnt[threadIdx.x] - cnt;  // image slots before t0
  for (int t = t0; t < t1; ++t) {
    const bool img = ids[t] == image_id;
    dst[t] = t + before * (rows_per_slot - 1);
    src[t] = img ? before * rows_per_slot : -1;
    before += img;
  }
  if (threadIdx.x == 1023) {
    totals[0] = s_cnt[1023];                              // number of image slots
    totals[1] = L + s_cnt[1023] * (rows_per_slot - 1);    // merged length S
  }
}

// pass 2: CTA (t, r) copies row r of input token t (text tokens have one row; an image slot has rows_per_slot rows).
// One CTA per OUTPUT row: a slot's 169 rows used to be copied serially by one CTA (1.4 MB each, 185 us per merge).
__global__ void __launch_bounds__(128)
merge_copy_kernel(const long long* __restrict__ ids, const int* __restrict__ dst, const int* __restrict__ src,
                  const bf16* __restrict__ table, long long ldt, long long vocab, const bf16* __restrict__ feats,
                  long long ldf, int n_feat_rows, bf16* __restrict__ out, long long ldo, int D, int rows_per_slot) {
  const int t = blockIdx.x, r = blockIdx.y;
  const int nchunk = D >> 3;
  const int s = src[t];
  if (s < 0) {
    if (r != 0) return;
    long long id = ids[t];
    if (id < 0) id = 0;            // Megatron clamps ignore/pad ids to row 0 (vlm_model.py:226)
    if (id >= vocab) id = vocab - 1;
    const uint4* sp = reinterpret_cast<const uint4*>(table + id * ldt);
    uint4* dp = reinterpret_cast<uint4*>(out + (long long)dst[t] * ldo);
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) dp[c] = sp[c];
  } else {
    if (s + r >= n_feat_rows) return;  // count mismatch is reported by the host wrapper
    const uint4* sp = reinterpret_cast<const uint4*>(feats + (long long)(s + r) * ldf);
    uint4* dp = reinterpret_cast<uint4*>(out + (long long)(dst[t] + r) * ldo);
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) dp[c] = sp[c];
  }
}

// ---- im2col for the 14x14 / stride 14 patch embedding ------------------------------------------------------------
// out[(n*G + py)*G + px, c*196 + ky*14 + kx] = pix[n, c, py*14 + ky, px*14 + kx]; columns [588, ldo) are zeroed.
template <typename T>
__global__ void im2col_patch14_kernel(const T* __restrict__ pix, bf16* __restrict__ out, long long ldo, int N, int C,
                                      int H, int W) {
  const int G_h = H / 14, G_w = W / 14;
  const int segs = C * 14 + 1;  // +1: the zero pad segment
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * G_h * G_w * segs;
  if (i >= total) return;
  const int seg = (int)(i % segs);
  const long long patch = i / segs;
  bf16* orow = out + patch * ldo;
  if (seg == C * 14) {
    for (int k = C * 196; k < ldo; ++k) orow[k] = __float2bfloat16(0.f);
    return;
  }
  const int c = seg / 14, ky = seg % 14;
  const int px = (int)(patch % G_w), py = (int)((patch / G_w) % G_h), n = (int)(patch / ((long long)G_w * G_h));
  const T* src = pix + (((long long)n * C + c) * H + (py * 14 + ky)) * W + px * 14;
  bf16* dst = orow + c * 196 + ky * 14;
#pragma unroll
  for (int kx = 0; kx < 14; kx += 2) {
    const float a = (float)src[kx], b = (float)src[kx + 1];
    *reinterpret_cast<uint32_t*>(dst + kx) = pack_bf16x2(a, b);
  }
}

// ---- standalone pixel shuffle (the fused path is lp_layernorm's shuffle store) ----------------------------------
// out[b, oh*half + ow, (dy*2+dx)*D + e] = x[b, (2oh+dy)*side + 2ow+dx, e]; one CTA per input row
__global__ void __launch_bounds__(128)
pixel_shuffle_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ out, long long ldo, int side, int D) {
  const int row = blockIdx.x;
  const int per = side * side, half = side >> 1;
  const int b = row / per, t = row % per, h = t / side, w = t % side;
  const long long orow = (long long)b * (half * half) + (h >> 1) * half + (w >> 1);
  const uint4* sp = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  uint4* dp = reinterpret_cast<uint4*>(out + orow * ldo + (((h & 1) << 1) | (w & 1)) * D);
  for (int c = threadIdx.x; c < (D >> 3); c += blockDim.x) dp[c] = sp[c];
}

// ---- per-token cross entropy over the vocabulary (Megatron forward with labels) -----------------------------------
// loss[r] = logsumexp(logits[r, :V]) - logits[r, label[r]]   (fp32 math; label < 0 -> 0); one CTA per row
template <typename T>
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                     float* __restrict__ loss, int V) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const T* row = logits + (long long)r * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += 256) mx = fmaxf(mx, (float)row[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += 256) s += __expf((float)row[i] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    const long long lab = labels[r];
    loss[r] = (lab >= 0 && lab < V) ? (logf(t) + mx - (float)row[lab]) : 0.f;
  }
}

}  // namespace lp

using namespace lp;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int lp_layernorm(const void* x, long long ldx, const float* gamma, const float* beta, void* out,
                            long long ldo, int M, int D, float eps, int shuffle_side, void* stream) {
  LP_REQUIRE(M > 0 && D > 0 && D % 8 == 0 && D <= NORM_MAX_D, "lp_layernorm: bad M=%d D=%d", M, D);
  LP_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && aligned16(x) && aligned16(out) && aligned16(gamma) &&
                 (beta == nullptr || aligned16(beta)),
             "lp_layernorm: 16-byte alignment required");
  LP_REQUIRE(shuffle_side == 0 || (shuffle_side % 2 == 0 && M % (shuffle_side * shuffle_side) == 0),
             "lp_layernorm: pixel-shuffle store needs an even grid side and M %% side^2 == 0");
  ProfScope prof(3, M, D, 0, reinterpret_cast<cudaStream_t>(stream));
  note_launch();
  launch_norm<false>(reinterpret_cast<const bf16*>(x), ldx, gamma, beta, reinterpret_cast<bf16*>(out), ldo, M, D, eps,
                     shuffle_side, reinterpret_cast<cudaStream_t>(stream));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_rmsnorm(const void* x, long long ldx, const float* gamma, void* out, long long ldo, int M, int D,
                          float eps, void* stream) {
  LP_REQUIRE(M > 0 && D > 0 && D % 8 == 0 && D <= NORM_MAX_D, "lp_rmsnorm: bad M=%d D=%d", M, D);
  LP_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && aligned16(x) && aligned16(out) && aligned16(gamma),
             "lp_rmsnorm: 16-byte alignment required");
  ProfScope prof(3, M, D, 1, reinterpret_cast<cudaStream_t>(stream));
  note_launch();
  launch_norm<true>(reinterpret_cast<const bf16*>(x), ldx, gamma, nullptr, reinterpret_cast<bf16*>(out), ldo, M, D, eps, 0,
                    reinterpret_cast<cudaStream_t>(stream));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_row_stats(const void* x, long long ldx, int rows, int width, float* stats, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(x && stats && rows > 0 && width > 0 && width % 8 == 0 && ldx % 8 == 0 && aligned16(x) &&
                 (reinterpret_cast<uintptr_t>(stats) & 7) == 0,
             "lp_row_stats: bad arguments rows=%d width=%d", rows, width);
  ProfScope prof(3, rows, width, 2, stream);
  note_launch();
  LP_CHECK_CUDA(launch_ex(row_stats_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, 1, prefill_pdl(),
                          reinterpret_cast<const bf16*>(x), ldx, rows, width, reinterpret_cast<float2*>(stats)));
  return 0;
}

extern "C" int lp_norm_set_variant(int variant) {
  const int old = g_norm_variant;
  g_norm_variant = variant;
  return old;
}

extern "C" int lp_rope_table(const int* pos, const float* inv_freq, float* cos_out, float* sin_out, int S, int half,
                             void* stream) {
  LP_REQUIRE(S > 0 && half > 0 && pos != nullptr, "lp_rope_table: bad S=%d half=%d", S, half);
  const int total = S * half;
  ProfScope prof(4, S, half, 0, reinterpret_cast<cudaStream_t>(stream));
  note_launch();
  rope_table_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pos, 0, inv_freq, cos_out,
                                                                                           sin_out, S, half);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace lp {
// positions pos_base .. pos_base + S - 1 without a device-side position array (decode step)
int rope_table_base(int pos_base, const float* inv_freq, float* cs, float* sn, int S, int half, cudaStream_t stream) {
  note_launch();
  rope_table_kernel<<<(S * half + 255) / 256, 256, 0, stream>>>(nullptr, pos_base, inv_freq, cs, sn, S, half);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace lp

extern "C" int lp_rope_inplace(void* x, long long ld, int S, int col0, int n_heads, int head_dim, const float* cos_t,
                               const float* sin_t, void* stream) {
  LP_REQUIRE(S > 0 && n_heads > 0 && head_dim % 16 == 0, "lp_rope_inplace: bad S=%d heads=%d dim=%d", S, n_heads,
             head_dim);
  LP_REQUIRE(ld % 8 == 0 && col0 % 8 == 0 && aligned16(x), "lp_rope_inplace: 16-byte alignment required");
  const int half = head_dim / 2;
  const long long total = (long long)S * n_heads * (half / 8);
  ProfScope prof(4, S, n_heads, 1, reinterpret_cast<cudaStream_t>(stream));
  note_launch();
  launch_ex(rope_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
            prefill_pdl(), reinterpret_cast<bf16*>(x), ld, S, col0, n_heads, half, cos_t, sin_t);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace lp {
// RoPE in place on the q|k heads of x [S, ld] + append K (rotated) and V rows 0 .. S-1 to the cache (see rope_kv_kernel);
// cos_t == nullptr: append only (x already holds rotated q / k)
int rope_kv_append(void* x, long long ld, int S, int n_q, int n_kv, int head_dim, const float* cos_t, const float* sin_t,
                   void* kc, void* vc, long long ld_cache, cudaStream_t stream) {
  LP_REQUIRE(S > 0 && head_dim % 16 == 0 && ld % 8 == 0 && ld_cache % 8 == 0 && aligned16(x) && aligned16(kc) && aligned16(vc),
             "rope_kv_append: bad S=%d head_dim=%d or alignment", S, head_dim);
  const int half = head_dim / 2;
  const long long total = (long long)S * (n_q + 2 * n_kv) * (half / 8);
  ProfScope prof(4, S, n_q + n_kv, 2, stream);
  note_launch();
  LP_CHECK_CUDA(launch_ex(rope_kv_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, 1, prefill_pdl(),
                          reinterpret_cast<bf16*>(x), ld, S, n_q, n_kv, half, cos_t, sin_t, reinterpret_cast<bf16*>(kc),
                          reinterpret_cast<bf16*>(vc), ld_cache));
  return 0;
}
}  // namespace lp

extern "C" int lp_embed_merge(const long long* ids, int L, long long image_id, int rows_per_slot, const void* table,
                              long long ldt, long long vocab, const void* feats, long long ldf, int n_feat_rows,
                              void* out, long long ldo, int max_out_rows, int D, int* plan /* [2*L + 2] */,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(L > 0 && D % 8 == 0 && rows_per_slot >= 1, "lp_embed_merge: bad L=%d D=%d rows_per_slot=%d", L, D,
             rows_per_slot);
  LP_REQUIRE(ldt % 8 == 0 && ldo % 8 == 0 && (feats == nullptr || ldf % 8 == 0) && aligned16(table) && aligned16(out),
             "lp_embed_merge: 16-byte alignment required");
  LP_REQUIRE(max_out_rows >= L + (long long)(n_feat_rows / rows_per_slot) * (rows_per_slot - 1),
             "lp_embed_merge: output buffer too small for %d feature rows", n_feat_rows);
  int* dst = plan;
  int* src = plan + L;
  int* totals = plan + 2 * L;
  ProfScope prof(5, L, D, rows_per_slot, stream);
  note_launch(2);
  merge_plan_kernel<<<1, 1024, 0, stream>>>(ids, L, image_id, rows_per_slot, dst, src, totals);
  LP_CHECK_CUDA(cudaGetLastError());
  merge_copy_kernel<<<dim3(L, rows_per_slot), 128, 0, stream>>>(ids, dst, src, reinterpret_cast<const bf16*>(table), ldt, vocab,
                                           reinterpret_cast<const bf16*>(feats), ldf, n_feat_rows,
                                           reinterpret_cast<bf16*>(out), ldo, D, rows_per_slot);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_im2col_patch14(const void* pixels, int pixels_fp32, void* out, long long ldo, int N, int C, int H,
                                 int W, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(N > 0 && C > 0 && H % 14 == 0 && W % 14 == 0, "lp_im2col_patch14: bad N=%d C=%d H=%d W=%d", N, C, H, W);
  LP_REQUIRE(ldo >= C * 196 && ldo % 8 == 0 && aligned16(out), "lp_im2col_patch14: ldo=%lld too small/unaligned", ldo);
  const long long total = (long long)N * (H / 14) * (W / 14) * (C * 14 + 1);
  const unsigned grid = (unsigned)((total + 255) / 256);
  ProfScope prof(6, N, H, W, stream);
  note_launch();
  if (pixels_fp32)
    im2col_patch14_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(pixels),
                                                           reinterpret_cast<bf16*>(out), ldo, N, C, H, W);
  else
    im2col_patch14_kernel<bf16><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(pixels),
                                                          reinterpret_cast<bf16*>(out), ldo, N, C, H, W);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_pixel_shuffle(const void* x, long long ldx, void* out, long long ldo, int M, int D, int side,
                                void* stream) {
  LP_REQUIRE(M > 0 && D % 8 == 0 && side > 0 && side % 2 == 0 && M % (side * side) == 0,
             "lp_pixel_shuffle: bad M=%d D=%d side=%d", M, D, side);
  LP_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && aligned16(x) && aligned16(out), "lp_pixel_shuffle: alignment");
  ProfScope prof(7, M, D, side, reinterpret_cast<cudaStream_t>(stream));
  note_launch();
  pixel_shuffle_kernel<<<M, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<bf16*>(out), ldo, side, D);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int lp_cross_entropy(const void* logits, int logits_fp32, long long ld, const long long* labels, float* loss,
                                int rows, int V, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(rows > 0 && V > 0 && logits && labels && loss, "lp_cross_entropy: bad arguments rows=%d V=%d", rows, V);
  ProfScope prof(9, rows, V, 0, stream);
  note_launch();
  if (logits_fp32)
    cross_entropy_kernel<float><<<rows, 256, 0, stream>>>(reinterpret_cast<const float*>(logits), ld, labels, loss, V);
  else
    cross_entropy_kernel<bf16><<<rows, 256, 0, stream>>>(reinterpret_cast<const bf16*>(logits), ld, labels, loss, V);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
