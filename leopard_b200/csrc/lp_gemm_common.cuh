// leopard_b200 — pieces shared by the tcgen05 GEMM kernels (lp_gemm.cu: one CTA or a CTA pair per tile; lp_gemm4.cu: two pairs
// per 4-CTA cluster sharing the A tile by TMA multicast): tile geometry, launch arguments, shared-memory plan, tile
// rasterisation and the epilogue (bias / activation / residual / fp32 or swizzled-staging bf16 stores).
#pragma once
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_EPI_THREADS = 256;

struct GemmArgs {
  int M, N, K;
  void* C;
  long long ldc;
  int out_fp32;
  const float* bias;
  const bf16* residual;
  long long ldr;
  int res_row_mod;
  int act;
  int num_m_blk, num_n_blk, num_k_blk;
  int group_m;
  int prefetch_a;
  unsigned epi_sleep_ns;         // epilogue warps sleep this long between probes of the accumulator-full barrier (0 = spin)
  // ---- fused row normalisation of the A operand (consumer side; SURVEY.md K2/K5/K9: LayerNorm / RMSNorm folded into the GEMM).
  // The gain is folded into W offline (W' = W diag(gamma)), so norm(x) W^T = rs[row] * (x W'^T - mu[row] * colsum[col]) (+ beta W^T
  // + b, passed as `bias`).  mu / rs come from per-row partial (sum, sum of squares) pairs the PRODUCING GEMM's epilogue wrote.
  int norm_mode;                 // LP_NORM_NONE / LP_NORM_RMS / LP_NORM_LAYER
  const float2* stats_in;        // [M, stats_in_parts]
  int stats_in_parts;
  const float* colsum;           // [N] fp32: sum_k W'[n, k] (layer norm only)
  float norm_eps, norm_inv_k;    // 1 / (features per normalised row)
  // ---- producer side: per-row (sum, sum of squares) of the bf16-ROUNDED outputs of this GEMM, one pair per (column block,
  // epilogue warpgroup): stats_out[row * 2 * num_n_blk + 2 * n_blk + wg]; summed in index order by the consumer (deterministic)
  float2* stats_out;
  // ---- rotate-half RoPE on the first `rope_heads` 128-column heads (decoder QKV; tables [M, 64] fp32; BN = 256 only)
  const float* rope_cos;
  const float* rope_sin;
  int rope_heads;
  // ---- tensor-parallel fusions (lp_gemm_bf16_tp).  tp_size > 0: the M rows are tp_size segments of seg_rows rows,
  // segment r owned by rank r (Megatron sequence parallelism); tiles never straddle a segment (seg_m_blk row blocks per segment)
  // and the rank's own segment is scheduled first.
  int tp_size, tp_rank, seg_rows, seg_m_blk;
  unsigned epoch;                 // call number: what the cross-rank signal words are compared with
  // all-gather prologue (ag_counter != null): the epilogue warps of all CTAs copy the other ranks' rows of A from peer memory
  // (ag_src[r] = rank r's copy of the A buffer, NVLink loads) into the local A buffer while the first tiles (own rows) run
  const void* ag_src[8];
  void* ag_dst;
  long long ag_ld;                // row stride (elements) of the A buffers, identical on every rank
  int ag_cols;                    // columns to copy (K)
  unsigned* ag_counter;           // local, zero on entry: CTAs that finished their share of the copy
  const unsigned* ag_wait[8];     // local signal words, written by rank r: >= epoch once its rows may be read
  unsigned* ag_post[8];           // rank r's signal word for THIS rank (remote store): "my rows are ready"
  // reduce-scatter epilogue (rs_counter != null): the tile of segment r is stored through GemmSegMaps.m[r] (rank r's staging memory,
  // TMA stores over NVLink); the last CTA to finish posts `epoch` to every owner
  unsigned* rs_counter;           // local, zero on entry
  unsigned* rs_post[8];           // rank r's "partials from this rank have landed" word (remote store)
  unsigned* recycle;              // a counter slot of an older call to zero for reuse
  // ---- stream-K tail (sk_tiles > 0; lp_gemm_set_streamk_workspace): the first dp_tiles tiles (a multiple of the unit count) are
  // scheduled one tile per unit per wave as usual; the K loops of the last sk_tiles tiles (n_units <= sk_tiles < 2 n_units) are
  // laid end to end and cut into n_units equal ranges, so every unit ends at the same time instead of a mostly idle last wave.
  // A tile cut between units u and u + 1 is STARTED by u (right after its data-parallel tiles: k-blocks [0, b), raw fp32 accumulator
  // dumped to sk_ws) and FINISHED by u + 1 (its last item: the dump is loaded back into TMEM and the k loop continues at b) — the
  // accumulation order of every output element is exactly that of the unsplit kernel, so results are bit-identical to it.
  int sk_tiles, dp_tiles;
  float* sk_ws;                   // [n_units][CTAs per unit][BN / 32 chunks][8][128 rows][4] fp32
  unsigned* sk_flags;             // [n_units][CTAs per unit]: 1 while the unit's START dump waits to be consumed
};

// ---- work items of one unit (a CTA or a CTA pair)
enum { GEMM_ITEM_FULL = 0, GEMM_ITEM_START = 1, GEMM_ITEM_END = 2 };
struct GemmItem {
  int tile, kb0, kb1, kind;
};
struct GemmSchedule {
  int n_items, unit, n_units;
  int dp_items;                   // data-parallel tiles of this unit: tile = unit + i * n_units
  int has_start, has_end, n_whole;
  int start_tile, start_kb1;      // START piece: k-blocks [0, start_kb1) of start_tile
  int end_tile, end_kb0;          // END piece: k-blocks [end_kb0, num_k_blk) of end_tile
  int whole0;                     // first whole stream-K tile
};

__device__ __forceinline__ GemmSchedule gemm_schedule(const GemmArgs& a, int unit, int n_units) {
  GemmSchedule s = {};
  s.unit = unit;
  s.n_units = n_units;
  const int num_tiles = a.num_m_blk * a.num_n_blk;
  if (a.sk_tiles <= 0) {
    s.dp_items = unit < num_tiles ? (num_tiles - unit + n_units - 1) / n_units : 0;
    s.n_items = s.dp_items;
    return s;
  }
  const int KB = a.num_k_blk;
  const long long total = (long long)a.sk_tiles * KB;
  const long long b = (long long)unit * total / n_units, e = (long long)(unit + 1) * total / n_units;  // e - b >= KB
  const int t_first = (int)(b / KB), b_off = (int)(b - (long long)t_first * KB);
  const int t_last = (int)((e - 1) / KB), e_off = (int)(e - (long long)t_last * KB);   // 1 .. KB
  s.dp_items = a.dp_tiles / n_units;
  s.has_end = b_off > 0;
  s.has_start = e_off < KB;
  s.end_tile = a.dp_tiles + t_first;
  s.end_kb0 = b_off;
  s.start_tile = a.dp_tiles + t_last;
  s.start_kb1 = e_off;
  s.whole0 = a.dp_tiles + t_first + s.has_end;
  s.n_whole = (t_last - s.has_start) - (t_first + s.has_end) + 1;
  s.n_items = s.has_start + s.dp_items + s.n_whole + s.has_end;
  return s;
}

// item order: [data-parallel tiles] [START piece] [whole stream-K tiles] [END piece].  The data-parallel waves run exactly as in
// the plain schedule (all units in step along K, which is what keeps the operand slices shared through L2).  The START piece comes
// before the whole tiles so that its dump is in memory before the neighbouring unit reaches its END item (by construction that
// happens (range - num_k_blk) k-blocks later); unit 0 never has an END item and all units are co-resident (persistent grid), so
// the waits cannot form a cycle.
__device__ __forceinline__ GemmItem gemm_item(const GemmArgs& a, const GemmSchedule& s, int i) {
  GemmItem w;
  w.kb0 = 0;
  w.kb1 = a.num_k_blk;
  w.kind = GEMM_ITEM_FULL;
  if (i < s.dp_items) { w.tile = s.unit + i * s.n_units; return w; }
  i -= s.dp_items;
  if (s.has_start) {
    if (i == 0) {
      w.tile = s.start_tile; w.kb1 = s.start_kb1; w.kind = GEMM_ITEM_START;
      return w;
    }
    --i;
  }
  if (i < s.n_whole) { w.tile = s.whole0 + i; return w; }
  w.tile = s.end_tile; w.kb0 = s.end_kb0; w.kind = GEMM_ITEM_END;
  return w;
}


struct GemmSegMaps {
  CUtensorMap m[8];
};

enum { LP_NORM_NONE_ = 0, LP_NORM_RMS_ = 1, LP_NORM_LAYER_ = 2 };

#ifndef LP_EMULATE
// 4-CTA-cluster variant (lp_gemm4.cu); args.num_n_blk counts column-block pairs, tmAh has 64-row boxes
int launch_gemm4(const CUtensorMap& tmAh, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmArgs& args, int block_n,
                 cudaStream_t stream);
#endif

// EPI_BUFS: 2 KiB staging tiles per epilogue warp — 1 = output staging only; 3 = ring shared by the TMA-loaded residual tile and
// the output tile (the TMARES kernels: residual chunk g + 1 lands while chunk g is added in place and chunk g - 1 is stored)
template <int BN, int STAGES, bool CTA2 = false, int EPI_BUFS = 1>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;  // CTA pair: each CTA stages half of the W tile
  static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int RES_BAR_OFF = BAR_OFF + 256;                                // 8 warps x 3 residual-tile barriers
  static constexpr int BIAS_OFF = BAR_OFF + 512;
  static constexpr int COLSUM_OFF = BIAS_OFF + 2 * BN * 4;                         // layer-norm column sums, double-buffered like bias
  static constexpr int STAGE_OUT_OFF = (COLSUM_OFF + 2 * BN * 4 + 1023) & ~1023;  // 8 warps x EPI_BUFS x 2 KiB staging
  static constexpr int TOTAL = STAGE_OUT_OFF + 8 * EPI_BUFS * 2048 + 1024;  // +1024: manual 1 KiB alignment of the base
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void epi_named_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// tile index -> (m_blk, n_blk): groups of `group_m` row-blocks, row-block fastest inside a group
__device__ __forceinline__ void tile_coords(const GemmArgs& a, int tile, int& m_blk, int& n_blk) {
  const int per_group = a.group_m * a.num_n_blk;
  const int g = tile / per_group, r = tile - g * per_group;
  const int m0 = g * a.group_m;
  const int gsz = min(a.group_m, a.num_m_blk - m0);
  n_blk = r / gsz;
  m_blk = m0 + (r - n_blk * gsz);
}

// first row of a row block, its segment (tensor-parallel launches: own segment first) and the row limit of that segment
__device__ __forceinline__ int tile_row0(const GemmArgs& a, int m_blk, int rows_per_blk, int& seg, int& row_limit) {
  if (a.tp_size <= 0) {
    seg = 0;
    row_limit = a.M;
    return m_blk * rows_per_blk;
  }
  const int s = m_blk / a.seg_m_blk;
  seg = (s + a.tp_rank) % a.tp_size;
  row_limit = min(a.M, (seg + 1) * a.seg_rows);
  return seg * a.seg_rows + (m_blk - s * a.seg_m_blk) * rows_per_blk;
}



// applies bias/activation/residual to 32 accumulator columns of one row and stores them
__device__ __forceinline__ void epilogue_store_chunk(const GemmArgs& args, float* v, const float* bias_s, int row,
                                                     int rrow, bool row_ok, int col0, int n_limit) {
  if (args.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias_s + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (args.act == LP_ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
  } else if (args.act == LP_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  }
  if (!row_ok) return;
  if (args.residual) {
    const bf16* rp = args.residual + (long long)rrow * args.ldr + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (col0 + g * 8 < n_limit) {
        const uint4 u = *reinterpret_cast<const uint4*>(rp + g * 8);
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                     f3 = unpack_bf16x2(u.w);
        float* w = v + g * 8;
        w[0] += f0.x; w[1] += f0.y; w[2] += f1.x; w[3] += f1.y;
        w[4] += f2.x; w[5] += f2.y; w[6] += f3.x; w[7] += f3.y;
      }
    }
  }
  if (args.out_fp32) {
    float* cp = reinterpret_cast<float*>(args.C) + (long long)row * args.ldc + col0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (col0 + g * 4 < n_limit)
        *reinterpret_cast<float4*>(cp + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    }
  } else {
    bf16* cp = reinterpret_cast<bf16*>(args.C) + (long long)row * args.ldc + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (col0 + g * 8 < n_limit) {
        const float* w = v + g * 8;
        uint4 u;
        u.x = pack_bf16x2(w[0], w[1]); u.y = pack_bf16x2(w[2], w[3]);
        u.z = pack_bf16x2(w[4], w[5]); u.w = pack_bf16x2(w[6], w[7]);
        *reinterpret_cast<uint4*>(cp + g * 8) = u;
      }
    }
  }
}

// ---- staged epilogue (bf16 outputs) ------------------------------------------------------------------------------
// Each epilogue warp owns a 32-row x 32-column bf16 staging tile (2 KiB, 64-byte rows, TMA SWIZZLE_64B layout:
// 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)).  Residual rows are fetched with coalesced 16-byte loads
// (4 lanes per row), handed to the owning thread through the tile, the finished rows go back into the tile and leave
// with one TMA store per chunk (the tensor map clips rows >= M / columns >= N', so no store predication is needed).
__device__ __forceinline__ uint32_t stage_off(int row, int chunk) {
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

struct ResidualRegs { uint4 v[4]; };

// coalesced fetch of the residual chunk [32 rows x 32 cols] for rows row_base .. row_base+31, columns col0 .. col0+31
__device__ __forceinline__ void residual_fetch(const GemmArgs& args, ResidualRegs& rr, int row_base, int col0, int n_limit,
                                               int lane) {
  const int seg = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row_base + i * 8 + (lane >> 2);
    rr.v[i] = make_uint4(0u, 0u, 0u, 0u);
    if (row < args.M && col0 + seg * 8 < n_limit) {
      const int rrow = args.res_row_mod > 0 ? row % args.res_row_mod : row;
      rr.v[i] = *reinterpret_cast<const uint4*>(args.residual + (long long)rrow * args.ldr + col0 + seg * 8);
    }
  }
}

__device__ __forceinline__ void staged_store_chunk(const GemmArgs& args, const CUtensorMap* tmC, uint8_t* stage,
                                                   float* v, const ResidualRegs* rr, int row_base, int col0, int lane,
                                                   float2* row_stats = nullptr) {
  // previous TMA store of this warp must have finished reading the staging tile
  if (lane == 0) tma_store_wait_read<0>();
  __syncwarp();
  if (rr != nullptr) {
    const int seg = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<uint4*>(stage + stage_off(i * 8 + (lane >> 2), seg)) = rr->v[i];
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 u = *reinterpret_cast<const uint4*>(stage + stage_off(lane, g));
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      float* w = v + g * 8;
      w[0] += f0.x; w[1] += f0.y; w[2] += f1.x; w[3] += f1.y;
      w[4] += f2.x; w[5] += f2.y; w[6] += f3.x; w[7] += f3.y;
    }
    __syncwarp();
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float* w = v + g * 8;
    uint4 u;
    u.x = pack_bf16x2(w[0], w[1]); u.y = pack_bf16x2(w[2], w[3]);
    u.z = pack_bf16x2(w[4], w[5]); u.w = pack_bf16x2(w[6], w[7]);
    *reinterpret_cast<uint4*>(stage + stage_off(lane, g)) = u;
    if (row_stats != nullptr) {  // statistics of what the consumer will read: the rounded values
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      row_stats->x += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
      row_stats->y += ((f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y)) +
                      ((f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y));
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmC, stage, col0, row_base);
    tma_store_commit();
  }
}

// TMARES kernels: `buf` (32 x 32 bf16, SWIZZLE_64B) holds the residual tile delivered by TMA; every thread adds its own row to the
// 32 finished values, writes the rounded sums back in place (a thread touches only its own row: no intra-warp ordering needed) and
// the tile leaves from the same buffer with one TMA store
__device__ __forceinline__ void ring_add_store_chunk(const CUtensorMap* tmC, uint8_t* buf, float* v, int row_base, int col0, int lane,
                                                     float2* row_stats) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint8_t* p = buf + stage_off(lane, g);
    const uint4 r = lds_u4(p);
    const float2 r0 = unpack_bf16x2(r.x), r1 = unpack_bf16x2(r.y), r2 = unpack_bf16x2(r.z), r3 = unpack_bf16x2(r.w);
    const float* w = v + g * 8;
    uint4 u;
    u.x = pack_bf16x2(w[0] + r0.x, w[1] + r0.y); u.y = pack_bf16x2(w[2] + r1.x, w[3] + r1.y);
    u.z = pack_bf16x2(w[4] + r2.x, w[5] + r2.y); u.w = pack_bf16x2(w[6] + r3.x, w[7] + r3.y);
    sts_u4(p, u);
    if (row_stats != nullptr) {  // statistics of what the consumer will read: the rounded values (same order as staged_store_chunk)
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      row_stats->x += ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
      row_stats->y += ((f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y)) +
                      ((f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y));
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmC, buf, col0, row_base);
    tma_store_commit();
  }
}

// mu / rstd of one A row from the producer's partial sums (fixed summation order -> bit-deterministic)
__device__ __forceinline__ void row_norm_coeffs(const GemmArgs& args, int row, bool row_ok, float& rs, float& nmu) {
  rs = 1.f;
  nmu = 0.f;
  if (args.norm_mode == LP_NORM_NONE_ || !row_ok) return;
  const float2* p = args.stats_in + (long long)row * args.stats_in_parts;
  float s1 = 0.f, s2 = 0.f;
  int i = 0;
  for (; i + 4 <= args.stats_in_parts; i += 4) {   // four partials in flight at once; the additions keep the index order
    const float2 t0 = p[i], t1 = p[i + 1], t2 = p[i + 2], t3 = p[i + 3];
    s1 += t0.x; s2 += t0.y;
    s1 += t1.x; s2 += t1.y;
    s1 += t2.x; s2 += t2.y;
    s1 += t3.x; s2 += t3.y;
  }
  for (; i < args.stats_in_parts; ++i) {
    const float2 t = p[i];
    s1 += t.x;
    s2 += t.y;
  }
  if (args.norm_mode == LP_NORM_RMS_) {
    rs = rsqrtf(s2 * args.norm_inv_k + args.norm_eps);
  } else {
    const float mu = s1 * args.norm_inv_k;
    const float var = fmaxf(s2 * args.norm_inv_k - mu * mu, 0.f);
    rs = rsqrtf(var + args.norm_eps);
    nmu = -rs * mu;
  }
}

// v = rs * acc + nmu * colsum  (before bias / activation)
__device__ __forceinline__ void apply_row_norm(const GemmArgs& args, float* v, const float* colsum_s, float rs, float nmu) {
  if (args.norm_mode == LP_NORM_RMS_) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rs;
  } else if (args.norm_mode == LP_NORM_LAYER_) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 c = *reinterpret_cast<const float4*>(colsum_s + j);
      v[j] = fmaf(rs, v[j], nmu * c.x); v[j + 1] = fmaf(rs, v[j + 1], nmu * c.y);
      v[j + 2] = fmaf(rs, v[j + 2], nmu * c.z); v[j + 3] = fmaf(rs, v[j + 3], nmu * c.w);
    }
  }
}

__device__ __forceinline__ void apply_bias_act(const GemmArgs& args, float* v, const float* bias_s) {
  if (args.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias_s + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (args.act == LP_ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
  } else if (args.act == LP_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  }
}

}  // namespace lp
