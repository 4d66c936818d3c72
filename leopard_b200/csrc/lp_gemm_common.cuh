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
};

#ifndef LP_EMULATE
// 4-CTA-cluster variant (lp_gemm4.cu); args.num_n_blk counts column-block pairs, tmAh has 64-row boxes
int launch_gemm4(const CUtensorMap& tmAh, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmArgs& args, int block_n,
                 cudaStream_t stream);
#endif

template <int BN, int STAGES, bool CTA2 = false>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;  // CTA pair: each CTA stages half of the W tile
  static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int BIAS_OFF = BAR_OFF + 256;
  static constexpr int STAGE_OUT_OFF = (BIAS_OFF + 2 * BN * 4 + 1023) & ~1023;  // 8 warps x 2 KiB output staging
  static constexpr int TOTAL = STAGE_OUT_OFF + 8 * 2048 + 1024;  // +1024: manual 1 KiB alignment of the base
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
                                                   float* v, const ResidualRegs* rr, int row_base, int col0, int lane) {
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
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmC, stage, col0, row_base);
    tma_store_commit();
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
