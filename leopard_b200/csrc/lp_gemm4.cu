// leopard_b200 — tcgen05 GEMM on 4-CTA clusters: two CTA pairs on column-adjacent 256 x BN tiles share the A tile by TMA
// multicast (sm_100a).
//
// STATUS: written after the round's GPU budget was spent; logic verified under the CPU emulator
// (tests/test_gemm_kernel_emulated.py), NOT yet run on hardware.  Opt-in: lp_gemm_set_pair_mode(2).
//
// Why: in full waves the pair kernel (lp_gemm.cu) is bound by the operand feed — 25.6 us per 256 x 256 x 4096 tile against
// 17.3 us at the MMA rate.  Per k-block each CTA pulls 16 KB of A and 16 KB of W out of L2.  Here the cluster's
// two pairs work on tiles (m, 2 n2) and (m, 2 n2 + 1): both need the same 256 rows of A, so every CTA requests only HALF of
// its 128-row A tile (64 rows, 8 KB) and the TMA multicasts it to the CTA with the same row half in the other pair.  L2 reads
// per tile drop from 64 KB to 48 KB per k-block (-25 %); shared-memory fill and the MMA schedule are unchanged.
//   cluster rank = 2 p + r : pair p (column block 2 n2 + p), row half r inside the pair (cta_group::2 pairs are ranks 2p, 2p+1)
//   full[stage]   on each pair leader: its own arrive.expect_tx(2 x stage bytes); bytes arrive from the four multicast A halves
//                 that land in the pair's two CTAs (two of them issued by the OTHER pair) and the pair's two W loads
//   empty[stage]  on every CTA, count 2: one tcgen05.commit per pair, multicast to all four CTAs — a slot is rewritten by
//                 producers of both pairs, so both pairs' MMAs must have retired
//   tfull / tempty (accumulator hand-off) stay per pair.  A pair whose column block lies beyond N (odd number of column
//   blocks) keeps feeding the protocol — its A halves, zero-filled W boxes, empty commits — and skips MMA and epilogue.
// Epilogue, tile rasterisation (over column-block PAIRS) and the shared-memory plan are those of the pair kernel.
#include "lp_gemm_common.cuh"

namespace lp {

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm4_bf16_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const GemmArgs args) {
  using L = GemmSmem<BN, STAGES, true>;
  constexpr int BMT = 2 * GEMM_BM;
  const uint32_t cta_rank = cluster_ctarank();
  const uint32_t pair = cta_rank >> 1, r = cta_rank & 1u;
  const uint32_t leader_rank = cta_rank & ~1u;
  const bool leader = r == 0;
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
  const uint16_t a_mask = (uint16_t)((1u << r) | (1u << (2 + r)));   // the CTAs holding row half r: one in each pair
  const int unit = (int)(blockIdx.x >> 2), n_units = (int)(gridDim.x >> 2);
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256u : 512u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_smem = reinterpret_cast<float*>(smem + L::BIAS_OFF);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmB);
    if (!args.out_fp32) tma_prefetch_desc(&tmC);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 2);   // one commit per pair
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);  // one arrive per epilogue warp of both CTAs of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  // work units are (row block, column-block PAIR): args.num_n_blk counts pairs; this CTA's pair takes column block 2 n2 + pair
  const int num_tiles = args.num_m_blk * args.num_n_blk;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < num_tiles; tile += n_units) {
        int m_blk, n2;
        tile_coords(args, tile, m_blk, n2);
        const int n_blk = 2 * n2 + (int)pair;
        const bool valid = n_blk * BN < args.N;
        // this CTA requests rows [64 pair, 64 pair + 64) of the 128-row A tile that belongs to row half r
        const int a_row = m_blk * BMT + (int)r * GEMM_BM + (int)pair * (GEMM_BM / 2);
        const int n_rt = valid ? min(BN, ((args.N - n_blk * BN) + 31) & ~31) : BN;
        const int b_row = n_blk * BN + (int)r * (n_rt / 2);   // beyond N for an invalid pair: zero-filled boxes
        for (int kb = 0; kb < args.num_k_blk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), leader_rank);
          tma_load_2d_pair_mcast(sa + pair * (L::A_BYTES / 2), &tmAh, bar, kb * GEMM_BK, a_row, a_mask);
          tma_load_2d_pair(sb, &tmB, bar, kb * GEMM_BK, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (each pair's leader)
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = unit; tile < num_tiles; tile += n_units) {
        int m_blk_unused, n2;
        tile_coords(args, tile, m_blk_unused, n2);
        const int n_blk = 2 * n2 + (int)pair;
        const bool valid = n_blk * BN < args.N;
        const uint32_t idesc = umma_idesc_bf16(BMT, valid ? min(BN, ((args.N - n_blk * BN) + 31) & ~31) : BN, 0, 0);
        if (valid) {
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < args.num_k_blk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (valid) {
            const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
            const uint64_t adesc = umma_desc_sw128(a_addr, 0, 1024);
            const uint64_t bdesc = umma_desc_sw128(a_addr + L::A_BYTES, 0, 1024);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // the slot is free for all four producers once BOTH pairs have committed (with nothing issued the commit
          // arrives at once)
          umma_commit_mask(&empty_bar[stage], (uint16_t)0xF);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (valid) {
          umma_commit_mask(&tfull_bar[as], pair_mask);
          as ^= 1;
          if (as == 0) aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9), as in the pair kernel
    const int lg = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    int as = 0;
    uint32_t aphase = 0;
    const bool swiglu = args.act == LP_ACT_SWIGLU;
    constexpr int NC = BN / 32;
    constexpr int NP = BN / 64;
    for (int tile = unit; tile < num_tiles; tile += n_units) {
      int m_blk, n2;
      tile_coords(args, tile, m_blk, n2);
      const int n_blk = 2 * n2 + (int)pair;
      if (n_blk * BN >= args.N) continue;   // this pair has no tile in the last (odd) column-block pair
      const int m0 = m_blk * BMT + (int)r * GEMM_BM, n0 = n_blk * BN;
      float* bs = bias_smem + as * BN;
      if (args.bias) {
        for (int i = et; i < BN; i += GEMM_EPI_THREADS) bs[i] = (n0 + i < args.N) ? args.bias[n0 + i] : 0.f;
      }
      epi_named_barrier();
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m0 + lg * 32 + lane;
      const bool row_ok = row < args.M;
      const int rrow = args.res_row_mod > 0 ? row % args.res_row_mod : row;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN;
      uint8_t* stage = smem + L::STAGE_OUT_OFF + (warp - 2) * 2048;
      const int row_base = m0 + lg * 32;
      const bool has_res = args.residual != nullptr;
      if (!swiglu) {
        const int c_begin = wg == 0 ? 0 : (NC + 1) / 2, c_end = wg == 0 ? (NC + 1) / 2 : NC;
        ResidualRegs rr;
        if (has_res && !args.out_fp32 && n0 + c_begin * 32 < args.N)
          residual_fetch(args, rr, row_base, n0 + c_begin * 32, args.N, lane);
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= args.N) break;
          uint32_t rg[32];
          LP_TMEM_LD_32(rg, taddr0 + c * 32);
          tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rg[j]);
          if (args.out_fp32) {
            epilogue_store_chunk(args, v, bs + c * 32, row, rrow, row_ok, col0, args.N);
          } else {
            apply_bias_act(args, v, bs + c * 32);
            ResidualRegs cur = rr;
            if (has_res && c + 1 < c_end && col0 + 32 < args.N)
              residual_fetch(args, rr, row_base, col0 + 32, args.N, lane);
            staged_store_chunk(args, &tmC, stage, v, has_res ? &cur : nullptr, row_base, col0, lane);
          }
        }
      } else {
        const int c_begin = wg == 0 ? 0 : (NP + 1) / 2, c_end = wg == 0 ? (NP + 1) / 2 : NP;
        ResidualRegs rr;
        if (has_res && n0 + c_begin * 64 < args.N)
          residual_fetch(args, rr, row_base, (n0 + c_begin * 64) / 2, args.N / 2, lane);
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 64;
          if (col0 >= args.N) break;
          uint32_t rg[32], ru[32];
          LP_TMEM_LD_32(rg, taddr0 + c * 64);
          LP_TMEM_LD_32(ru, taddr0 + c * 64 + 32);
          tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float g = __uint_as_float(rg[j]), u = __uint_as_float(ru[j]);
            if (args.bias) { g += bs[c * 64 + j]; u += bs[c * 64 + 32 + j]; }
            v[j] = silu_f(g) * u;
          }
          ResidualRegs cur = rr;
          if (has_res && c + 1 < c_end && col0 + 64 < args.N)
            residual_fetch(args, rr, row_base, (col0 + 64) / 2, args.N / 2, lane);
          staged_store_chunk(args, &tmC, stage, v, has_res ? &cur : nullptr, row_base, col0 / 2, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), leader_rank));
        else mbar_arrive(&tempty_bar[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();   // MMAs read the pair peer's smem and multicast loads write the other pair's: leave together
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

#ifndef LP_EMULATE
template <int BN, int STAGES>
static int launch4(const CUtensorMap& tmAh, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmArgs& args,
                   cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, true>;
  auto kern = gemm4_bf16_kernel<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  // persistent kernel with a static round-robin schedule: the grid must not exceed the clusters that are co-resident (GPC
  // boundaries can leave fewer than SMs / 4 slots for 4-CTA clusters)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * (sm_count() / 4));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    LP_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    max_clusters = n > 0 ? n : 1;
  }
  const int tiles = args.num_m_blk * args.num_n_blk;
  int clusters = sm_count() / 4;
  if (clusters > max_clusters) clusters = max_clusters;
  LP_CHECK_CUDA(launch_ex(kern, dim3(4 * (tiles < clusters ? tiles : clusters)), dim3(GEMM_THREADS), L::TOTAL, stream, 4,
                          prefill_pdl(), tmAh, tmB, tmC, args));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// args as prepared by lp_gemm_bf16 for pair mode, except num_n_blk = column-block PAIRS; tmAh: 64-row A boxes
int launch_gemm4(const CUtensorMap& tmAh, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmArgs& args, int block_n,
                 cudaStream_t stream) {
  switch (block_n) {
    case 256: return launch4<256, 6>(tmAh, tmB, tmC, args, stream);
    case 192: return launch4<192, 7>(tmAh, tmB, tmC, args, stream);
    case 128: return launch4<128, 8>(tmAh, tmB, tmC, args, stream);
    default: set_error("launch_gemm4: block_n %d is not instantiated (128, 192, 256)", block_n); return -1;
  }
}
#endif

}  // namespace lp
