// leopard_b200 — the 2-layer MLP projector as ONE kernel: two chained tcgen05 GEMMs in a single persistent launch.
//
//   out[M, N2] = ( act( X[M, K1] · W1[N1, K1]^T + b1 ) ) · W2[N2, N1]^T + b2          (K2 == N1)
//
// Reference: myLlavaMultiModalProjector.forward (evaluations/models/llava_multiimg_siglip_anyres.py:187-192:
// linear_1 -> GELU -> linear_2 on the pixel-shuffled rows; Megatron twin megatron_patch/model/llava/vlm_model.py:468-475).  X is
// the pixel-shuffled ln_post output — the shuffle itself is an address computation in ln_post's store (lp_layernorm), so
// shuffle + projector cost exactly this one launch.
//
// Structure = the CTA-pair GEMM of lp_gemm.cu (TMA producer warp, one MMA-issuing thread, 8 epilogue warps, double-buffered TMEM
// accumulator, 256 x 256 pair tiles), with a work list that holds the tiles of BOTH layers: [layer-1 tiles | layer-2 tiles], row
// block major, strided over the persistent grid.  The intermediate [M, N1] (39 MB at 28 crops) lives in the caller's scratch and
// stays in the 126 MB L2.  Dataflow instead of a launch boundary: a layer-2 tile of row block m may load its A operand once all
// layer-1 tiles of row block m have LANDED — every epilogue warp waits for its bulk stores, fences and bumps row_done[m]; the
// producer of the layer-2 tile spins on it (bounded) and issues a proxy fence before its TMA loads.  A unit's layer-1 items
// precede its layer-2 items and layer-1 items never wait, so the grid (co-resident: at most one CTA per SM) cannot deadlock.
// What it buys over two launches: no launch gap, and the last wave of layer 1 (304 tiles on 74 pairs = 4.1 waves at 28 crops)
// is filled with layer-2 tiles of row blocks that are already complete: 608 tiles = 8.2 waves instead of 5 + 5.
#include "lp_gemm_common.cuh"

namespace lp {

struct ChainArgs {
  int M;
  int N[2], K[2];
  int num_m_blk, num_n_blk[2], num_k_blk[2];
  const float* bias[2];
  int act[2];
  unsigned* row_done;     // [num_m_blk], zero on entry: epilogue warps whose stores of a layer-1 tile of the row block have landed
  unsigned* exit_count;   // zero on entry; the last CTA to leave re-zeroes the counters (ready for the next launch / graph replay)
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_chain_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ CUtensorMap tmC1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC2, const ChainArgs args) {
  using L = GemmSmem<BN, STAGES, true>;
  constexpr int BMT = 2 * GEMM_BM;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int unit = (int)(blockIdx.x >> 1), n_units = (int)(gridDim.x >> 1);
  constexpr uint32_t TMEM_COLS = 512u;
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
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmC1);
    tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmC2);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  const int tiles1 = args.num_m_blk * args.num_n_blk[0];
  const int tiles_all = tiles1 + args.num_m_blk * args.num_n_blk[1];
  // tile index -> (layer, row block, column block): row block major inside a layer, so row blocks complete early and in order
  auto coords = [&](int t, int& layer, int& m_blk, int& n_blk) {
    layer = t >= tiles1 ? 1 : 0;
    const int local = t - layer * tiles1;
    m_blk = local / args.num_n_blk[layer];
    n_blk = local - m_blk * args.num_n_blk[layer];
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit; t < tiles_all; t += n_units) {
        int layer, m_blk, n_blk;
        coords(t, layer, m_blk, n_blk);
        const CUtensorMap* ta = layer ? &tmA2 : &tmA1;
        const CUtensorMap* tb = layer ? &tmB2 : &tmB1;
        const int a_row = m_blk * BMT + (int)cta_rank * GEMM_BM;
        const int n_rt = min(BN, ((args.N[layer] - n_blk * BN) + 31) & ~31);
        const int b_row = n_blk * BN + (int)cta_rank * (n_rt / 2);
        if (layer == 1) {
          // the intermediate rows of this row block: every epilogue warp (8 per CTA, 2 CTAs) of every layer-1 column block has
          // seen its bulk stores complete; the loads below go through the async proxy, hence the proxy fence after the acquire
          spin_until_gpu(args.row_done + m_blk, (unsigned)(args.num_n_blk[0] * 16));
          fence_proxy_async_all();
        }
        for (int kb = 0; kb < args.num_k_blk[layer]; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_pair(sa, ta, bar, kb * GEMM_BK, a_row);
          tma_load_2d_pair(sb, tb, bar, kb * GEMM_BK, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = unit; t < tiles_all; t += n_units) {
        int layer, m_blk, n_blk;
        coords(t, layer, m_blk, n_blk);
        const uint32_t idesc = umma_idesc_bf16(BMT, min(BN, ((args.N[layer] - n_blk * BN) + 31) & ~31), 0, 0);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < args.num_k_blk[layer]; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(a_addr, 0, 1024);
          const uint64_t bdesc = umma_desc_sw128(a_addr + L::A_BYTES, 0, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9): bias + activation -> bf16 tile stores
    const int lg = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    constexpr int NC = BN / 32;
    const int c_begin = wg == 0 ? 0 : (NC + 1) / 2, c_end = wg == 0 ? (NC + 1) / 2 : NC;
    uint8_t* stage = smem + L::STAGE_OUT_OFF + (warp - 2) * 2048;
    GemmArgs unused = {};
    int as = 0;
    uint32_t aphase = 0;
    for (int t = unit; t < tiles_all; t += n_units) {
      int layer, m_blk, n_blk;
      coords(t, layer, m_blk, n_blk);
      const int N = args.N[layer], act = args.act[layer];
      const float* bias = args.bias[layer];
      const CUtensorMap* tmc = layer ? &tmC2 : &tmC1;
      const int n0 = n_blk * BN;
      float* bs = bias_smem + as * BN;
      if (et < BN) bs[et] = (bias != nullptr && n0 + et < N) ? bias[n0 + et] : 0.f;
      epi_named_barrier();
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN;
      const int row_base = m_blk * BMT + (int)cta_rank * GEMM_BM + lg * 32;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= N) break;
        uint32_t r[32];
        LP_TMEM_LD_32(r, taddr0 + c * 32);
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(bs + c * 32 + j);
          v[j] = __uint_as_float(r[j]) + b.x; v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
          v[j + 2] = __uint_as_float(r[j + 2]) + b.z; v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
        }
        if (act == LP_ACT_GELU_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
        } else if (act == LP_ACT_GELU_ERF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
        }
        staged_store_chunk(unused, tmc, stage, v, nullptr, row_base, col0, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
        else mbar_arrive(&tempty_bar[as]);
        if (layer == 0) {
          // publish: this warp's part of the intermediate tile is in memory (bulk stores complete, not merely read)
          tma_store_wait<0>();
          fence_proxy_async_all();
          __threadfence();
          atomicAdd(args.row_done + m_blk, 1u);
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(args.exit_count, 1u) == gridDim.x - 1) {   // every CTA is past its last read of row_done
      for (int m = 0; m < args.num_m_blk; ++m) args.row_done[m] = 0u;
      *args.exit_count = 0u;
    }
  }
}

#ifndef LP_EMULATE
// counters live in the stream-K workspace's flag page (lp_gemm.cu): words [512, 512 + num_m_blk] of the first 4 KiB
unsigned* gemm_flag_page();

int launch_gemm_chain(const void* X, long long ldx, const void* W1, const float* b1, int act1, void* mid, long long ld_mid,
                      const void* W2, const float* b2, void* out, long long ldo, int M, int N1, int K1, int N2, cudaStream_t stream) {
  constexpr int BN = 256, STAGES = 6;
  using L = GemmSmem<BN, STAGES, true>;
  unsigned* page = gemm_flag_page();
  LP_REQUIRE(page != nullptr, "lp_project: chained launch needs the GEMM flag workspace (lp_gemm_set_streamk_workspace)");
  ChainArgs a = {};
  a.M = M;
  a.N[0] = N1; a.K[0] = K1; a.N[1] = N2; a.K[1] = N1;
  a.num_m_blk = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  for (int l = 0; l < 2; ++l) {
    a.num_n_blk[l] = (a.N[l] + BN - 1) / BN;
    a.num_k_blk[l] = (a.K[l] + GEMM_BK - 1) / GEMM_BK;
  }
  LP_REQUIRE(a.num_m_blk <= 448, "lp_project: %d row blocks exceed the chained launch's counter page", a.num_m_blk);
  a.bias[0] = b1; a.bias[1] = b2;
  a.act[0] = act1; a.act[1] = LP_ACT_NONE;
  a.row_done = page + 512;
  a.exit_count = page + 511;
  CUtensorMap tmA1, tmB1, tmC1, tmA2, tmB2, tmC2;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmA1, X, (uint64_t)K1, (uint64_t)M, (uint64_t)ldx, GEMM_BK, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB1, W1, (uint64_t)K1, (uint64_t)N1, (uint64_t)K1, GEMM_BK, BN / 2))) return rc;
  if ((rc = make_tmap_2d(&tmC1, mid, 2, (uint64_t)N1, (uint64_t)M, (uint64_t)ld_mid, 32, 32, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmA2, mid, (uint64_t)N1, (uint64_t)M, (uint64_t)ld_mid, GEMM_BK, GEMM_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB2, W2, (uint64_t)N1, (uint64_t)N2, (uint64_t)N1, GEMM_BK, BN / 2))) return rc;
  if ((rc = make_tmap_2d(&tmC2, out, 2, (uint64_t)N2, (uint64_t)M, (uint64_t)ldo, 32, 32, 64))) return rc;
  auto kern = gemm_chain_kernel<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = a.num_m_blk * (a.num_n_blk[0] + a.num_n_blk[1]);
  const int pairs = sm_count() / 2;
  ProfScope prof(1, M, N1 + N2, (K1 + N1) / 2, stream);   // one record: both layers
  note_launch();
  LP_CHECK_CUDA(launch_ex(kern, dim3(2 * (tiles < pairs ? tiles : pairs)), dim3(GEMM_THREADS), L::TOTAL, stream, 2, prefill_pdl(), tmA1,
                          tmB1, tmC1, tmA2, tmB2, tmC2, a));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
#endif

}  // namespace lp
