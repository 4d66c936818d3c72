// leopard_b200 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] · W[N,K]^T )          A, W bf16 row-major (K contiguous), fp32 accumulate in TMEM
//
// Replaces every dense contraction on Leopard's prefill path, which the reference delegates to cuBLAS /
// TransformerEngine through torch:
//   K1  patch-embed GEMM + bias + position embedding   (HF SiglipVisionEmbeddings; Megatron
//       Pai-Megatron-Patch/megatron_patch/model/idefics2/idefics_vision_tower.py:118-150)
//   K2/K4/K5 ViT qkv / out-proj / fc1(GELU-tanh) / fc2  (megatron_patch/model/llava/clip_encoder.py:51-72)
//   K7  projector linear_1(GELU) / linear_2            (evaluations/models/llava_multiimg_siglip_anyres.py:179-192)
//   K10/K13/K14/K15/K16 decoder qkv / o / gate‖up(SwiGLU) / down / lm_head
//       (megatron_patch/model/llava/transformer.py:116-176,577-624)
//
// Structure (one CTA per SM, static round-robin tile scheduler):
//   warp 0      TMA producer   : cp.async.bulk.tensor A/W k-blocks -> 128B-swizzled smem ring (STAGES deep)
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered
//                                TMEM accumulator (2 x BN columns), tcgen05.commit frees smem / signals epilogue
//   warps 2-9   epilogue       : two warpgroups split the tile's columns; tcgen05.ld accumulator rows -> bias /
//                                activation / residual -> 16-byte stores
// Tiles are rasterised in groups of `group_m` row-blocks so that whichever operand does not fit the 126 MB L2
// is streamed from HBM once (host heuristic), and the tile width BN in {128,160,192,224,256} is chosen per problem
// to minimise wave quantisation (ceil(tiles / SMs) * BN).
// Ragged M, N, K are handled by TMA zero fill on loads and predication on stores.
#include "lp_gemm_common.cuh"

namespace lp {

// CTA2 = true: CTA pair (cluster of 2, tcgen05 cta_group::2).  The pair owns a 256 x BN tile: CTA r stages rows
// [m0 + 128 r, +128) of A and rows [n0 + r BN/2, + BN/2) of W, the leader's MMA thread issues M = 256 MMAs that read
// both CTAs' shared memory and write each CTA's 128 accumulator rows into its own TMEM.  Halving the W bytes staged
// per SM relieves the per-SM L2->SMEM bandwidth (the 1-CTA kernel runs at ~75 % of it) and deepens the ring.
// TP = true: tensor-parallel launch (lp_gemm_bf16_tp) — segmented rows, output through per-segment tensor maps `segC`, optional
// all-gather prologue and reduce-scatter signalling (see GemmArgs)
// TMARES = true: the residual launch of the layer loops (bf16 output, residual rows == output rows, no SwiGLU / RoPE): the residual
// tile travels by TMA (map tmR) into a 3-deep per-warp ring one chunk ahead of its use, is added in place and leaves from the same
// buffer — no residual registers, no global-load latency in the epilogue's dependency chain
// RES: 0 = the kernel has no residual path at all (leaner epilogue for the qkv / fc1 / gate-up launches), 1 = residual through
// registers (any layout: fp32 output, row-modulo residual = position embedding, one CTA per tile), 2 = TMARES
// A_MN / B_MN: the operand is given TRANSPOSED in memory — A as At [K, M] row-major, W as Wt [K, N] row-major (the contraction index
// is the ROW index) — and is fed to the tensor core as an MN-major operand: TMA brings 64 (k rows) x 64 (m or n columns) boxes into
// 8 KiB slabs, the UMMA descriptor walks them with LBO = slab stride, a K = 16 step advances 16 rows = 2 KiB.  This is what the SFT
// step's dgrad (dX = dY W: W is "Wt") and wgrad (dW = dY^T X: both operands) need; no transposed copies are materialised.
// Tile width BN = 256 columns; a ragged last column block runs the full-width MMA on zero-filled columns.
// Phase trace (tools/micro/gemm_trace.cu compiles this file with -DLP_GEMM_TRACE): lane 0 of the MMA warp and of the first epilogue
// warp of each warpgroup of ONE chosen CTA stamp clock64() at the points below; no code in normal builds.
#ifdef LP_GEMM_TRACE
__device__ unsigned long long* g_gemm_trace = nullptr;   // [3 slots][4096]: slot 0 = MMA issuer, 1 / 2 = epilogue warp 2 / 6
__device__ int g_gemm_trace_blk = 0;
#define G_TRACE_INIT()                                                                                                  \
  unsigned long long* tr_buf = nullptr;                                                                                 \
  int tr_n = 0;                                                                                                         \
  if ((int)blockIdx.x == g_gemm_trace_blk && lane == 0 && g_gemm_trace != nullptr && (warp == 1 || warp == 2 || warp == 6)) \
    tr_buf = g_gemm_trace + (warp == 1 ? 0 : warp == 2 ? 1 : 2) * 4096
#define G_TRACE(id)                                                                                                     \
  do {                                                                                                                  \
    if (tr_buf != nullptr && tr_n < 4094) {                                                                             \
      tr_buf[1 + tr_n] = ((unsigned long long)(id) << 48) | ((unsigned long long)clock64() & 0xffffffffffffULL);        \
      ++tr_n;                                                                                                           \
      tr_buf[0] = (unsigned long long)tr_n;                                                                             \
    }                                                                                                                   \
  } while (0)
#else
#define G_TRACE_INIT() do {} while (0)
#define G_TRACE(id) do {} while (0)
#endif

template <int BN, int STAGES, bool CTA2, bool TP, int RES, bool A_MN = false, bool B_MN = false>
__device__ __forceinline__ void gemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                          const CUtensorMap* segC, const CUtensorMap& tmR, const GemmArgs& args) {
  constexpr bool TMARES = RES == 2;
  using L = GemmSmem<BN, STAGES, CTA2, TMARES ? 3 : 1>;
  constexpr int BMT = CTA2 ? 2 * GEMM_BM : GEMM_BM;  // rows per work unit
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int unit = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256u : 512u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* tpre_bar = tempty_bar + 3;   // stream-K: the END item's accumulator has been loaded back into TMEM
  float* bias_smem = reinterpret_cast<float*>(smem + L::BIAS_OFF);
  float* colsum_smem = reinterpret_cast<float*>(smem + L::COLSUM_OFF);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  G_TRACE_INIT();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!args.out_fp32) tma_prefetch_desc(&tmC);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CTA2 ? 16 : 8);  // one arrive per epilogue warp (of both CTAs)
    }
    mbar_init(tpre_bar, CTA2 ? 16 : 8);
    if (TMARES) {
      tma_prefetch_desc(&tmR);
      for (int i = 0; i < 24; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem + L::RES_BAR_OFF) + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) { tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr_smem, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // barrier init / TMEM allocation / descriptor prefetch above overlap the previous kernel's tail; nothing below may
  // touch global memory before the prerequisite grids have completed
  pdl_launch_dependents();
  pdl_wait();

  const GemmSchedule sch = gemm_schedule(args, unit, n_units);   // this unit's work items (whole tiles + stream-K pieces)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool gathered = false;
      for (int it = 0; it < sch.n_items; ++it) {
        const GemmItem w = gemm_item(args, sch, it);
        int m_blk, n_blk, seg, row_limit;
        tile_coords(args, w.tile, m_blk, n_blk);
        const int a_row = tile_row0(args, m_blk, BMT, seg, row_limit) + (int)cta_rank * GEMM_BM;
        if (TP && args.ag_counter != nullptr && seg != args.tp_rank && !gathered) {
          // rows of other ranks: wait until every CTA's epilogue warps have finished copying them in (all-gather prologue)
          spin_until_gpu(args.ag_counter, gridDim.x);
          fence_proxy_async_all();
          gathered = true;
        }
        // ragged last column block: the MMA runs with N = n_rt (multiple of 32), so a pair splits n_rt, not BN
        // (transposed operands always run full-width MMAs: their 64-column slabs would not line up with a narrower last block; the
        //  columns beyond N are TMA zero fill and are clipped by the stores)
        const int n_rt = (A_MN || B_MN) ? BN : min(BN, ((args.N - n_blk * BN) + 31) & ~31);
        const int b_row = n_blk * BN + (CTA2 ? (int)cta_rank * (n_rt / 2) : 0);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          // (an explicit cp.async.bulk.prefetch.tensor.L2 of the k-block 8 ahead was measured: every CTA of a
          //  row-block group re-requests the same W tile, and the doubled TMA/L2 request rate cost 10-30 %: removed)
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (CTA2) {
            // both CTAs' bytes are counted on the LEADER's full barrier (it alone arrives, expecting 2 stages)
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (A_MN) {      // At [K, M]: two 64-column slabs of this CTA's 128 output rows
              tma_load_2d_pair(sa, &tmA, bar, a_row, kb * GEMM_BK);
              tma_load_2d_pair(sa + 8192, &tmA, bar, a_row + 64, kb * GEMM_BK);
            } else {
              tma_load_2d_pair(sa, &tmA, bar, kb * GEMM_BK, a_row);
            }
            if (B_MN) {      // Wt [K, N]: this CTA's BN / 2 columns in 64-column slabs
#pragma unroll
              for (int sl = 0; sl < BN / 128; ++sl) tma_load_2d_pair(sb + sl * 8192, &tmB, bar, b_row + sl * 64, kb * GEMM_BK);
            } else {
              tma_load_2d_pair(sb, &tmB, bar, kb * GEMM_BK, b_row);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, a_row);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int it = 0; it < sch.n_items; ++it) {
        const GemmItem w = gemm_item(args, sch, it);
        int m_blk_unused, n_blk;
        tile_coords(args, w.tile, m_blk_unused, n_blk);
        // the last column block of a ragged N issues narrower MMAs (N = valid columns rounded up to 32)
        const uint32_t idesc = umma_idesc_bf16(BMT, (A_MN || B_MN) ? BN : min(BN, ((args.N - n_blk * BN) + 31) & ~31), A_MN ? 1 : 0,
                                               B_MN ? 1 : 0);
        G_TRACE(20);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        G_TRACE(21);
        const bool resume = w.kind == GEMM_ITEM_END;   // the accumulator continues another unit's partial sums
        if (resume) mbar_wait(tpre_bar, 0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(a_addr, A_MN ? 8192 : 0, 1024);
          const uint64_t bdesc = umma_desc_sw128(a_addr + L::A_BYTES, B_MN ? 8192 : 0, 1024);
          constexpr uint32_t a_step = A_MN ? 128 : 2, b_step = B_MN ? 128 : 2;   // descriptor units of 16 B per K = 16 step
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {  // +32 B per K=16 step inside the 128 B swizzle row
            const uint32_t acc = (resume || kb > w.kb0 || k > 0) ? 1u : 0u;
            if (CTA2) umma_ss_pair(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, acc);
            else umma_ss(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, acc);
          }
          // smem slot reusable once these MMAs retire (in both CTAs of a pair)
          if (CTA2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CTA2) umma_commit_pair(&tfull_bar[as]); else umma_commit(&tfull_bar[as]);
        G_TRACE(22);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int lg = warp & 3;            // TMEM lane quarter this warp may read
    const int wg = (warp - 2) >> 2;     // warpgroup 0 / 1: first / second part of the tile's columns
    const int et = threadIdx.x - 64;
    int as = 0;
    uint32_t aphase = 0;
    const bool swiglu = args.act == LP_ACT_SWIGLU;
    constexpr int NC = BN / 32;         // 32-column chunks per tile
    constexpr int NP = BN / 64;         // 64-column gate/up pairs per tile
    if (TP && args.ag_counter != nullptr) {
      // ---- all-gather prologue: while the MMA warps work on this rank's own rows, the 256 epilogue threads of every CTA pull
      // the other ranks' rows of A over NVLink (16-byte loads from peer memory) into the local A buffer.
      if (blockIdx.x == 0 && et == 0) {      // this rank's rows are final (stream order): tell the peers
        for (int r = 1; r < args.tp_size; ++r) st_release_sys(args.ag_post[(args.tp_rank + r) % args.tp_size], args.epoch);
      }
      const int vec_per_row = args.ag_cols >> 3;
      const long long total = (long long)args.seg_rows * vec_per_row;
      const long long ldv = args.ag_ld >> 3;
      for (int r = 1; r < args.tp_size; ++r) {
        const int p = (args.tp_rank + r) % args.tp_size;
        if (et == 0) spin_until_sys(args.ag_wait[p], args.epoch);
        epi_named_barrier();
        const uint4* src = reinterpret_cast<const uint4*>(args.ag_src[p]) + (long long)p * args.seg_rows * ldv;
        uint4* dst = reinterpret_cast<uint4*>(args.ag_dst) + (long long)p * args.seg_rows * ldv;
        const long long stride = (long long)gridDim.x * GEMM_EPI_THREADS;
        long long i = (long long)blockIdx.x * GEMM_EPI_THREADS + et;
        for (; i + 3 * stride < total; i += 4 * stride) {       // four independent 16-byte loads in flight per thread
          long long o[4];
          uint4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const long long j = i + u * stride;
            o[u] = (j / vec_per_row) * ldv + (j % vec_per_row);
            v[u] = src[o[u]];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) dst[o[u]] = v[u];
        }
        for (; i < total; i += stride) {
          const long long o = (i / vec_per_row) * ldv + (i % vec_per_row);
          dst[o] = src[o];
        }
      }
      __threadfence();
      fence_proxy_async_all();     // the copied rows are read by TMA (async proxy)
      epi_named_barrier();
      if (et == 0) atomicAdd(args.ag_counter, 1u);
    }
    const int sk_c0 = wg == 0 ? 0 : (NC + 1) / 2, sk_c1 = wg == 0 ? (NC + 1) / 2 : NC;   // chunk split of the stream-K dump / reload
    const int sk_slot = unit * (CTA2 ? 2 : 1) + (int)cta_rank;
    bool sk_reloaded = false;
    int stat_row = -1;
    float rs = 1.f, nmu = 0.f;
    // bias / layer-norm column sums of a tile sit in shared memory (double-buffered with the accumulator stage); the values of item
    // it + 1 are fetched into a register while item it is processed, so their global-load latency is off the epilogue's critical path
    const bool col_thread = et < BN;
    auto fetch_cols = [&](const GemmItem& wi, float& bv, float& cv) {
      int mb, nb;
      tile_coords(args, wi.tile, mb, nb);
      const int col = nb * BN + et;
      const bool ok = col_thread && col < args.N;
      bv = (ok && args.bias) ? args.bias[col] : 0.f;
      cv = (ok && args.norm_mode == LP_NORM_LAYER_) ? args.colsum[col] : 0.f;
    };
    if (sch.n_items > 0 && col_thread) {
      float bv, cv;
      fetch_cols(gemm_item(args, sch, 0), bv, cv);
      bias_smem[et] = bv;
      colsum_smem[et] = cv;
    }
    // TMARES: residual ring of this warp (3 x 2 KiB) and its barriers; rg counts the chunks this warp has sent through it
    uint8_t* ring = smem + L::STAGE_OUT_OFF + (warp - 2) * (TMARES ? 3 : 1) * 2048;
    uint64_t* rbar = reinterpret_cast<uint64_t*>(smem + L::RES_BAR_OFF) + (warp - 2) * 3;
    uint32_t rg = 0;
    auto sk_reload = [&]() {
        // ---- stream-K END item (always the last one): the unit before this one computed k-blocks [0, kb0) of the tile and dumped its
        // accumulator; load it into the TMEM stage the END item will use (free: its previous user's epilogue is behind us) while
        // the MMA warp still works on the item before it
        sk_reloaded = true;
        if (et == 0) {
          spin_until_gpu(args.sk_flags + sk_slot - (CTA2 ? 2 : 1), 1u);
          st_relaxed_gpu(args.sk_flags + sk_slot - (CTA2 ? 2 : 1), 0u);   // consumed: ready for the next launch
        }
        epi_named_barrier();
        const float* src = args.sk_ws + (size_t)(sk_slot - (CTA2 ? 2 : 1)) * (BN * GEMM_BM) + (size_t)(lg * 32 + lane) * 4;
        const uint32_t tdst = tmem_base + ((uint32_t)(lg * 32) << 16) + ((sch.n_items - 1) & 1) * BN;
#pragma unroll 1
        for (int c = sk_c0; c < sk_c1; ++c) {
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 u = ld_cg_u4(src + (size_t)(c * 8 + j) * (GEMM_BM * 4));
            r[4 * j] = u.x; r[4 * j + 1] = u.y; r[4 * j + 2] = u.z; r[4 * j + 3] = u.w;
          }
          LP_TMEM_ST_32(r, tdst + c * 32);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2 && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(tpre_bar), 0));
          else mbar_arrive(tpre_bar);
        }
          };
    for (int it = 0; it < sch.n_items; ++it) {
      const GemmItem w = gemm_item(args, sch, it);
      if (sch.has_end && !sk_reloaded && it + 2 >= sch.n_items && w.kind != GEMM_ITEM_START) sk_reload();
      if (w.kind == GEMM_ITEM_START) {
        // ---- stream-K START item (after the data-parallel tiles): k-blocks [0, kb1) only; the raw fp32 accumulator goes to this unit's slot
        // of the workspace ([chunk][4-column group][row][4]: every warp store covers 512 contiguous bytes) for the next unit
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        float* dst = args.sk_ws + (size_t)sk_slot * (BN * GEMM_BM) + (size_t)(lg * 32 + lane) * 4;
        const uint32_t tsrc = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN;
#pragma unroll 1
        for (int c = sk_c0; c < sk_c1; ++c) {
          uint32_t r[32];
          LP_TMEM_LD_32(r, tsrc + c * 32);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_cg_u4(dst + (size_t)(c * 8 + j) * (GEMM_BM * 4), make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2 && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
          else mbar_arrive(&tempty_bar[as]);
        }
        __threadfence();
        epi_named_barrier();
        if (et == 0) st_release_gpu(args.sk_flags + sk_slot, 1u);
        as ^= 1;
        if (as == 0) aphase ^= 1;
        if (it + 1 < sch.n_items && col_thread) {   // bias / column sums of the next item (every warp has left item it - 1: barrier above)
          float bv, cv;
          fetch_cols(gemm_item(args, sch, it + 1), bv, cv);
          bias_smem[as * BN + et] = bv;
          colsum_smem[as * BN + et] = cv;
        }
        if (sch.has_end && !sk_reloaded && it + 2 >= sch.n_items) sk_reload();   // [.., START, END]: own dump first (no chain of waits)
        continue;
      }
      int m_blk, n_blk, seg, row_limit;
      tile_coords(args, w.tile, m_blk, n_blk);
      const int m0 = tile_row0(args, m_blk, BMT, seg, row_limit) + (int)cta_rank * GEMM_BM, n0 = n_blk * BN;
      // tensor-parallel launches store through the segment's own map (rows clipped at the segment end; a reduce-scatter
      // launch's maps point at the owner rank's staging memory), with row coordinates relative to the segment
      const CUtensorMap* tmc = TP ? &segC[seg] : &tmC;
      const int st_shift = TP ? seg * args.seg_rows : 0;
      float* bs = bias_smem + as * BN;
      float* cs = colsum_smem + as * BN;
      const int row = m0 + lg * 32 + lane;
      const bool row_ok = row < row_limit;
      // row statistics of the A operand are fetched while the tile's MMAs are still running (and kept while the unit stays in
      // the same row block, which is every tile when A is L2-resident and the raster walks the column blocks of one row block)
      if (row != stat_row) {
        row_norm_coeffs(args, row, row_ok, rs, nmu);
        stat_row = row;
      }
      const int row_base = m0 + lg * 32;
      const int c_begin_n = wg == 0 ? 0 : (NC + 1) / 2, c_end_n = wg == 0 ? (NC + 1) / 2 : NC;   // this warp's 32-column chunks
      auto res_issue = [&](uint32_t g, int col0) {      // lane 0: residual tile of chunk g -> ring buffer g % 3
        mbar_arrive_expect_tx(&rbar[g % 3], 2048);
        tma_load_2d(ring + (g % 3) * 2048, &tmR, &rbar[g % 3], col0, row_base);
      };
      if (TMARES && lane == 0) {
        // the first two residual chunks of the tile are requested before the accumulator is even complete; buffers g % 3 and
        // (g + 1) % 3 were last read by the stores of chunks g - 3 and g - 2: at most the newest store may still be reading
        tma_store_wait_read<1>();
        if (n0 + c_begin_n * 32 < args.N) res_issue(rg, n0 + c_begin_n * 32);
        if (c_begin_n + 1 < c_end_n && n0 + (c_begin_n + 1) * 32 < args.N) res_issue(rg + 1, n0 + (c_begin_n + 1) * 32);
      }
      epi_named_barrier();   // bias / column sums of this item (written during the previous item) are visible
      float nbv = 0.f, ncv = 0.f;
      const bool prefetch_cols = it + 1 < sch.n_items;
      if (prefetch_cols) fetch_cols(gemm_item(args, sch, it + 1), nbv, ncv);
      G_TRACE(30);
      mbar_wait_relaxed(&tfull_bar[as], aphase, args.epi_sleep_ns);
      G_TRACE(31);
      tc_fence_after();
      const int rrow = args.res_row_mod > 0 ? row % args.res_row_mod : row;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN;
      uint8_t* stage = ring;
      const bool has_res = RES == 1 && args.residual != nullptr;
      float2 ostat = make_float2(0.f, 0.f);
      float2* ostat_p = args.stats_out != nullptr ? &ostat : nullptr;
      if (!swiglu) {
        const int c_begin = wg == 0 ? 0 : (NC + 1) / 2, c_end = wg == 0 ? (NC + 1) / 2 : NC;
        // RoPE tiles (BN = 256): each warpgroup owns one 128-column head = chunks c_begin .. c_begin + 3; column i pairs with i + 64
        const bool rope_tile = BN == 256 && args.rope_heads > 0 && ((n0 >> 7) + wg) < args.rope_heads;
        if (rope_tile) {
#pragma unroll 1
          for (int p = 0; p < 2; ++p) {
            uint32_t rl[32], rh[32];
            LP_TMEM_LD_32(rl, taddr0 + (c_begin + p) * 32);
            LP_TMEM_LD_32(rh, taddr0 + (c_begin + p + 2) * 32);
            tmem_wait_ld();
            float lo[32], hi[32];
            const float* cp = args.rope_cos + (long long)(row_ok ? row : 0) * 64 + p * 32;
            const float* sp = args.rope_sin + (long long)(row_ok ? row : 0) * 64 + p * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 c4 = *reinterpret_cast<const float4*>(cp + j);
              const float4 s4 = *reinterpret_cast<const float4*>(sp + j);
#define LP_ROPE_PAIR(T, CV, SV)                                                                                       \
  {                                                                                                                    \
    const float a = rs * __uint_as_float(rl[j + T]), b = rs * __uint_as_float(rh[j + T]);                              \
    lo[j + T] = a * (CV) - b * (SV);                                                                                   \
    hi[j + T] = b * (CV) + a * (SV);                                                                                   \
  }
              LP_ROPE_PAIR(0, c4.x, s4.x) LP_ROPE_PAIR(1, c4.y, s4.y) LP_ROPE_PAIR(2, c4.z, s4.z) LP_ROPE_PAIR(3, c4.w, s4.w)
#undef LP_ROPE_PAIR
            }
            staged_store_chunk(args, tmc, stage, lo, nullptr, row_base - st_shift, n0 + (c_begin + p) * 32, lane, ostat_p);
            staged_store_chunk(args, tmc, stage, hi, nullptr, row_base - st_shift, n0 + (c_begin + p + 2) * 32, lane, ostat_p);
          }
        } else if (TMARES) {
          uint32_t done = 0;
#pragma unroll 1
          for (int c = c_begin; c < c_end; ++c) {
            const int col0 = n0 + c * 32;
            if (col0 >= args.N) break;
            const uint32_t g = rg + done;
            if (c > c_begin && lane == 0) {
              // chunk c + 1's residual goes to the buffer chunk c - 2 was stored from: all but the newest store must have been read
              tma_store_wait_read<1>();
              if (c + 1 < c_end && col0 + 32 < args.N) res_issue(g + 1, col0 + 32);
            }
            G_TRACE(35);
            uint32_t r[32];
            LP_TMEM_LD_32(r, taddr0 + c * 32);
            tmem_wait_ld();
            G_TRACE(32);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            apply_row_norm(args, v, cs + c * 32, rs, nmu);
            apply_bias_act(args, v, bs + c * 32);
            G_TRACE(33);
            mbar_wait(&rbar[g % 3], (g / 3) & 1);
            G_TRACE(36);
            ring_add_store_chunk(tmc, ring + (g % 3) * 2048, v, row_base, col0, lane, ostat_p);
            G_TRACE(34);
            ++done;
          }
          rg += done;
        } else {
        ResidualRegs rr;
        if (has_res && !args.out_fp32 && n0 + c_begin * 32 < args.N)
          residual_fetch(args, rr, row_base, n0 + c_begin * 32, args.N, lane);
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= args.N) break;
          uint32_t r[32];
          LP_TMEM_LD_32(r, taddr0 + c * 32);
          tmem_wait_ld();
          G_TRACE(32);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          apply_row_norm(args, v, cs + c * 32, rs, nmu);
          if (args.out_fp32) {
            epilogue_store_chunk(args, v, bs + c * 32, row, rrow, row_ok, col0, args.N);
          } else {
            apply_bias_act(args, v, bs + c * 32);
            ResidualRegs cur = rr;
            if (has_res && c + 1 < c_end && col0 + 32 < args.N)  // prefetch the next chunk's residual
              residual_fetch(args, rr, row_base, col0 + 32, args.N, lane);
            G_TRACE(33);
            staged_store_chunk(args, tmc, stage, v, has_res ? &cur : nullptr, row_base - st_shift, col0, lane, ostat_p);
            G_TRACE(34);
          }
        }
        }
      } else {
        // columns are interleaved in 32-wide granules [gate | up | gate | up ...]; output has N/2 columns
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
            float g = rs * __uint_as_float(rg[j]), u = rs * __uint_as_float(ru[j]);   // rs = 1 without a fused RMSNorm
            if (args.bias) { g += bs[c * 64 + j]; u += bs[c * 64 + 32 + j]; }
            v[j] = silu_f(g) * u;
          }
          ResidualRegs cur = rr;
          if (has_res && c + 1 < c_end && col0 + 64 < args.N)
            residual_fetch(args, rr, row_base, (col0 + 64) / 2, args.N / 2, lane);
          staged_store_chunk(args, tmc, stage, v, has_res ? &cur : nullptr, row_base - st_shift, col0 / 2, lane);
        }
      }
      if (args.stats_out != nullptr && row_ok)
        args.stats_out[(long long)row * (2 * args.num_n_blk) + 2 * n_blk + wg] = ostat;
      if (prefetch_cols && col_thread) {   // the other buffer: last read during item it - 1, which every warp has left (barrier above)
        bias_smem[(as ^ 1) * BN + et] = nbv;
        colsum_smem[(as ^ 1) * BN + et] = ncv;
      }
      G_TRACE(37);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2 && !leader) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
        else mbar_arrive(&tempty_bar[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();  // all bulk stores of this warp have landed before the CTA may exit
  }

  tc_fence_before();
  __syncwarp();  // role branches above leave warps 0/1 diverged; barrier.cluster.*.aligned needs converged warps
  if (CTA2) cluster_sync_all(); else __syncthreads();  // the leader's MMAs read the peer's smem: leave together
  if (warp == 1) {
    __syncwarp();
    if (CTA2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (TP && threadIdx.x == 0) {
    if (args.rs_counter != nullptr) {
      // reduce-scatter: every epilogue warp has waited for its bulk stores (partial tiles in the owners' memory); the last CTA to
      // get here tells every owner that this rank's partial sums have landed
      fence_proxy_async_all();
      __threadfence_system();
      if (atomicAdd(args.rs_counter, 1u) == gridDim.x - 1) {
        for (int r = 0; r < args.tp_size; ++r) st_release_sys(args.rs_post[r], args.epoch);
        if (args.recycle != nullptr) *args.recycle = 0u;
      }
    } else if (blockIdx.x == 0 && args.recycle != nullptr) {
      *args.recycle = 0u;
    }
  }
}

template <int BN, int STAGES, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmArgs args) {
  gemm_body<BN, STAGES, CTA2, false, 0>(tmA, tmB, tmC, nullptr, tmC, args);
}

// transposed-operand launches (SFT dgrad / wgrad): pair mode, no residual
template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_mn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmArgs args) {
  static_assert(!A_MN || B_MN || true, "");
  gemm_body<BN, STAGES, true, false, 0, A_MN, B_MN>(tmA, tmB, tmC, nullptr, tmC, args);
}

// residual through registers (see RES above)
template <int BN, int STAGES, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_lres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const GemmArgs args) {
  gemm_body<BN, STAGES, CTA2, false, 1>(tmA, tmB, tmC, nullptr, tmC, args);
}

// residual launches of the layer loops: residual tile by TMA (tmR; equal to tmC for the in-place residual stream)
template <int BN, int STAGES, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_res_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmArgs args) {
  gemm_body<BN, STAGES, CTA2, false, 2>(tmA, tmB, tmC, nullptr, tmR, args);
}

template <int BN, int STAGES, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GemmSegMaps segs, const GemmArgs args) {
  gemm_body<BN, STAGES, CTA2, true, 0>(tmA, tmB, segs.m[0], segs.m, segs.m[0], args);
}

template <int BN, int STAGES, bool CTA2>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmArgs& args,
                       cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, CTA2>;
  auto kern = args.residual != nullptr ? gemm_bf16_lres_kernel<BN, STAGES, CTA2> : gemm_bf16_kernel<BN, STAGES, CTA2>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    LP_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_lres_kernel<BN, STAGES, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = args.num_m_blk * args.num_n_blk;
  if (!CTA2) {
    const int grid = tiles < sm_count() ? tiles : sm_count();
    LP_CHECK_CUDA(launch_ex(kern, dim3(grid), dim3(GEMM_THREADS), L::TOTAL, stream, 1, prefill_pdl(), tmA, tmB, tmC, args));
  } else {
    const int pairs = sm_count() / 2;
    LP_CHECK_CUDA(launch_ex(kern, dim3(2 * (tiles < pairs ? tiles : pairs)), dim3(GEMM_THREADS), L::TOTAL, stream, 2,
                            prefill_pdl(), tmA, tmB, tmC, args));
  }
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// residual launches in pair mode: same structure, residual tile by TMA into a per-warp ring (gemm_bf16_res_kernel)
template <int BN, int STAGES>
static int launch_gemm_res(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                           const GemmArgs& args, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, true, 3>;
  auto kern = gemm_bf16_res_kernel<BN, STAGES, true>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = args.num_m_blk * args.num_n_blk;
  const int pairs = sm_count() / 2;
  LP_CHECK_CUDA(launch_ex(kern, dim3(2 * (tiles < pairs ? tiles : pairs)), dim3(GEMM_THREADS), L::TOTAL, stream, 2, prefill_pdl(), tmA,
                          tmB, tmC, tmR, args));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// transposed-operand launches (pair mode, 256-wide tiles): (A_MN, B_MN) = (0, 1) dgrad, (1, 1) wgrad, (1, 0) for completeness
template <bool A_MN, bool B_MN_unused>
static int launch_gemm_mn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmArgs& args,
                          cudaStream_t stream, bool b_t) {
  using L = GemmSmem<256, 6, true>;
  auto kern = b_t ? gemm_bf16_mn_kernel<256, 6, A_MN, true> : gemm_bf16_mn_kernel<256, 6, A_MN, false>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_mn_kernel<256, 6, A_MN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    LP_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_mn_kernel<256, 6, A_MN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = args.num_m_blk * args.num_n_blk;
  const int pairs = sm_count() / 2;
  LP_CHECK_CUDA(launch_ex(kern, dim3(2 * (tiles < pairs ? tiles : pairs)), dim3(GEMM_THREADS), L::TOTAL, stream, 2, prefill_pdl(), tmA,
                          tmB, tmC, args));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int g_gemm_tma_residual = 1;  // pair-mode residual launches fetch the residual tile by TMA (0 = register path, for A/B)
static unsigned g_gemm_epi_sleep_ns = 0;   // lp_gemm_set_epilogue_sleep: nanosleep between probes of the accumulator-full barrier
static int g_gemm_raster_balance = 1;  // row-major raster when W fits L2 and the last column block is ragged (lp_gemm_set_raster_balance)
static int g_gemm_pair_mode = 1;  // 0 = one CTA per tile, 1 = CTA pairs where the problem is large enough (default)
// stream-K tail (GemmArgs::sk_*): caller-provided workspace (flags in the first 4 KiB, fp32 partial accumulators after them)
static void* g_sk_ws = nullptr;
static size_t g_sk_bytes = 0;
static int g_sk_enable = 1;   // 0 = off, 1 = stream-K tail only (default), 2 = tail + pure stream-K for norm-consuming L2-resident launches (measured: no gain, opt-in)
constexpr size_t SK_FLAG_BYTES = 4096;

}  // namespace lp

using namespace lp;

namespace lp {
// tile width for a problem: wave-quantisation heuristic — minimise ceil(tiles / units) * (BN + per-tile overhead); the narrow
// tiles are shared-memory-bandwidth bound on one CTA (A 4 KiB + B BN*32 B per K=16 step), hence their penalty
static int pick_block_n(int M, int N, int act, bool pair, int sms) {
  const int units = pair ? sms / 2 : sms;
  const int bm_unit = pair ? 2 * GEMM_BM : GEMM_BM;
  const long long mb = (M + bm_unit - 1) / bm_unit;
  const int cand[5] = {256, 224, 192, 160, 128};
  const double pen[5] = {1.0, 1.15, 1.10, 1.15, 1.30};  // measured: wider tiles amortise the epilogue better
  double best = 1e30;
  int block_n = 256;
  for (int i = 0; i < 5; ++i) {
    if (act == LP_ACT_SWIGLU && cand[i] % 64 != 0) continue;
    if (pair && cand[i] == 160) continue;  // pair kernels are instantiated for 128 / 192 / 224 / 256
    const long long tiles = mb * ((N + cand[i] - 1) / cand[i]);
    const double cost = (double)((tiles + units - 1) / units) * (cand[i] + 12) * pen[i];
    if (cost < best) { best = cost; block_n = cand[i]; }
  }
  return block_n;
}
}  // namespace lp

extern "C" int lp_gemm_pick_block_n(int M, int N, int act) {
  return pick_block_n(M, N, act, g_gemm_pair_mode != 0 && M > GEMM_BM, sm_count());
}

extern "C" int lp_gemm_bf16_ex(const lp_gemm_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(d != nullptr, "lp_gemm_bf16_ex: null descriptor");
  const void* A = d->A; const void* W = d->W; void* C = d->C;
  const long long lda = d->lda, ldw = d->ldw, ldc = d->ldc, ldr = d->ldr;
  const int M = d->M, N = d->N, K = d->K, act = d->act, out_fp32 = d->out_fp32, res_row_mod = d->res_row_mod;
  const float* bias = d->bias; const void* residual = d->residual;
  int block_n = d->block_n;
  LP_REQUIRE(M > 0 && N > 0 && K > 0, "lp_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  LP_REQUIRE(N % 8 == 0, "lp_gemm_bf16: N=%d must be a multiple of 8", N);
  LP_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "lp_gemm_bf16: lda/ldw must be multiples of 8 elements (TMA 16 B rule)");
  LP_REQUIRE(ldc % (out_fp32 ? 4 : 8) == 0, "lp_gemm_bf16: ldc=%lld breaks 16 B store alignment", ldc);
  LP_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(C) & 15) == 0,
             "lp_gemm_bf16: A/W/C must be 16-byte aligned");
  LP_REQUIRE(residual == nullptr || (ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
             "lp_gemm_bf16: residual must be 16-byte aligned with ldr %% 8 == 0");
  LP_REQUIRE(act >= LP_ACT_NONE && act <= LP_ACT_SWIGLU, "lp_gemm_bf16: bad act %d", act);
  LP_REQUIRE(act != LP_ACT_SWIGLU || N % 64 == 0, "lp_gemm_bf16: SwiGLU needs N %% 64 == 0 (32-wide gate/up granules)");
  LP_REQUIRE(d->norm_mode >= LP_NORM_NONE && d->norm_mode <= LP_NORM_LAYER, "lp_gemm_bf16_ex: bad norm_mode %d", d->norm_mode);
  LP_REQUIRE(d->norm_mode == LP_NORM_NONE || (d->stats_in != nullptr && d->stats_in_parts > 0 &&
                                              (reinterpret_cast<uintptr_t>(d->stats_in) & 7) == 0),
             "lp_gemm_bf16_ex: fused normalisation needs the producer's row statistics");
  LP_REQUIRE(d->norm_mode != LP_NORM_LAYER || (d->colsum != nullptr && act != LP_ACT_SWIGLU),
             "lp_gemm_bf16_ex: layer-norm mode needs the column sums of the folded weight (and no SwiGLU)");
  LP_REQUIRE(d->stats_out == nullptr || (!out_fp32 && act != LP_ACT_SWIGLU && (reinterpret_cast<uintptr_t>(d->stats_out) & 7) == 0),
             "lp_gemm_bf16_ex: row statistics are produced by bf16-output, non-SwiGLU epilogues");
  LP_REQUIRE(d->rope_heads == 0 || (d->rope_cos && d->rope_sin && !out_fp32 && act == LP_ACT_NONE && bias == nullptr &&
                                    residual == nullptr && N % 128 == 0 && d->rope_heads * 128 <= N &&
                                    d->norm_mode != LP_NORM_LAYER && (block_n == 0 || block_n == 256)),
             "lp_gemm_bf16_ex: fused RoPE needs 128-column heads, bf16 output, no bias / activation / residual, block_n 256");
  const int sms = sm_count();
  // CTA pairs (256-row tiles) when asked for and the problem has at least one tile per pair
  bool pair = g_gemm_pair_mode != 0 && M > GEMM_BM;
  const int units = pair ? sms / 2 : sms;
  const int bm_unit = pair ? 2 * GEMM_BM : GEMM_BM;
  const long long mb = (M + bm_unit - 1) / bm_unit;
  if (d->rope_heads > 0 || d->a_t || d->b_t) block_n = 256;
  if (block_n == 0) block_n = pick_block_n(M, N, act, pair, sms);
  LP_REQUIRE(block_n == 128 || block_n == 160 || block_n == 192 || block_n == 224 || block_n == 256,
             "lp_gemm_bf16: block_n must be 0, 128, 160, 192, 224 or 256");
  LP_REQUIRE(act != LP_ACT_SWIGLU || block_n % 64 == 0, "lp_gemm_bf16: SwiGLU needs block_n %% 64 == 0");

  GemmArgs args = {};
  args.M = M; args.N = N; args.K = K;
  args.C = C; args.ldc = ldc; args.out_fp32 = out_fp32;
  args.bias = bias;
  args.residual = reinterpret_cast<const bf16*>(residual);
  args.ldr = ldr; args.res_row_mod = res_row_mod;
  args.act = act;
  args.epi_sleep_ns = g_gemm_epi_sleep_ns;
  args.norm_mode = d->norm_mode;
  args.stats_in = reinterpret_cast<const float2*>(d->stats_in);
  args.stats_in_parts = d->stats_in_parts;
  args.colsum = d->colsum;
  args.norm_eps = d->norm_eps;
  args.norm_inv_k = 1.0f / (float)K;
  args.stats_out = reinterpret_cast<float2*>(d->stats_out);
  args.rope_cos = d->rope_cos; args.rope_sin = d->rope_sin; args.rope_heads = d->rope_heads;
  if (pair && (block_n == 160 || mb * ((N + block_n - 1) / block_n) < units)) pair = false;
  args.num_m_blk = pair ? (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM) : (M + GEMM_BM - 1) / GEMM_BM;
  args.num_n_blk = (N + block_n - 1) / block_n;
  args.num_k_blk = (K + GEMM_BK - 1) / GEMM_BK;
  LP_REQUIRE(d->stats_out == nullptr || d->stats_out_parts == 2 * args.num_n_blk,
             "lp_gemm_bf16_ex: stats_out_parts=%d, this launch writes %d per row (2 x column blocks of %d)", d->stats_out_parts,
             2 * args.num_n_blk, block_n);
  {
    // L2-aware rasterisation: keep the operand that fits L2 resident, stream the other from HBM once
    const double l2_budget = 56e6;
    const double a_bytes = 2.0 * M * K, w_bytes = 2.0 * N * K;
    if (a_bytes <= l2_budget) args.group_m = args.num_m_blk;                 // A resident, W tiles read once
    else if (w_bytes <= l2_budget) {
      args.group_m = max(1, sms / args.num_n_blk);  // W resident, A read once
      // A ragged last column block (N = 1152: 4 x 256 + 128) makes its tiles cheaper; units take tiles round-robin over the raster
      // index, and with groups of ~29 row blocks the 74 half-width tiles of the ViT out-proj / fc2 land on 35 of the 74 CTA pairs —
      // the others run five full tiles while the mean is 4.5 (-11 %).  Row-major order (groups of one row block) hands every unit
      // each column block equally often and still reads A once (a row block's tiles run concurrently on neighbouring units).
      if (g_gemm_raster_balance && N % block_n != 0) args.group_m = 1;
    }
    else {  // neither fits: keep a group's A rows (~half of the 126 MB L2) resident while W streams once per group
      const double a_blk_bytes = 2.0 * bm_unit * K;
      args.group_m = (int)(l2_budget / a_blk_bytes);
      if (args.group_m < 1) args.group_m = 1;
    }
    if (args.group_m > args.num_m_blk) args.group_m = args.num_m_blk;
    args.prefetch_a = a_bytes > l2_budget ? 1 : 0;  // activations that cannot stay L2-resident are prefetched too
  }

  CUtensorMap tmA, tmB, tmC;
  int rc = 0;
  if (!out_fp32) {  // output tile map: 32 x 32 bf16 boxes with 64-byte swizzle (see staged_store_chunk)
    rc = make_tmap_2d(&tmC, C, 2, (uint64_t)(act == LP_ACT_SWIGLU ? N / 2 : N), (uint64_t)M, (uint64_t)ldc, 32, 32, 64);
    if (rc) return rc;
  } else {
    tmC = CUtensorMap{};
  }
  const bool a_t = d->a_t != 0, b_t = d->b_t != 0;
  if (a_t || b_t) {
    LP_REQUIRE(pair && block_n == 256 && (!a_t || M % 8 == 0),
               "lp_gemm_bf16_ex: transposed operands need the CTA-pair mode (M=%d N=%d: at least one 256-row tile per SM pair) and "
               "block_n 256", M, N);
    LP_REQUIRE(residual == nullptr && act == LP_ACT_NONE && d->norm_mode == LP_NORM_NONE && d->stats_out == nullptr && d->rope_heads == 0,
               "lp_gemm_bf16_ex: transposed operands take the plain epilogue (bias / fp32 output only)");
  }
  // K-major operand: [rows, K] with 64-wide k boxes; transposed (MN-major) operand: [K, rows] with 64 (rows) x 64 (k) boxes
  rc = a_t ? make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, GEMM_BK)
           : make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, GEMM_BM);
  if (rc) return rc;
  rc = b_t ? make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 64, GEMM_BK)
           : make_tmap_bf16_2d(&tmB, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, GEMM_BK, (uint32_t)(pair ? block_n / 2 : block_n));
  if (rc) return rc;
  ProfScope prof(1, M, N, K, stream);
  note_launch();
  const bool plain = d->norm_mode == LP_NORM_NONE && d->stats_out == nullptr && d->rope_heads == 0;
  if (plain && pair && g_gemm_pair_mode == 2 && (block_n == 256 || block_n == 192 || block_n == 128) && args.num_n_blk >= 2) {
    // two pairs per 4-CTA cluster on column-adjacent tiles, A multicast between them (lp_gemm4.cu): work units are
    // (row block, column-block pair); A is requested in 64-row boxes
    GemmArgs a4 = args;
    a4.num_n_blk = (args.num_n_blk + 1) / 2;
    CUtensorMap tmAh;
    rc = make_tmap_bf16_2d(&tmAh, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, GEMM_BM / 2);
    if (rc) return rc;
    return launch_gemm4(tmAh, tmB, tmC, a4, block_n, stream);
  }
  if (g_sk_enable && g_sk_ws != nullptr && M >= 1024) {
    // stream-K tail: when the last wave would be mostly idle, the last 1 + fraction waves of tiles are cut along K so that every
    // unit gets the same number of k-blocks (gemm_schedule); bit-identical results (the accumulator is handed over, not re-added)
    const int T = args.num_m_blk * args.num_n_blk;
    const int n_units = pair ? sms / 2 : sms, per_unit = pair ? 2 : 1;
    if (T > n_units && T % n_units != 0 && args.num_k_blk >= 8) {
      // measured on the B200 (profiles/r02i_*): worth it when the last wave is mostly empty (S = 4796 o / down: 4.11 waves, +6 ..
      // +10 %; qkv 6.16: +1 %); with a half-full last wave (S = 2937: 2.59 waves) the staggered k ranges of the stream-K phase
      // cost more L2 locality than the idle units did (-4 %), so those keep the plain schedule
      const double frac = (double)T / n_units - (double)(T / n_units);
      const size_t need = SK_FLAG_BYTES + (size_t)n_units * per_unit * GEMM_BM * block_n * sizeof(float);
      if (frac <= 0.40 && need <= g_sk_bytes) {
        args.dp_tiles = (T / n_units - 1) * n_units;
        args.sk_tiles = T - args.dp_tiles;
        args.sk_flags = reinterpret_cast<unsigned*>(g_sk_ws);
        args.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_sk_ws) + SK_FLAG_BYTES);
      }
    }
    // Norm-consuming launches whose operands are L2-resident (the ViT's qkv / fc1 at any crop count): PURE stream-K over a
    // column-block-fastest raster — every unit owns one contiguous range of the (row block, column block, k) space, so it stays in
    // one or two row blocks (the row statistics of the fused LayerNorm are fetched once instead of per tile) and every unit ends
    // at the same time whatever M is (17 crops = 45 row blocks on 74 pairs would otherwise run 9.1 waves as 10).  With M a
    // multiple of the unit count (28 crops = 74 row blocks) this is exactly the plain schedule.  Measured on the B200
    // (profiles/r02o_ab_pure_streamk.jsonl): cfg3 80.3 -> 81.0 ms, cfg2 51.4 -> 51.7 ms per step — no gain (the column-block-fastest
    // raster shares each W tile between fewer concurrent units), so it stays opt-in (lp_gemm_set_streamk(2)).
    const size_t need_pure = SK_FLAG_BYTES + (size_t)n_units * per_unit * GEMM_BM * block_n * sizeof(float);
    if (g_sk_enable == 2 && d->norm_mode != LP_NORM_NONE && T >= n_units && args.num_k_blk >= 8 && need_pure <= g_sk_bytes &&
        2.0 * M * K <= 56e6 && 2.0 * N * K <= 56e6) {
      args.group_m = 1;
      args.dp_tiles = 0;
      args.sk_tiles = T;
      args.sk_flags = reinterpret_cast<unsigned*>(g_sk_ws);
      args.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_sk_ws) + SK_FLAG_BYTES);
    }
  }
  if (a_t || b_t) {
    if (a_t) return launch_gemm_mn<true, true>(tmA, tmB, tmC, args, stream, b_t);
    return launch_gemm_mn<false, true>(tmA, tmB, tmC, args, stream, b_t);
  }
  if (pair && g_gemm_tma_residual && residual != nullptr && res_row_mod == 0 && !out_fp32 && act != LP_ACT_SWIGLU &&
      d->rope_heads == 0) {
    CUtensorMap tmR;
    rc = make_tmap_2d(&tmR, residual, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldr, 32, 32, 64);
    if (rc) return rc;
    switch (block_n) {
      case 256: return launch_gemm_res<256, 5>(tmA, tmB, tmC, tmR, args, stream);
      case 224: return launch_gemm_res<224, 5>(tmA, tmB, tmC, tmR, args, stream);
      case 192: return launch_gemm_res<192, 6>(tmA, tmB, tmC, tmR, args, stream);
      default: return launch_gemm_res<128, 7>(tmA, tmB, tmC, tmR, args, stream);
    }
  }
  if (pair) {
    switch (block_n) {
      case 256: return launch_gemm<256, 6, true>(tmA, tmB, tmC, args, stream);
      case 224: return launch_gemm<224, 6, true>(tmA, tmB, tmC, args, stream);
      case 192: return launch_gemm<192, 7, true>(tmA, tmB, tmC, args, stream);
      default: return launch_gemm<128, 8, true>(tmA, tmB, tmC, args, stream);
    }
  }
  switch (block_n) {
    case 256: return launch_gemm<256, 4, false>(tmA, tmB, tmC, args, stream);
    case 224: return launch_gemm<224, 4, false>(tmA, tmB, tmC, args, stream);
    case 192: return launch_gemm<192, 5, false>(tmA, tmB, tmC, args, stream);
    case 160: return launch_gemm<160, 5, false>(tmA, tmB, tmC, args, stream);
    default: return launch_gemm<128, 6, false>(tmA, tmB, tmC, args, stream);
  }
}

extern "C" int lp_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M,
                            int N, int K, const float* bias, const void* residual, long long ldr, int res_row_mod,
                            int act, int out_fp32, int block_n, void* stream_) {
  lp_gemm_desc d = {};
  d.A = A; d.lda = lda; d.W = W; d.ldw = ldw; d.C = C; d.ldc = ldc;
  d.M = M; d.N = N; d.K = K;
  d.bias = bias; d.residual = residual; d.ldr = ldr; d.res_row_mod = res_row_mod;
  d.act = act; d.out_fp32 = out_fp32; d.block_n = block_n;
  return lp_gemm_bf16_ex(&d, stream_);
}

// ---- tensor-parallel launches: all-gather -> GEMM and GEMM -> reduce-scatter in ONE kernel over NVLink peer memory
extern "C" int lp_gemm_bf16_tp(const lp_gemm_desc* d, const lp_tp_desc* t, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(d && t, "lp_gemm_bf16_tp: null descriptor");
  LP_REQUIRE(t->tp_size >= 2 && t->tp_size <= 8 && t->tp_rank >= 0 && t->tp_rank < t->tp_size, "lp_gemm_bf16_tp: tp_size=%d rank=%d",
             t->tp_size, t->tp_rank);
  LP_REQUIRE(t->mode == LP_TP_ALLGATHER || t->mode == LP_TP_REDUCE_SCATTER, "lp_gemm_bf16_tp: bad mode %d", t->mode);
  LP_REQUIRE(d->M == t->tp_size * t->seg_rows && t->seg_rows > 0, "lp_gemm_bf16_tp: M=%d must be tp_size * seg_rows (%d x %d)", d->M,
             t->tp_size, t->seg_rows);
  LP_REQUIRE(d->N % 8 == 0 && d->K % 8 == 0 && d->lda % 8 == 0 && d->ldw % 8 == 0 && d->ldc % 8 == 0 && !d->out_fp32,
             "lp_gemm_bf16_tp: bf16 output, 16-byte aligned strides");
  LP_REQUIRE(d->residual == nullptr && d->norm_mode == LP_NORM_NONE && d->stats_out == nullptr && d->rope_heads == 0,
             "lp_gemm_bf16_tp: residual / norm / RoPE fusions are not combined with the tensor-parallel launches");
  LP_REQUIRE(d->act == LP_ACT_NONE || (d->act == LP_ACT_SWIGLU && d->N % 64 == 0 && t->mode == LP_TP_ALLGATHER),
             "lp_gemm_bf16_tp: act must be NONE (or SWIGLU after an all-gather)");
  LP_REQUIRE(t->counter != nullptr && t->signal_local != nullptr, "lp_gemm_bf16_tp: counter / signal words missing");
  constexpr int BN = 256;
  GemmArgs args = {};
  args.M = d->M; args.N = d->N; args.K = d->K;
  args.C = d->C; args.ldc = d->ldc; args.out_fp32 = 0;
  args.bias = d->bias; args.act = d->act;
  args.tp_size = t->tp_size; args.tp_rank = t->tp_rank; args.seg_rows = t->seg_rows;
  args.seg_m_blk = (t->seg_rows + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  args.num_m_blk = t->tp_size * args.seg_m_blk;
  args.num_n_blk = (d->N + BN - 1) / BN;
  args.num_k_blk = (d->K + GEMM_BK - 1) / GEMM_BK;
  args.group_m = args.seg_m_blk;          // one segment per raster group: the rank's own rows (all column blocks) come first
  args.epoch = t->epoch;
  args.recycle = t->recycle;
  const int n_out = d->act == LP_ACT_SWIGLU ? d->N / 2 : d->N;
  GemmSegMaps segs;
  int rc = 0;
  if (t->mode == LP_TP_ALLGATHER) {
    args.ag_counter = t->counter;
    args.ag_dst = const_cast<void*>(d->A);
    args.ag_ld = d->lda; args.ag_cols = d->K;
    for (int r = 0; r < t->tp_size; ++r) {
      LP_REQUIRE(r == t->tp_rank || (t->peer_A[r] != nullptr && t->signal_peer[r] != nullptr), "lp_gemm_bf16_tp: peer %d pointers missing", r);
      args.ag_src[r] = t->peer_A[r];
      args.ag_wait[r] = t->signal_local + r;                                   // written by rank r
      args.ag_post[r] = t->signal_peer[r] ? t->signal_peer[r] + t->tp_rank : nullptr;   // my word on rank r
      rc = make_tmap_2d(&segs.m[r], reinterpret_cast<__nv_bfloat16*>(d->C) + (size_t)r * t->seg_rows * d->ldc, 2, (uint64_t)n_out,
                        (uint64_t)t->seg_rows, (uint64_t)d->ldc, 32, 32, 64);
      if (rc) return rc;
    }
  } else {
    args.rs_counter = t->counter;
    for (int r = 0; r < t->tp_size; ++r) {
      LP_REQUIRE(t->peer_stage[r] != nullptr && t->signal_peer[r] != nullptr, "lp_gemm_bf16_tp: peer %d staging / signal pointers missing", r);
      args.rs_post[r] = t->signal_peer[r] + 8 + t->tp_rank;
      // rank r's staging [tp_size, seg_rows, ldc]: this rank's partial sums for r's rows go to slot [tp_rank]
      rc = make_tmap_2d(&segs.m[r], reinterpret_cast<__nv_bfloat16*>(t->peer_stage[r]) + (size_t)t->tp_rank * t->seg_rows * d->ldc, 2,
                        (uint64_t)n_out, (uint64_t)t->seg_rows, (uint64_t)d->ldc, 32, 32, 64);
      if (rc) return rc;
    }
  }
  for (int r = t->tp_size; r < 8; ++r) segs.m[r] = segs.m[0];
  CUtensorMap tmA, tmB;
  rc = make_tmap_bf16_2d(&tmA, d->A, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, GEMM_BK, GEMM_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, d->W, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldw, GEMM_BK, BN / 2);
  if (rc) return rc;
  using L = GemmSmem<BN, 6, true>;
  auto kern = gemm_bf16_tp_kernel<BN, 6, true>;
  static bool configured = false;
  if (!configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = args.num_m_blk * args.num_n_blk;
  const int pairs = sm_count() / 2;
  ProfScope prof(1, d->M, d->N, d->K, stream);
  note_launch();
  LP_CHECK_CUDA(launch_ex(kern, dim3(2 * (tiles < pairs ? tiles : pairs)), dim3(GEMM_THREADS), L::TOTAL, stream, 2, false, tmA, tmB, segs,
                          args));
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Stream-K workspace: `ws` = zero-filled device memory of at least lp_gemm_streamk_workspace_bytes() bytes that stays valid (and is
// used by ONE stream at a time) until it is replaced; NULL switches the stream-K tail off.
extern "C" size_t lp_gemm_streamk_workspace_bytes(void) {
  return SK_FLAG_BYTES + (size_t)sm_count() * GEMM_BM * 256 * sizeof(float);
}
extern "C" int lp_gemm_set_streamk_workspace(void* ws, size_t bytes) {
  LP_REQUIRE(ws == nullptr || ((reinterpret_cast<uintptr_t>(ws) & 255) == 0 && bytes > SK_FLAG_BYTES),
             "lp_gemm_set_streamk_workspace: workspace must be 256-byte aligned and larger than %zu bytes", SK_FLAG_BYTES);
  g_sk_ws = ws;
  g_sk_bytes = ws ? bytes : 0;
  return 0;
}
namespace lp {
// the workspace's flag page, shared with the chained projector launch (lp_gemm_chain.cu: words 511 ..)
unsigned* gemm_flag_page() { return reinterpret_cast<unsigned*>(g_sk_ws); }
}  // namespace lp

extern "C" int lp_gemm_set_epilogue_sleep(int ns) {
  const int old = (int)g_gemm_epi_sleep_ns;
  g_gemm_epi_sleep_ns = ns < 0 ? 0u : (unsigned)ns;
  return old;
}
extern "C" int lp_gemm_set_raster_balance(int enable) {
  const int old = g_gemm_raster_balance;
  g_gemm_raster_balance = enable ? 1 : 0;
  return old;
}
extern "C" int lp_gemm_set_tma_residual(int enable) {
  const int old = g_gemm_tma_residual;
  g_gemm_tma_residual = enable ? 1 : 0;
  return old;
}
extern "C" int lp_gemm_set_streamk(int enable) {
  const int old = g_sk_enable;
  g_sk_enable = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  return old;
}

extern "C" int lp_gemm_set_pair_mode(int enable) {
  const int old = g_gemm_pair_mode;
  g_gemm_pair_mode = enable == 2 ? 2 : (enable ? 1 : 0);   // 2 = 4-CTA clusters with A multicast (lp_gemm4.cu, opt-in)
  return old;
}
