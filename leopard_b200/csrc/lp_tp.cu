// leopard_b200 — tensor-parallel reduce step: the owner-side half of the fused GEMM -> reduce-scatter (lp_gemm_bf16_tp,
// LP_TP_REDUCE_SCATTER).  Every rank's GEMM epilogue has TMA-stored its partial sums for this rank's rows into slot [src] of the
// local staging buffer over NVLink and posted its epoch; this kernel waits for all of them and adds the slots in rank order
// (deterministic) to the residual stream.  Reference: reduce_scatter_to_sequence_parallel_region + bias_dropout_add,
// Megatron-LM-240603/megatron/core/tensor_parallel/mappings.py:119-199, megatron_patch/model/llava/transformer.py:888-894.
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

using namespace lp;

namespace {
struct ReduceWait {
  const unsigned* w[8];
};

__global__ void __launch_bounds__(256)
tp_reduce_kernel(const bf16* __restrict__ stage, long long ld, int tp, int seg_rows, const bf16* __restrict__ residual, long long ldr,
                 bf16* __restrict__ out, long long ldo, int cols, ReduceWait wait, unsigned epoch) {
  if (threadIdx.x == 0) {
    for (int r = 0; r < tp; ++r) spin_until_sys(wait.w[r], epoch);
  }
  __syncthreads();
  const int vec_per_row = cols >> 3;
  const long long total = (long long)seg_rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long row = i / vec_per_row;
    const int c = (int)(i % vec_per_row) * 8;
    float acc[8];
    {
      const uint4 u = *reinterpret_cast<const uint4*>(residual + row * ldr + c);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc[0] = a.x; acc[1] = a.y; acc[2] = b.x; acc[3] = b.y; acc[4] = cc.x; acc[5] = cc.y; acc[6] = d.x; acc[7] = d.y;
    }
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < tp; ++r) {      // rank order: the sum does not depend on arrival order
      const uint4 u = *reinterpret_cast<const uint4*>(stage + ((long long)r * seg_rows + row) * ld + c);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += cc.x; s[5] += cc.y; s[6] += d.x; s[7] += d.y;
    }
    uint4 o;
    // the reference sums bf16 partials in the collective (bf16 result) and then adds the residual in bf16
    o.x = pack_bf16x2(acc[0] + __bfloat162float(__float2bfloat16_rn(s[0])), acc[1] + __bfloat162float(__float2bfloat16_rn(s[1])));
    o.y = pack_bf16x2(acc[2] + __bfloat162float(__float2bfloat16_rn(s[2])), acc[3] + __bfloat162float(__float2bfloat16_rn(s[3])));
    o.z = pack_bf16x2(acc[4] + __bfloat162float(__float2bfloat16_rn(s[4])), acc[5] + __bfloat162float(__float2bfloat16_rn(s[5])));
    o.w = pack_bf16x2(acc[6] + __bfloat162float(__float2bfloat16_rn(s[6])), acc[7] + __bfloat162float(__float2bfloat16_rn(s[7])));
    *reinterpret_cast<uint4*>(out + row * ldo + c) = o;
  }
}
}  // namespace

extern "C" int lp_tp_reduce(const lp_tp_desc* t, const void* stage, long long ld, const void* residual, long long ldr, void* out,
                            long long ldo, int cols, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(t && stage && residual && out && t->signal_local, "lp_tp_reduce: null argument");
  LP_REQUIRE(t->tp_size >= 2 && t->tp_size <= 8 && t->seg_rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && ldr % 8 == 0 &&
                 ldo % 8 == 0,
             "lp_tp_reduce: bad shape tp=%d seg_rows=%d cols=%d", t->tp_size, t->seg_rows, cols);
  ReduceWait w;
  for (int r = 0; r < 8; ++r) w.w[r] = t->signal_local + 8 + (r < t->tp_size ? r : 0);
  const long long total = (long long)t->seg_rows * (cols / 8);
  const unsigned grid = (unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  ProfScope prof(3, t->seg_rows, cols, 3, stream);
  note_launch();
  tp_reduce_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(stage), ld, t->tp_size, t->seg_rows,
                                             reinterpret_cast<const bf16*>(residual), ldr, reinterpret_cast<bf16*>(out), ldo, cols, w,
                                             t->epoch);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
