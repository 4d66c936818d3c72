// Phase trace of the persistent GEMM (leopard_b200/csrc/lp_gemm.cu compiled with LP_GEMM_TRACE): the MMA-issuing thread and lane 0 of
// the first epilogue warp of each warpgroup of one CTA stamp clock64(); prints per-interval means — is a short-K GEMM waiting for
// operands, for the accumulator buffer, or for its epilogue (TMEM load / math / staging + TMA store)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lcuda -o tools/micro/gemm_trace tools/micro/gemm_trace.cu
//   tools/micro/gemm_trace <M> <N> <K> <bias 0|1> <act> <residual 0|1> <cta to trace>
#define LP_GEMM_TRACE 1
#include "../../leopard_b200/csrc/lp_host.cu"
#include "../../leopard_b200/csrc/lp_gemm4.cu"
#include "../../leopard_b200/csrc/lp_gemm.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <map>

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 18928, N = argc > 2 ? atoi(argv[2]) : 3840, K = argc > 3 ? atoi(argv[3]) : 1152;
  const int has_bias = argc > 4 ? atoi(argv[4]) : 1, act = argc > 5 ? atoi(argv[5]) : 0, has_res = argc > 6 ? atoi(argv[6]) : 0;
  const int blk = argc > 7 ? atoi(argv[7]) : 10;
  std::vector<__nv_bfloat16> ha((size_t)M * K), hw((size_t)N * K);
  unsigned s = 777u;
  for (auto& x : ha) { s = s * 1664525u + 1013904223u; x = __float2bfloat16(((int)(s >> 16) % 2001 - 1000) * 1e-3f); }
  for (auto& x : hw) { s = s * 1664525u + 1013904223u; x = __float2bfloat16(((int)(s >> 16) % 2001 - 1000) * 3e-5f); }
  __nv_bfloat16 *A, *W, *C;
  float* bias;
  cudaMalloc(&A, ha.size() * 2); cudaMalloc(&W, hw.size() * 2); cudaMalloc(&C, (size_t)M * N * 2); cudaMalloc(&bias, N * 4);
  cudaMemcpy(A, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(W, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(C, 0, (size_t)M * N * 2); cudaMemset(bias, 0, N * 4);
  void* ws; size_t ws_bytes = 64u << 20;
  cudaMalloc(&ws, ws_bytes); cudaMemset(ws, 0, ws_bytes);
  lp_gemm_set_streamk_workspace(ws, ws_bytes);
  lp_gemm_set_pair_mode(1);
  unsigned long long* tr;
  cudaMalloc(&tr, 3 * 4096 * 8); cudaMemset(tr, 0, 3 * 4096 * 8);
  lp_gemm_desc d = {};
  d.A = A; d.lda = K; d.W = W; d.ldw = K; d.C = C; d.ldc = N; d.M = M; d.N = N; d.K = K;
  d.bias = has_bias ? bias : nullptr; d.act = act;
  if (has_res) { d.residual = C; d.ldr = N; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) if (lp_gemm_bf16_ex(&d, 0)) { printf("gemm failed: %s\n", lp_last_error()); return 1; }
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) lp_gemm_bf16_ex(&d, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  cudaMemcpyToSymbol(lp::g_gemm_trace_blk, &blk, sizeof(int));
  cudaMemcpyToSymbol(lp::g_gemm_trace, &tr, sizeof(tr));
  lp_gemm_bf16_ex(&d, 0);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(err)); return 1; }
  std::vector<unsigned long long> t(3 * 4096);
  cudaMemcpy(t.data(), tr, t.size() * 8, cudaMemcpyDeviceToHost);
  printf("M=%d N=%d K=%d bias=%d act=%d res=%d: %.4f ms = %.0f TFLOP/s; traced CTA %d\n", M, N, K, has_bias, act, has_res, ms,
         2.0 * M * N * K / ms / 1e9, blk);
  const char* names[3] = {"mma issuer", "epilogue warp 2 (wg0)", "epilogue warp 6 (wg1)"};
  for (int sl = 0; sl < 3; ++sl) {
    const int n = (int)t[sl * 4096];
    if (n < 2) continue;
    std::map<std::pair<int, int>, std::pair<double, int>> acc;
    for (int i = 0; i + 1 < n; ++i) {
      const unsigned long long a = t[sl * 4096 + 1 + i], b = t[sl * 4096 + 2 + i];
      auto& e = acc[{(int)(a >> 48), (int)(b >> 48)}];
      e.first += (double)((b & 0xffffffffffffULL) - (a & 0xffffffffffffULL));
      e.second += 1;
    }
    const double span = (double)((t[sl * 4096 + n] & 0xffffffffffffULL) - (t[sl * 4096 + 1] & 0xffffffffffffULL));
    printf(" %s: %d stamps, %.0f cycles\n", names[sl], n, span);
    for (auto& kv : acc)
      printf("   %2d -> %2d  n=%4d  mean %8.0f  total %9.0f (%4.1f %%)\n", kv.first.first, kv.first.second, kv.second.second,
             kv.second.first / kv.second.second, kv.second.first, 100.0 * kv.second.first / span);
  }
  printf(" ids: 20 before / 21 after the wait for a free accumulator buffer, 22 tile's MMAs issued + committed; 30 before / 31 after the wait for the\n"
         " accumulator, per 32-column chunk: 35 chunk start (ring kernels), 32 TMEM load done, 33 math done, 36 residual tile arrived, 34 staging +\n"
         " TMA store issued, 37 tile done (before the buffer is handed back)\n");
  return 0;
}
