// Phase trace of the v3 attention kernel (leopard_b200/csrc/lp_attn3.cu compiled with LP_ATTN_TRACE): one CTA's issuer threads and the
// first lane of each softmax warpgroup stamp clock64() at every protocol point; this tool launches the kernel on synthetic Q/K/V at a
// bench shape and prints, per stamped interval, the mean cycles per half step — where the ~4800 cycles of a 128-key step go.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/micro/attn_trace tools/micro/attn_trace.cu
//   tools/micro/attn_trace <T> <heads> <kv_heads> <head_dim 72|128> <causal> <pingpong> <blockIdx.x to trace>
#define LP_ATTN_TRACE 1
#include "../../leopard_b200/csrc/lp_attn3.cu"
#include "../../leopard_b200/csrc/lp_host.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace lp;

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 4796, H = argc > 2 ? atoi(argv[2]) : 32, HK = argc > 3 ? atoi(argv[3]) : 8;
  const int dh = argc > 4 ? atoi(argv[4]) : 128, causal = argc > 5 ? atoi(argv[5]) : 1, ping = argc > 6 ? atoi(argv[6]) : 1;
  const int blk = argc > 7 ? atoi(argv[7]) : 4, B = argc > 8 ? atoi(argv[8]) : 1;
  const int hs = dh == 72 ? 80 : 128;
  const long long ld = (long long)(H + 2 * HK) * hs;
  std::vector<__nv_bfloat16> h((size_t)B * T * ld);
  unsigned s = 12345u;
  for (size_t i = 0; i < h.size(); ++i) {
    s = s * 1664525u + 1013904223u;
    const int col = (int)(i % ld) % hs;
    h[i] = __float2bfloat16(col < dh ? ((int)(s >> 16) % 2001 - 1000) * 1e-3f : 0.f);
  }
  __nv_bfloat16 *qkv, *out;
  cudaMalloc(&qkv, h.size() * 2);
  cudaMalloc(&out, (size_t)B * T * H * dh * 2);
  cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  unsigned long long* tr;
  cudaMalloc(&tr, 4 * 4096 * 8);
  cudaMemset(tr, 0, 4 * 4096 * 8);
  AttnArgs a;
  a.T = T; a.Tk = T; a.B = B; a.n_heads = H; a.group = H / HK;
  a.q_off = 0; a.k_off = H * hs; a.v_off = (H + HK) * hs; a.hs = hs; a.dh_out = dh;
  a.O = out; a.ldo = (long long)H * dh;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)dh);
  a.pingpong = ping;
  CUtensorMap tm;
  if (make_tmap_bf16_3d(&tm, qkv, (uint64_t)ld, (uint64_t)T, (uint64_t)B, (uint64_t)ld, (uint64_t)T * ld, 64, 128, 1)) {
    printf("tensor map failed: %s\n", get_error());
    return 1;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch_attn3(tm, a, dh, causal, 0, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) launch_attn3(tm, a, dh, causal, 0, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpyToSymbol(g_a3_trace_blk, &blk, sizeof(int));
  cudaMemcpyToSymbol(g_a3_trace, &tr, sizeof(tr));
  launch_attn3(tm, a, dh, causal, 0, 0);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("cuda error %s / %s\n", cudaGetErrorString(err), get_error()); return 1; }
  std::vector<unsigned long long> t(4 * 4096);
  cudaMemcpy(t.data(), tr, t.size() * 8, cudaMemcpyDeviceToHost);
  printf("{\"T\": %d, \"heads\": %d, \"head_dim\": %d, \"causal\": %d, \"pingpong\": %d, \"block\": %d, \"ms_untraced\": %.4f, \"trace\": {", T, H, dh,
         causal, ping, blk, ms / 10);
  const char* names[4] = {"issuer_A", "issuer_B", "softmax_A", "softmax_B"};
  unsigned long long t0 = ~0ULL;
  for (int sl = 0; sl < 4; ++sl)
    if (t[sl * 4096] > 0) t0 = std::min(t0, t[sl * 4096 + 1] & 0xffffffffffffULL);
  for (int sl = 0; sl < 4; ++sl) {
    const int n = (int)t[sl * 4096];
    printf("%s\"%s\": [", sl ? ", " : "", names[sl]);
    for (int i = 0; i < n; ++i) {
      const unsigned long long w = t[sl * 4096 + 1 + i];
      printf("%s[%d, %llu]", i ? ", " : "", (int)(w >> 48), (w & 0xffffffffffffULL) - t0);
    }
    printf("]");
  }
  printf("}}\n");
  return 0;
}
