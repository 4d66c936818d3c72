// Microbenchmark (development aid, not part of the product): issue rates that bound the softmax section of the attention kernels on
// sm_100a — MUFU.EX2 (f32 and bf16x2), packed FFMA2 / FADD2, F2FP, and a replica of the kernel's 64-element exp half-step —
// for 1, 2 and 4 warps per SM sub-partition.  Prints cycles per warp-instruction (per SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[64], w[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) { v[i] = -0.01f * (threadIdx.x + i); w[i] = -0.02f * (threadIdx.x + i); }
  __syncthreads();
  const long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 64 independent MUFU.EX2
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = ex2(v[i]);
    } else if (MODE == 1) {   // 32 FFMA2
      const uint64_t s2 = pk2(1.0001f, 1.0001f), c2 = pk2(-0.001f, -0.001f);
#pragma unroll
      for (int i = 0; i < 64; i += 2) { uint64_t r = fma2(pk2(v[i], v[i + 1]), s2, c2); upk2(r, v[i], v[i + 1]); }
    } else if (MODE == 2) {   // the kernel's half-step: fma2 -> 2 x ex2 -> add2 (row sum) -> pack bf16x2
      const uint64_t s2 = pk2(1.0001f, 1.0001f), c2 = pk2(-0.001f, -0.001f);
      uint64_t a2 = pk2(0.f, 0.f);
      uint32_t pkd = 0;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float x0, x1;
        upk2(fma2(pk2(v[i], v[i + 1]), s2, c2), x0, x1);
        const float p0 = ex2(x0), p1 = ex2(x1);
        a2 = add2(a2, pk2(p0, p1));
        __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
        pkd ^= *reinterpret_cast<uint32_t*>(&b);
        v[i] = x0; v[i + 1] = x1;
      }
      float s0, s1; upk2(a2, s0, s1); acc += s0 + s1 + __uint_as_float(pkd & 0x7fu);
    } else if (MODE == 4) {   // three passes with pinned order: 32 FFMA2, then 64 MUFU back to back, then 32 FADD2 + packs
      const uint64_t s2 = pk2(1.0001f, 1.0001f), c2 = pk2(-0.001f, -0.001f);
#pragma unroll
      for (int i = 0; i < 64; i += 2) { uint64_t r = fma2(pk2(v[i], v[i + 1]), s2, c2); upk2(r, v[i], v[i + 1]); }
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = ex2(v[i]);
      uint64_t a2 = pk2(0.f, 0.f), b2 = pk2(0.f, 0.f);
      uint32_t pkd = 0;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        a2 = add2(a2, pk2(v[i], v[i + 1]));
        b2 = add2(b2, pk2(v[i + 2], v[i + 3]));
        __nv_bfloat162 q0 = __floats2bfloat162_rn(v[i], v[i + 1]), q1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
        pkd ^= *reinterpret_cast<uint32_t*>(&q0) ^ *reinterpret_cast<uint32_t*>(&q1);
      }
      float s0, s1, s3, s4; upk2(a2, s0, s1); upk2(b2, s3, s4); acc += s0 + s1 + s3 + s4 + __uint_as_float(pkd & 0x7fu);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = -v[i] * 0.5f - 0.01f;
    } else if (MODE == 5) {   // software-pipelined in 16-value groups: MUFU of group g + 1 issued before the adds / packs of group g
      const uint64_t s2 = pk2(1.0001f, 1.0001f), c2 = pk2(-0.001f, -0.001f);
      uint64_t a2 = pk2(0.f, 0.f), b2 = pk2(0.f, 0.f);
      uint32_t pkd = 0;
#pragma unroll
      for (int i = 0; i < 16; i += 2) { uint64_t r = fma2(pk2(v[i], v[i + 1]), s2, c2); upk2(r, v[i], v[i + 1]); }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ex2(v[i]);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g < 3) {
#pragma unroll
          for (int i = 16 * (g + 1); i < 16 * (g + 2); i += 2) { uint64_t r = fma2(pk2(v[i], v[i + 1]), s2, c2); upk2(r, v[i], v[i + 1]); }
#pragma unroll
          for (int i = 16 * (g + 1); i < 16 * (g + 2); ++i) v[i] = ex2(v[i]);
        }
#pragma unroll
        for (int i = 16 * g; i < 16 * (g + 1); i += 4) {
          a2 = add2(a2, pk2(v[i], v[i + 1]));
          b2 = add2(b2, pk2(v[i + 2], v[i + 3]));
          __nv_bfloat162 q0 = __floats2bfloat162_rn(v[i], v[i + 1]), q1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
          pkd ^= *reinterpret_cast<uint32_t*>(&q0) ^ *reinterpret_cast<uint32_t*>(&q1);
        }
      }
      float s0, s1, s3, s4; upk2(a2, s0, s1); upk2(b2, s3, s4); acc += s0 + s1 + s3 + s4 + __uint_as_float(pkd & 0x7fu);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = -v[i] * 0.5f - 0.01f;
    } else if (MODE == 6) {   // TWO independent 64-value streams in one warp (rows of two query tiles per thread): 128 values
      const uint64_t s2 = pk2(1.0001f, 1.0001f), c2 = pk2(-0.001f, -0.001f);
      uint64_t a2 = pk2(0.f, 0.f), b2 = pk2(0.f, 0.f);
      uint32_t pka = 0, pkb = 0;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float x0, x1, y0, y1;
        upk2(fma2(pk2(v[i], v[i + 1]), s2, c2), x0, x1);
        upk2(fma2(pk2(w[i], w[i + 1]), s2, c2), y0, y1);
        const float p0 = ex2(x0), p1 = ex2(x1), q0 = ex2(y0), q1 = ex2(y1);
        a2 = add2(a2, pk2(p0, p1));
        b2 = add2(b2, pk2(q0, q1));
        __nv_bfloat162 ba = __floats2bfloat162_rn(p0, p1), bb = __floats2bfloat162_rn(q0, q1);
        pka ^= *reinterpret_cast<uint32_t*>(&ba);
        pkb ^= *reinterpret_cast<uint32_t*>(&bb);
        v[i] = x0; v[i + 1] = x1; w[i] = y0; w[i + 1] = y1;
      }
      float s0, s1, s3, s4; upk2(a2, s0, s1); upk2(b2, s3, s4); acc += s0 + s1 + s3 + s4 + __uint_as_float((pka ^ pkb) & 0x7fu);
    } else if (MODE == 3) {   // bf16x2 exponentials: 32 MUFU-class ops for 64 values
      uint32_t pkd = 0;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        __nv_bfloat162 b = __floats2bfloat162_rn(v[i], v[i + 1]);
        uint32_t y = ex2_bf16x2(*reinterpret_cast<uint32_t*>(&b));
        pkd ^= y;
      }
      acc += __uint_as_float(pkd & 0x7fu);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] += 1e-7f;
    }
  }
  const long long t1 = clock64();
  float s = acc;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += v[i] + (MODE == 6 ? w[i] : 0.f);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter_instr) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp, iters = 2000;
    k<MODE><<<148, threads>>>(out, cyc, iters);
    k<MODE><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("{\"case\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_iteration\": %.1f, \"smsp_cycles_per_warp_instr\": %.2f}\n", name,
           warps_per_smsp, avg / iters, avg / iters / (per_iter_instr * warps_per_smsp));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("mufu_ex2_f32 (64 per iter)", 64);
  run<1>("ffma2 (32 per iter)", 32);
  run<2>("softmax half-step replica (64 values)", 64);
  run<3>("ex2 bf16x2 (32 per iter, 64 values)", 32);
  run<4>("three passes, pinned order (64 values)", 64);
  run<5>("software-pipelined 16-value groups (64 values)", 64);
  run<6>("two 64-value streams in one warp (128 values, 128 MUFU)", 128);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
