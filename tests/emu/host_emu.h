// Test infrastructure: the handful of CUDA runtime names a kernel file's HOST code uses, so that whole .cu files (kernels + the
// host function that sequences them) can run on the CPU under cuda_emu.h once the test has rewritten every
// `kernel<<<grid, block, smem, stream>>>(args);` into `emu::launch(dim3(grid), dim3(block), [&] { kernel(args); });`.
#pragma once
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
struct int2 { int x, y; };
inline int2 make_int2(int a, int b) { return {a, b}; }

// fp64 arithmetic without contraction (the test compiles with -ffp-contract=off)
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }

namespace lp {
typedef __nv_bfloat16 bf16;
inline char g_emu_err[512];
inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(g_emu_err, sizeof(g_emu_err), fmt, ap);
  va_end(ap);
}
inline void note_launch(int = 1) {}
struct ProfScope { ProfScope(int, int, int, int, cudaStream_t) {} };
}  // namespace lp

#define LP_CHECK_CUDA(expr)                                        \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) { lp::set_error("cuda error"); return -2; } \
  } while (0)
#define LP_REQUIRE(cond, ...)     \
  do {                            \
    if (!(cond)) {                \
      lp::set_error(__VA_ARGS__); \
      return -1;                  \
    }                             \
  } while (0)
