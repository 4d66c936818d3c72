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
// cudaLaunchKernelEx (programmatic-dependent / cooperative / cluster launch attributes are accepted and ignored: blocks run one
// after another, so a cooperative kernel that needs a grid-wide barrier cannot run here)
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 1, cudaLaunchAttributeCooperative = 2,
                             cudaLaunchAttributeClusterDimension = 3 };
struct cudaLaunchAttribute {
  cudaLaunchAttributeID id;
  struct { int programmaticStreamSerializationAllowed; int cooperative; struct { unsigned x, y, z; } clusterDim; } val;
};
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes = 0;
  cudaStream_t stream = nullptr;
  cudaLaunchAttribute* attrs = nullptr;
  unsigned numAttrs = 0;
};
template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), Args... args) {
  emu::launch(cfg->gridDim, cfg->blockDim, [&] { kern(static_cast<KArgs>(args)...); });
  return cudaSuccess;
}
// streams / graphs / copies as far as host code names them
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
constexpr int cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaMemcpyDeviceToDevice = 3;
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t*, int) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return 1; }      // graphs are not emulated
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return 1; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, int) { return 1; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return 1; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline void __threadfence() {}
inline long long clock64() { return 0; }
inline void __trap() { std::abort(); }
template <typename T> inline T __ldcg(const T* p) { return *p; }
inline float __bfloat162float(__nv_bfloat16 v) { return (float)v; }
inline __nv_bfloat16 __ushort_as_bfloat16(unsigned short u) { __nv_bfloat16 r; r.bits = u; return r; }
#define __align__(n) alignas(n)
#define LP_WATCHDOG_CYCLES (1LL << 62)
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
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline int sm_count() { return 4; }
inline uint32_t smem_u32(const void* p) { return (uint32_t)(uintptr_t)p; }
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
