// Test infrastructure: the decode step (leopard_b200/csrc/lp_decode.cu — the GEMV / split-K attention / embed / arg-max kernels AND
// the host functions that sequence them: lp_decode_step, lp_decode_greedy_step in their default mode) run on the CPU.
// `decode_src.inc` is the file itself with its #include lines removed, the dynamic shared-memory declarations dropped (buffers
// below), its one-line inline-PTX helpers replaced by plain C++ (streaming loads, cp.async as a synchronous copy, ex2) and its
// launches rewritten (tests/emu/launch_rewrite.py).  Not emulated: CUDA graphs, the cooperative single-kernel mode (it needs
// co-resident CTAs), the first implementation's calls into the prefill kernels (stubs below return an error).
#include "host_emu.h"
#include "../../include/leopard_b200.h"

namespace lp {
static uint4 xs[8192];                 // gemv_kernel's staged activation row
static uint4 xs3[8192];                // gemv3 / mega kernels
alignas(16) static uint8_t da_smem[96 * 1024];
}
#include "elem_helpers.inc"

// entry points of other translation units that only the first implementation (mode 0) calls
extern "C" {
int lp_rmsnorm(const void*, long long, const float*, void*, long long, int, int, float, void*) { return -1; }
int lp_gemm_bf16(const void*, long long, const void*, long long, void*, long long, int, int, int, const float*, const void*, long long,
                 int, int, int, int, void*) { return -1; }
int lp_rope_inplace(void*, long long, int, int, int, int, const float*, const float*, void*) { return -1; }
int lp_embed_merge(const long long*, int, long long, int, const void*, long long, long long, const void*, long long, int, void*,
                   long long, int, int, int*, void*) { return -1; }
long long lp_launch_count(void) { return 0; }
}
namespace lp {
int rope_table_base(int, const float*, float*, float*, int, int, cudaStream_t) { return -1; }
}
#include "decode_src.inc"
