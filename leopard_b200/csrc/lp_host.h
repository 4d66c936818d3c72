// leopard_b200 — host-side helpers shared by the .cu translation units (not part of the public C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace lp {

// error plumbing: every entry point returns 0 or a negative lp_status and records a message
void set_error(const char* fmt, ...);
const char* get_error();

#define LP_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      lp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return -2; /* LP_ERR_CUDA */                                                            \
    }                                                                                         \
  } while (0)

#define LP_REQUIRE(cond, ...)           \
  do {                                  \
    if (!(cond)) {                      \
      lp::set_error(__VA_ARGS__);       \
      return -1; /* LP_ERR_ARG */       \
    }                                   \
  } while (0)

int sm_count();

// prefill kernels are launched with the programmatic-dependent-launch attribute when this is on (lp_set_prefill_pdl):
// every such kernel calls griddepcontrol.wait before its first dependent global access
bool prefill_pdl();

// kernel launch with optional cluster dimension and optional programmatic dependent launch
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                             bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// launch accounting + optional per-launch CUDA-event timing (off by default; see lp_profile_* in the public header)
void note_launch(int n = 1);
struct ProfScope {
  int idx;
  cudaStream_t stream;
  ProfScope(int kind, int a, int b, int c, cudaStream_t s);
  ~ProfScope();
};

// RoPE in place on the q|k heads of x [S, ld] fused with the KV-cache append of the same rows (lp_elem.cu)
int rope_kv_append(void* x, long long ld, int S, int n_q, int n_kv, int head_dim, const float* cos_t, const float* sin_t,
                   void* kc, void* vc, long long ld_cache, cudaStream_t stream);

// attention launch arguments (shared by lp_attn.cu and lp_attn2.cu)
struct AttnArgs {
  int T, B, n_heads, group;   // T = query rows per sequence
  int Tk;                     // key/value rows per sequence (== T for self-attention)
  int q_off, k_off, v_off, hs;
  int dh_out;
  __nv_bfloat16* O;
  long long ldo;
  float scale_log2;
  int pingpong;               // lp_attn2: strict A/B alternation of the softmax warpgroups (turn protocol)
};
int launch_attn2(const CUtensorMap& tm, const AttnArgs& a, int head_dim, int causal, int emu, cudaStream_t stream);
int launch_attn3(const CUtensorMap& tm, const AttnArgs& a, int head_dim, int causal, int emu, cudaStream_t stream);
int launch_attn4(const CUtensorMap& tm, const AttnArgs& a, int head_dim, int causal, int emu, cudaStream_t stream);
// persistent v3 for short non-causal sequences (lp_attn5.cu): the ViT tower's T = 676
bool attn5_eligible(const AttnArgs& a, int head_dim, int causal);
int launch_attn5(const CUtensorMap& tm, const AttnArgs& a, cudaStream_t stream);

// 2-D / 3-D bf16 tensor maps with 128B swizzle (inner box = 64 elements = 128 B), zero OOB fill.
// dims/strides are innermost-first; strides are in ELEMENTS for dims >= 1.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer);
// generic 2-D map: elem_bytes 2 (bf16); swizzle_bytes in {0, 32, 64, 128}; inner box bytes must equal the swizzle span
int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                 uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1_elems,
                      uint64_t ld2_elems, uint32_t b0, uint32_t b1, uint32_t b2);

}  // namespace lp
