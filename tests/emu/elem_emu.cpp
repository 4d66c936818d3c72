// Test infrastructure: the HBM-bound glue kernels of the prefill path (leopard_b200/csrc/lp_elem.cu) run on the CPU under
// tests/emu/cuda_emu.h.  `elem_kernels.inc` is the kernel part of lp_elem.cu cut by the test (the host launcher in the middle of
// it removed); the kernels are parity-green on the B200 — this is their CPU regression check.
#include "cuda_emu.h"
#include <cstdio>

namespace lp {
typedef __nv_bfloat16 bf16;
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
}
#include "elem_helpers.inc"
#include "elem_kernels.inc"

using namespace lp;
static const bf16* B(const void* p) { return reinterpret_cast<const bf16*>(p); }
static bf16* B(void* p) { return reinterpret_cast<bf16*>(p); }

extern "C" {

// variant: 0 = norm_kernel (LayerNorm when beta != NULL or rms == 0), 1 = rmsnorm_packed_kernel<16>, 2 = rmsnorm_reread_kernel
void emu_norm(const void* x, long long ldx, const float* gamma, const float* beta, void* out, long long ldo, int M, int D, float eps,
              int shuffle_side, int rms, int variant) {
  if (variant == 1) {
    emu::launch(dim3((M + 7) / 8), dim3(256), [&] { rmsnorm_packed_kernel<16>(B(x), ldx, gamma, B(out), ldo, M, D, eps); });
  } else if (variant == 2) {
    emu::launch(dim3((M + 7) / 8), dim3(256), [&] { rmsnorm_reread_kernel(B(x), ldx, gamma, B(out), ldo, M, D, eps); });
  } else {
    const int grid = (M + NORM_THREADS / 32 - 1) / (NORM_THREADS / 32);
    emu::launch(dim3(grid), dim3(NORM_THREADS), [&] {
      if (D <= 32 * 8 * 5) {
        if (rms) norm_kernel<true, 5>(B(x), ldx, gamma, beta, B(out), ldo, M, D, eps, shuffle_side);
        else norm_kernel<false, 5>(B(x), ldx, gamma, beta, B(out), ldo, M, D, eps, shuffle_side);
      } else {
        if (rms) norm_kernel<true, 16>(B(x), ldx, gamma, beta, B(out), ldo, M, D, eps, shuffle_side);
        else norm_kernel<false, 16>(B(x), ldx, gamma, beta, B(out), ldo, M, D, eps, shuffle_side);
      }
    });
  }
}

void emu_rope_table(const int* pos, const float* inv_freq, float* cs, float* sn, int S, int half) {
  emu::launch(dim3((S * half + 255) / 256), dim3(256), [&] { rope_table_kernel(pos, 0, inv_freq, cs, sn, S, half); });
}

void emu_rope(void* x, long long ld, int S, int col0, int n_heads, int head_dim, const float* cs, const float* sn) {
  const int half = head_dim / 2;
  const long long total = (long long)S * n_heads * (half / 8);
  emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), [&] { rope_kernel(B(x), ld, S, col0, n_heads, half, cs, sn); });
}

void emu_rope_kv(void* x, long long ld, int S, int n_q, int n_kv, int head_dim, const float* cs, const float* sn, void* kc, void* vc,
                 long long ld_cache) {
  const int half = head_dim / 2;
  const long long total = (long long)S * (n_q + 2 * n_kv) * (half / 8);
  emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256),
              [&] { rope_kv_kernel(B(x), ld, S, n_q, n_kv, half, cs, sn, B(kc), B(vc), ld_cache); });
}

void emu_embed_merge(const long long* ids, int L, long long image_id, int rows_per_slot, const void* table, long long ldt,
                     long long vocab, const void* feats, long long ldf, int n_feat_rows, void* out, long long ldo, int D, int* plan) {
  int* dst = plan;
  int* src = plan + L;
  int* totals = plan + 2 * L;
  emu::launch(dim3(1), dim3(1024), [&] { merge_plan_kernel(ids, L, image_id, rows_per_slot, dst, src, totals); });
  emu::launch(dim3(L, rows_per_slot), dim3(128), [&] {
    merge_copy_kernel(ids, dst, src, B(table), ldt, vocab, B(feats), ldf, n_feat_rows, B(out), ldo, D, rows_per_slot);
  });
}

void emu_im2col(const void* pix, int fp32, void* out, long long ldo, int N, int C, int H, int W) {
  const long long total = (long long)N * (H / 14) * (W / 14) * (C * 14 + 1);
  emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), [&] {
    if (fp32) im2col_patch14_kernel<float>(reinterpret_cast<const float*>(pix), B(out), ldo, N, C, H, W);
    else im2col_patch14_kernel<bf16>(B(pix), B(out), ldo, N, C, H, W);
  });
}

void emu_pixel_shuffle(const void* x, long long ldx, void* out, long long ldo, int M, int D, int side) {
  emu::launch(dim3(M), dim3(128), [&] { pixel_shuffle_kernel(B(x), ldx, B(out), ldo, side, D); });
}

void emu_cross_entropy(const void* logits, int fp32, long long ld, const long long* labels, float* loss, int rows, int V) {
  emu::launch(dim3(rows), dim3(256), [&] {
    if (fp32) cross_entropy_kernel<float>(reinterpret_cast<const float*>(logits), ld, labels, loss, V);
    else cross_entropy_kernel<bf16>(B(logits), ld, labels, loss, V);
  });
}

}  // extern "C"
