// leopard_b200 — GPU image front-end: Pillow-exact bicubic resize + pad + 364x364 tiling + SigLIP normalisation.
//
// Replaces the per-crop CPU work the reference does before the model sees an image
// (evaluations/models/llava_multiimg_siglip_anyres.py):
//   resize_and_pad_image :102-140  aspect-preserving PIL.Image.resize (default BICUBIC) pasted centred on black
//   divide_to_patches    :143-162  row-major 364 x 364 tiles
//   image_processor.preprocess :403-404 (SiglipImageProcessor, slow/PIL backend of transformers <= 4.4x):
//                                  resize to 364 x 364 (bicubic) -> x * 1/255 (float64, rounded to float32)
//                                  -> (x - mean) / std in float32
// (training twin: Pai-Megatron-Patch/megatron_patch/data/llava/mm_pretrain_dataset.py:141-201,264-303).
//
// The arithmetic is Pillow's ImagingResample (src/libImaging/Resample.c, not vendored by the reference): separable,
// horizontal pass first, uint8 intermediate; per output sample the filter taps are computed in double precision,
// normalised, quantised to 22 fractional bits, and each pass computes clip8((sum + 2^21) >> 22).  All of it is integer
// / byte work bound by HBM traffic, so the kernels are about coalescing:
//   resample_coeff_kernel   one thread per output sample builds its taps in fp64 with explicitly rounded operations
//                           (__dmul_rn / __dadd_rn / __ddiv_rn — no FMA contraction), stored transposed [tap][sample]
//   resample_h_kernel       one CTA per image row: the row is staged in shared memory with 16-byte loads, outputs are
//                           staged and leave with 16-byte stores into a pitched (16 B aligned) intermediate
//   resample_v_kernel       one thread per 4 consecutive bytes of an output row: 32-bit coalesced loads per tap, the
//                           epilogue applies the 256-entry normalisation table and scatters into the planar
//                           [crop, channel, 364, 364] layout (pad + tile + normalise fused into the store)
//   pad_fill_kernel         writes the normalised black value outside the pasted rectangle
// Results are bit-identical to Pillow + the PIL-backed SiglipImageProcessor (tests/test_gpu_frontend.py).
#include "lp_common.cuh"
#include "lp_host.h"
#include "../../include/leopard_b200.h"

namespace lp {

constexpr int FE_PRECISION_BITS = 32 - 8 - 2;
constexpr int FE_CROP = 364;
constexpr int FE_MAX_KSIZE = 131;  // downscale factors up to 32x
enum { FE_BICUBIC = 0, FE_BILINEAR = 1 };

struct FeLut { float v[3][256]; };  // normalised fp32 value of every uint8 level, per channel

// ---- coefficients (Pillow precompute_coeffs + normalize_coeffs_8bpc, bicubic a = -0.5, whole-image box)
__device__ __forceinline__ double fe_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0)  // ((a + 2) x - (a + 3)) x x + 1
    return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(a + 2.0, x), a + 3.0), x), x), 1.0);
  if (x < 2.0)  // (((x - 5) x + 8) x - 4) a
    return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}

// triangle filter, support 1 (Pillow `bilinear_filter`): Idefics2ImageProcessor's resample
__device__ __forceinline__ double fe_bilinear(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}
__device__ __forceinline__ double fe_filter(int filter, double x) { return filter == FE_BILINEAR ? fe_bilinear(x) : fe_bicubic(x); }

// bounds[o] = (first input sample, tap count); coeffs[t * out_size + o] = tap t of output sample o (0 beyond count).
// identity != 0: the pass is skipped by Pillow (size unchanged) — a single unit tap reproduces the copy exactly.
__global__ void resample_coeff_kernel(int in_size, int out_size, int ksize, int identity, int filter, int2* __restrict__ bounds,
                                      int* __restrict__ coeffs) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out_size) return;
  if (identity) {
    bounds[o] = make_int2(o, 1);
    coeffs[o] = 1 << FE_PRECISION_BITS;
    return;
  }
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(filter == FE_BILINEAR ? 1.0 : 2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn((double)o + 0.5, scale));
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > ksize) xmax = ksize;  // cannot happen (ksize = 2 ceil(support) + 1); keeps stores in range regardless
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    ww = __dadd_rn(ww, fe_filter(filter, arg));
  }
  for (int x = 0; x < ksize; ++x) {
    int q = 0;
    if (x < xmax) {
      const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
      double k = fe_filter(filter, arg);
      if (ww != 0.0) k = __ddiv_rn(k, ww);
      const double v = __dmul_rn(k, (double)(1 << FE_PRECISION_BITS));
      q = k < 0.0 ? (int)__dadd_rn(-0.5, v) : (int)__dadd_rn(0.5, v);
    }
    coeffs[(long long)x * out_size + o] = q;
  }
  bounds[o] = make_int2(xmin, xmax);
}

__device__ __forceinline__ int fe_clip8(int acc) {
  const int v = acc >> FE_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// ---- horizontal pass: in [rows, w_in, 3] u8 (row pitch in_pitch bytes, any alignment) -> out [rows, w_out, 3] u8 with
// a 16-byte aligned pitch.  One CTA per row.
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ in, long long in_pitch, long long in_bytes, int w_in,
                  uint8_t* __restrict__ out, long long out_pitch, int w_out, const int2* __restrict__ bounds,
                  const int* __restrict__ coeffs) {
  extern __shared__ uint4 fe_smem[];
  const int row = blockIdx.x;
  const int row_bytes = w_in * 3;
  const uint8_t* src = in + (long long)row * in_pitch;
  const int head = (int)(reinterpret_cast<uintptr_t>(src) & 15);  // stage from the enclosing 16-byte window
  const uint8_t* win = src - head;
  const int nvec = (head + row_bytes + 15) >> 4;
  uint8_t* s_in = reinterpret_cast<uint8_t*>(fe_smem);
  uint8_t* s_out = s_in + (((size_t)nvec * 16 + 15) & ~size_t(15));
  const uint8_t* buf_end = in + in_bytes;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint8_t* p = win + (long long)v * 16;
    if (p + 16 <= buf_end) {
      fe_smem[v] = *reinterpret_cast<const uint4*>(p);
    } else {  // last vector of the image: never read past the caller's buffer
      for (int b = 0; b < 16; ++b) s_in[v * 16 + b] = (p + b < buf_end) ? p[b] : (uint8_t)0;
    }
  }
  __syncthreads();
  const uint8_t* px = s_in + head;
  for (int o = threadIdx.x; o < w_out; o += blockDim.x) {
    const int2 b = bounds[o];
    int a0 = 1 << (FE_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    const uint8_t* p = px + b.x * 3;
    for (int t = 0; t < b.y; ++t) {
      const int k = coeffs[(long long)t * w_out + o];
      a0 += (int)p[0] * k; a1 += (int)p[1] * k; a2 += (int)p[2] * k;
      p += 3;
    }
    s_out[o * 3 + 0] = (uint8_t)fe_clip8(a0);
    s_out[o * 3 + 1] = (uint8_t)fe_clip8(a1);
    s_out[o * 3 + 2] = (uint8_t)fe_clip8(a2);
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(out + (long long)row * out_pitch);
  const uint4* s_out4 = reinterpret_cast<const uint4*>(s_out);
  const int ovec = (w_out * 3 + 15) >> 4;  // the pitch is a multiple of 16: whole vectors may be written
  for (int v = threadIdx.x; v < ovec; v += blockDim.x) dst[v] = s_out4[v];
}

// ---- vertical pass + pad / tile / normalise store.
// tmp [h_in, w*3 bytes] (pitch multiple of 16) -> resized rows oy in [0, h_out); pixel (oy, x) lands at canvas position
// (paste_y + oy, paste_x + x) of a canvas that is cut into 364 x 364 tiles, tiles_x per canvas row; tile t is written
// at out + t * tile_stride elements, planar [3][364][364].
// TILE = FE_CROP: the LLaVA crop grid with compile-time divisions; TILE = 0: tile_w_rt x tile_h_rt (Idefics2: one planar canvas)
template <typename OutT, int TILE>
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, long long pitch, int w, int h_out, const int2* __restrict__ bounds,
                  const int* __restrict__ coeffs, int paste_x, int paste_y, int tiles_x, int tile_w_rt, int tile_h_rt,
                  OutT* __restrict__ out, long long tile_stride, const FeLut lut) {
  const int tile_w = TILE ? TILE : tile_w_rt, tile_h = TILE ? TILE : tile_h_rt;
  __shared__ float s_lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i >> 8][i & 255] = lut.v[i >> 8][i & 255];
  __syncthreads();
  const int oy = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;  // 4-byte group within the row
  const int row_bytes = w * 3;
  if (g * 4 >= row_bytes) return;
  const int2 b = bounds[oy];
  int acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = 1 << (FE_PRECISION_BITS - 1);
  const uint8_t* p = tmp + (long long)b.x * pitch + (long long)g * 4;
  for (int t = 0; t < b.y; ++t) {
    const int k = coeffs[(long long)t * h_out + oy];
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    acc[0] += (int)(v & 255u) * k;
    acc[1] += (int)((v >> 8) & 255u) * k;
    acc[2] += (int)((v >> 16) & 255u) * k;
    acc[3] += (int)(v >> 24) * k;
    p += pitch;
  }
  const int cy = paste_y + oy;
  const int ty = cy / tile_h, iy = cy - ty * tile_h;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int byte = g * 4 + q;
    if (byte >= row_bytes) break;
    const int x = byte / 3, c = byte - x * 3;
    const int cx = paste_x + x;
    const int tx = cx / tile_w, ix = cx - tx * tile_w;
    OutT* o = out + (long long)(ty * tiles_x + tx) * tile_stride + ((long long)c * tile_h + iy) * tile_w + ix;
    *o = (OutT)s_lut[c][fe_clip8(acc[q])];
  }
}

// vertical pass with a uint8 result (no normalisation / tiling): used when the vertical pass has to run FIRST.
// tmp [h_in, w*3 bytes] (pitch multiple of 16) -> out [h_out, w*3 bytes] with the same pitch.
__global__ void __launch_bounds__(256)
resample_v_raw_kernel(const uint8_t* __restrict__ tmp, long long pitch, int w, int h_out, const int2* __restrict__ bounds,
                      const int* __restrict__ coeffs, uint8_t* __restrict__ out) {
  const int oy = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 4 >= w * 3) return;
  const int2 b = bounds[oy];
  int acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = 1 << (FE_PRECISION_BITS - 1);
  const uint8_t* p = tmp + (long long)b.x * pitch + (long long)g * 4;
  for (int t = 0; t < b.y; ++t) {
    const int k = coeffs[(long long)t * h_out + oy];
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    acc[0] += (int)(v & 255u) * k;
    acc[1] += (int)((v >> 8) & 255u) * k;
    acc[2] += (int)((v >> 16) & 255u) * k;
    acc[3] += (int)(v >> 24) * k;
    p += pitch;
  }
  const uint32_t r = (uint32_t)fe_clip8(acc[0]) | ((uint32_t)fe_clip8(acc[1]) << 8) | ((uint32_t)fe_clip8(acc[2]) << 16) |
                     ((uint32_t)fe_clip8(acc[3]) << 24);
  *reinterpret_cast<uint32_t*>(out + (long long)oy * pitch + (long long)g * 4) = r;
}

// canvas pixels outside the pasted rectangle get the normalised value of black
template <typename OutT, int TILE>
__global__ void __launch_bounds__(256)
pad_fill_kernel(int canvas_w, int canvas_h, int paste_x, int paste_y, int w, int h, int tiles_x, int tile_w_rt, int tile_h_rt,
                OutT* __restrict__ out, long long tile_stride, float pad0, float pad1, float pad2) {
  const int tile_w = TILE ? TILE : tile_w_rt, tile_h = TILE ? TILE : tile_h_rt;
  const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;
  if (cx >= canvas_w || cy >= canvas_h) return;
  if (cy >= paste_y && cy < paste_y + h && cx >= paste_x && cx < paste_x + w) return;
  const int ty = cy / tile_h, iy = cy - ty * tile_h, tx = cx / tile_w, ix = cx - tx * tile_w;
  OutT* o = out + (long long)(ty * tiles_x + tx) * tile_stride + (long long)iy * tile_w + ix;
  o[0] = (OutT)pad0;
  o[(long long)tile_w * tile_h] = (OutT)pad1;
  o[2LL * tile_w * tile_h] = (OutT)pad2;
}

static int fe_ksize(int in_size, int out_size, int filter = FE_BICUBIC) {
  if (in_size == out_size) return 1;
  double fs = (double)in_size / out_size;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil((filter == FE_BILINEAR ? 1.0 : 2.0) * fs) * 2 + 1;
}

static size_t fe_align(size_t n) { return (n + 255) & ~size_t(255); }

// Pillow >= 11 resamples vertically FIRST when the image is more than 100x taller than wide and shrinks vertically
// (Image.resize, Python side); the order is observable because the intermediate is rounded to uint8
static bool fe_vertical_first(int w_in, int h_in, int h_out) { return (long long)h_in > 100LL * w_in && h_out < h_in; }

// workspace for one resize (w_in, h_in) -> (w_out, h_out): intermediate(s) + coefficient tables
static size_t fe_resize_ws(int w_in, int h_in, int w_out, int h_out, int filter = FE_BICUBIC) {
  const size_t pitch = ((size_t)w_out * 3 + 15) & ~size_t(15);
  size_t n = fe_align(pitch * h_in) + fe_align((size_t)w_out * 8) + fe_align((size_t)w_out * fe_ksize(w_in, w_out, filter) * 4) +
             fe_align((size_t)h_out * 8) + fe_align((size_t)h_out * fe_ksize(h_in, h_out, filter) * 4);
  if (fe_vertical_first(w_in, h_in, h_out)) {  // + pitched copy of the source, vertical result, identity tables
    const size_t p0 = ((size_t)w_in * 3 + 15) & ~size_t(15);
    n += fe_align(p0 * h_in) + fe_align(p0 * h_out) + fe_align(pitch * h_out) + fe_align((size_t)w_in * 12) +
         fe_align((size_t)h_out * 12);
  }
  return n;
}

// where the resized image goes: a canvas cut into tile_w x tile_h planar tiles (LLaVA: 364 x 364 crops, padding = normalised black;
// Idefics2: one tile = the whole canvas, padding = 0.0 written after the normalisation)
struct FeTarget {
  int canvas_w, canvas_h, paste_x, paste_y, tile_w, tile_h, filter;
  float pad[3];
};

template <typename OutT>
static int fe_resize_into(const uint8_t* rgb, int H, int W, int w_out, int h_out, const FeTarget& tg, OutT* out, const FeLut& lut,
                          uint8_t* ws, cudaStream_t stream) {
  const int canvas_w = tg.canvas_w, canvas_h = tg.canvas_h, paste_x = tg.paste_x, paste_y = tg.paste_y, filter = tg.filter;
  const bool crop_tiles = tg.tile_w == FE_CROP && tg.tile_h == FE_CROP;
  const int kx = fe_ksize(W, w_out, filter), ky = fe_ksize(H, h_out, filter);
  LP_REQUIRE(kx <= FE_MAX_KSIZE && ky <= FE_MAX_KSIZE, "lp_image_frontend: downscale factor above 32 (%dx%d -> %dx%d)",
             W, H, w_out, h_out);
  const size_t pitch = ((size_t)w_out * 3 + 15) & ~size_t(15);
  if (fe_vertical_first(W, H, h_out)) {
    // rare order (aspect > 100): pitched copy of the source (identity horizontal pass) -> vertical pass to uint8 ->
    // horizontal pass -> identity vertical pass whose store applies normalisation / tiling.  Identity passes are exact.
    const size_t p0 = ((size_t)W * 3 + 15) & ~size_t(15);
    uint8_t* t0 = ws; ws += fe_align(p0 * H);
    uint8_t* t1 = ws; ws += fe_align(p0 * h_out);
    uint8_t* t2 = ws; ws += fe_align(pitch * h_out);
    int2* bxi = reinterpret_cast<int2*>(ws); ws += fe_align((size_t)W * 8);
    int* cxi = reinterpret_cast<int*>(ws); ws += fe_align((size_t)W * 4);
    int2* byi = reinterpret_cast<int2*>(ws); ws += fe_align((size_t)h_out * 8);
    int* cyi = reinterpret_cast<int*>(ws); ws += fe_align((size_t)h_out * 4);
    int2* bx = reinterpret_cast<int2*>(ws); ws += fe_align((size_t)w_out * 8);
    int* cxs = reinterpret_cast<int*>(ws); ws += fe_align((size_t)w_out * kx * 4);
    int2* by = reinterpret_cast<int2*>(ws); ws += fe_align((size_t)h_out * 8);
    int* cys = reinterpret_cast<int*>(ws);
    resample_coeff_kernel<<<(W + 127) / 128, 128, 0, stream>>>(W, W, 1, 1, filter, bxi, cxi);
    resample_coeff_kernel<<<(h_out + 127) / 128, 128, 0, stream>>>(h_out, h_out, 1, 1, filter, byi, cyi);
    resample_coeff_kernel<<<(w_out + 127) / 128, 128, 0, stream>>>(W, w_out, kx, W == w_out, filter, bx, cxs);
    resample_coeff_kernel<<<(h_out + 127) / 128, 128, 0, stream>>>(H, h_out, ky, 0, filter, by, cys);
    const size_t smem = 2 * p0 + pitch + 64;
    LP_REQUIRE(smem <= 48 * 1024, "lp_image_frontend: image row too wide for the vertical-first path (W=%d)", W);
    resample_h_kernel<<<H, 256, smem, stream>>>(rgb, (long long)W * 3, (long long)H * W * 3, W, t0, (long long)p0, W, bxi,
                                                cxi);
    const int g0 = (W * 3 + 3) / 4;
    resample_v_raw_kernel<<<dim3((g0 + 255) / 256, h_out), 256, 0, stream>>>(t0, (long long)p0, W, h_out, by, cys, t1);
    resample_h_kernel<<<h_out, 256, smem, stream>>>(t1, (long long)p0, (long long)p0 * h_out, W, t2, (long long)pitch,
                                                    w_out, bx, cxs);
    const int tiles_x1 = canvas_w / tg.tile_w;
    const long long tile_stride1 = 3LL * tg.tile_w * tg.tile_h;
    if (w_out != canvas_w || h_out != canvas_h) {
      if (crop_tiles) {
        pad_fill_kernel<OutT, FE_CROP><<<dim3((canvas_w + 255) / 256, canvas_h), 256, 0, stream>>>(
            canvas_w, canvas_h, paste_x, paste_y, w_out, h_out, tiles_x1, tg.tile_w, tg.tile_h, out, tile_stride1, tg.pad[0],
            tg.pad[1], tg.pad[2]);
      } else {
        pad_fill_kernel<OutT, 0><<<dim3((canvas_w + 255) / 256, canvas_h), 256, 0, stream>>>(
            canvas_w, canvas_h, paste_x, paste_y, w_out, h_out, tiles_x1, tg.tile_w, tg.tile_h, out, tile_stride1, tg.pad[0],
            tg.pad[1], tg.pad[2]);
      }
      note_launch();
    }
    const int g2 = (w_out * 3 + 3) / 4;
    if (crop_tiles) {
      resample_v_kernel<OutT, FE_CROP><<<dim3((g2 + 255) / 256, h_out), 256, 0, stream>>>(
          t2, (long long)pitch, w_out, h_out, byi, cyi, paste_x, paste_y, tiles_x1, tg.tile_w, tg.tile_h, out, tile_stride1, lut);
    } else {
      resample_v_kernel<OutT, 0><<<dim3((g2 + 255) / 256, h_out), 256, 0, stream>>>(
          t2, (long long)pitch, w_out, h_out, byi, cyi, paste_x, paste_y, tiles_x1, tg.tile_w, tg.tile_h, out, tile_stride1, lut);
    }
    note_launch(8);
    LP_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  uint8_t* tmp = ws; ws += fe_align(pitch * H);
  int2* bx = reinterpret_cast<int2*>(ws); ws += fe_align((size_t)w_out * 8);
  int* cxs = reinterpret_cast<int*>(ws); ws += fe_align((size_t)w_out * kx * 4);
  int2* by = reinterpret_cast<int2*>(ws); ws += fe_align((size_t)h_out * 8);
  int* cys = reinterpret_cast<int*>(ws);
  resample_coeff_kernel<<<(w_out + 127) / 128, 128, 0, stream>>>(W, w_out, kx, W == w_out, filter, bx, cxs);
  resample_coeff_kernel<<<(h_out + 127) / 128, 128, 0, stream>>>(H, h_out, ky, H == h_out, filter, by, cys);
  const size_t smem = (((size_t)W * 3 + 15 + 15) / 16 * 16 + 15) / 16 * 16 + pitch + 32;
  LP_REQUIRE(smem <= 200 * 1024, "lp_image_frontend: image row too wide for the staging buffer (W=%d)", W);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    LP_CHECK_CUDA(cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  resample_h_kernel<<<H, 256, smem, stream>>>(rgb, (long long)W * 3, (long long)H * W * 3, W, tmp, (long long)pitch,
                                              w_out, bx, cxs);
  const int tiles_x = canvas_w / tg.tile_w;
  const long long tile_stride = 3LL * tg.tile_w * tg.tile_h;
  if (w_out != canvas_w || h_out != canvas_h) {
    if (crop_tiles) {
      pad_fill_kernel<OutT, FE_CROP><<<dim3((canvas_w + 255) / 256, canvas_h), 256, 0, stream>>>(
          canvas_w, canvas_h, paste_x, paste_y, w_out, h_out, tiles_x, tg.tile_w, tg.tile_h, out, tile_stride, tg.pad[0], tg.pad[1],
          tg.pad[2]);
    } else {
      pad_fill_kernel<OutT, 0><<<dim3((canvas_w + 255) / 256, canvas_h), 256, 0, stream>>>(
          canvas_w, canvas_h, paste_x, paste_y, w_out, h_out, tiles_x, tg.tile_w, tg.tile_h, out, tile_stride, tg.pad[0], tg.pad[1],
          tg.pad[2]);
    }
    note_launch();
  }
  const int groups = (w_out * 3 + 3) / 4;
  if (crop_tiles) {
    resample_v_kernel<OutT, FE_CROP><<<dim3((groups + 255) / 256, h_out), 256, 0, stream>>>(
        tmp, (long long)pitch, w_out, h_out, by, cys, paste_x, paste_y, tiles_x, tg.tile_w, tg.tile_h, out, tile_stride, lut);
  } else {
    resample_v_kernel<OutT, 0><<<dim3((groups + 255) / 256, h_out), 256, 0, stream>>>(
        tmp, (long long)pitch, w_out, h_out, by, cys, paste_x, paste_y, tiles_x, tg.tile_w, tg.tile_h, out, tile_stride, lut);
  }
  note_launch(4);
  LP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// resize_and_pad_image's geometry (EV/...anyres.py:116-136), same double arithmetic as the Python original
static void fe_pad_geometry(int w, int h, int tw, int th, int* nw, int* nh, int* px, int* py) {
  const double scale_w = (double)tw / (double)w, scale_h = (double)th / (double)h;
  if (scale_w < scale_h) {
    *nw = tw;
    const int c = (int)ceil((double)h * scale_w);
    *nh = c < th ? c : th;
  } else {
    *nh = th;
    const int c = (int)ceil((double)w * scale_h);
    *nw = c < tw ? c : tw;
  }
  *px = (tw - *nw) / 2;
  *py = (th - *nh) / 2;
}

}  // namespace lp

using namespace lp;

extern "C" size_t lp_image_frontend_workspace_bytes(int H, int W, int target_w, int target_h) {
  size_t n = fe_resize_ws(W, H, FE_CROP, FE_CROP);
  if (target_w > 0 && target_h > 0) {
    int nw, nh, px, py;
    fe_pad_geometry(W, H, target_w, target_h, &nw, &nh, &px, &py);
    n += fe_resize_ws(W, H, nw, nh);
  }
  return n + 256;
}

extern "C" int lp_image_frontend(const void* rgb, int H, int W, int target_w, int target_h, void* out_global,
                                 void* out_tiles, int out_fp32, double rescale, const float* mean, const float* std_,
                                 void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(rgb && out_global && workspace && mean && std_, "lp_image_frontend: null argument");
  LP_REQUIRE(H > 0 && W > 0 && W <= 16384 && H <= 16384, "lp_image_frontend: bad image size %dx%d", W, H);
  LP_REQUIRE((reinterpret_cast<uintptr_t>(rgb) & 15) == 0, "lp_image_frontend: image must be 16-byte aligned");
  const bool tiles = target_w > 0 && target_h > 0;
  LP_REQUIRE(!tiles || (out_tiles && target_w % FE_CROP == 0 && target_h % FE_CROP == 0),
             "lp_image_frontend: target resolution %dx%d is not a grid of 364-pixel tiles (or out_tiles is null)",
             target_w, target_h);
  LP_REQUIRE(ws_bytes >= lp_image_frontend_workspace_bytes(H, W, target_w, target_h),
             "lp_image_frontend: workspace too small");
  FeLut lut;  // SiglipImageProcessor: rescale in float64 rounded to float32, then (x - mean) / std in float32
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < 256; ++i) {
      const float x = (float)((double)i * rescale);
      lut.v[c][i] = (x - mean[c]) / std_[c];
    }
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const uint8_t* img = reinterpret_cast<const uint8_t*>(rgb);
  ProfScope prof(14, H, W, tiles ? (target_w / FE_CROP) * (target_h / FE_CROP) : 0, stream);
  int rc;
  // the original image -> one 364 x 364 crop (aspect ratio not preserved, like the processor's resize)
  FeTarget tg = {FE_CROP, FE_CROP, 0, 0, FE_CROP, FE_CROP, FE_BICUBIC, {lut.v[0][0], lut.v[1][0], lut.v[2][0]}};
  if (out_fp32)
    rc = fe_resize_into<float>(img, H, W, FE_CROP, FE_CROP, tg, reinterpret_cast<float*>(out_global), lut, ws, stream);
  else
    rc = fe_resize_into<bf16>(img, H, W, FE_CROP, FE_CROP, tg, reinterpret_cast<bf16*>(out_global), lut, ws, stream);
  if (rc || !tiles) return rc;
  ws += fe_resize_ws(W, H, FE_CROP, FE_CROP);
  int nw, nh, px, py;
  fe_pad_geometry(W, H, target_w, target_h, &nw, &nh, &px, &py);
  tg.canvas_w = target_w; tg.canvas_h = target_h; tg.paste_x = px; tg.paste_y = py;
  if (out_fp32) return fe_resize_into<float>(img, H, W, nw, nh, tg, reinterpret_cast<float*>(out_tiles), lut, ws, stream);
  return fe_resize_into<bf16>(img, H, W, nw, nh, tg, reinterpret_cast<bf16*>(out_tiles), lut, ws, stream);
}

// ---- Idefics2ImageProcessor for one image (evaluations/models/idefics2_multiimg.py:22-25 configures it with
// size = {'longest_edge': R, 'shortest_edge': 0}, do_image_splitting = False): BILINEAR `Image.resize` to (out_w, out_h) — the
// caller plans the size with the processor's own integer rule —, x * rescale in float64 rounded to float32, (x - mean) / std in
// float32, pasted at the top-left corner of a planar [3, canvas_h, canvas_w] canvas whose remaining pixels are `pad_value`
// (the processor pads with 0.0 AFTER the normalisation) and `mask` [canvas_h, canvas_w] int64 = 1 on the image, 0 on the padding.
__global__ void __launch_bounds__(256) pixel_mask_kernel(long long* __restrict__ mask, int canvas_w, int canvas_h, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < canvas_w && y < canvas_h) mask[(long long)y * canvas_w + x] = (x < w && y < h) ? 1LL : 0LL;
}

extern "C" size_t lp_image_resize_workspace_bytes(int H, int W, int out_w, int out_h, int filter) {
  return fe_resize_ws(W, H, out_w, out_h, filter) + 256;
}

extern "C" int lp_image_resize(const void* rgb, int H, int W, int out_w, int out_h, int filter, int canvas_w, int canvas_h,
                               float pad_value, void* out, int out_fp32, long long* mask, double rescale, const float* mean,
                               const float* std_, void* workspace, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LP_REQUIRE(rgb && out && workspace && mean && std_, "lp_image_resize: null argument");
  LP_REQUIRE(H > 0 && W > 0 && W <= 16384 && H <= 16384, "lp_image_resize: bad image size %dx%d", W, H);
  LP_REQUIRE(out_w > 0 && out_h > 0 && canvas_w >= out_w && canvas_h >= out_h, "lp_image_resize: %dx%d does not fit the %dx%d canvas",
             out_w, out_h, canvas_w, canvas_h);
  LP_REQUIRE(filter == FE_BICUBIC || filter == FE_BILINEAR, "lp_image_resize: filter must be 0 (bicubic) or 1 (bilinear)");
  LP_REQUIRE((reinterpret_cast<uintptr_t>(rgb) & 15) == 0, "lp_image_resize: image must be 16-byte aligned");
  LP_REQUIRE(ws_bytes >= lp_image_resize_workspace_bytes(H, W, out_w, out_h, filter), "lp_image_resize: workspace too small");
  FeLut lut;
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < 256; ++i) {
      const float x = (float)((double)i * rescale);
      lut.v[c][i] = (x - mean[c]) / std_[c];
    }
  FeTarget tg = {canvas_w, canvas_h, 0, 0, canvas_w, canvas_h, filter, {pad_value, pad_value, pad_value}};
  ProfScope prof(14, H, W, 0, stream);
  int rc;
  if (out_fp32)
    rc = fe_resize_into<float>(reinterpret_cast<const uint8_t*>(rgb), H, W, out_w, out_h, tg, reinterpret_cast<float*>(out), lut,
                               reinterpret_cast<uint8_t*>(workspace), stream);
  else
    rc = fe_resize_into<bf16>(reinterpret_cast<const uint8_t*>(rgb), H, W, out_w, out_h, tg, reinterpret_cast<bf16*>(out), lut,
                              reinterpret_cast<uint8_t*>(workspace), stream);
  if (rc) return rc;
  if (mask != nullptr) {
    pixel_mask_kernel<<<dim3((canvas_w + 255) / 256, canvas_h), 256, 0, stream>>>(mask, canvas_w, canvas_h, out_w, out_h);
    note_launch();
    LP_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
