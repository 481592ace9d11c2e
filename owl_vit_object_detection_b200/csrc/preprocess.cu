// Image preprocessing on the device: what the reference does per item on the CPU in its DataLoader
// workers (src/dataset.py:64-71 -> HF OwlViTImageProcessor, transformers 4.30.2: PIL bicubic resize to 768 x 768,
// rescale by 1/255, CLIP mean / std, channels first) as three kernels per image on raw uint8 HWC pixels, so that the
// host ships 1 byte per channel instead of a 7 MB fp32 tensor.
//
// Bit-exactness against Pillow (oracle/preprocess_oracle.py): the resample is Pillow's own algorithm
// (libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc, PRECISION_BITS = 22, horizontal pass then
// vertical pass, each rounded to uint8), the coefficients are computed in IEEE double with separately rounded
// operations (__dmul_rn / __dadd_rn / __ddiv_rn: the C code is compiled without FMA contraction), everything after
// that is integer arithmetic; rescale + normalise is a 3 x 256 table computed by the host with the reference's op
// sequence, so the float output is bit-identical too.
#include "common.h"
#include <cuda_runtime.h>
#include <math.h>

namespace owl {

constexpr int PRE_BITS = 22;   // Pillow: PRECISION_BITS = 32 - 8 - 2

__device__ __forceinline__ double pre_bicubic(double x) {
  // Resample.c bicubic_filter, a = -0.5:  ((a + 2) x - (a + 3)) x x + 1   |   (((x - 5) x + 8) x - 4) a
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// One thread per output coordinate of one axis: bounds[xx] = (first source index, tap count), kk[xx][0..ksize) int32.
__global__ void pre_coeff_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
  pdl_grid_wait();
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = __ddiv_rn(static_cast<double>(static_cast<float>(in_size)), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale));
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  int* k = kk + 1LL * xx * ksize;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x)
    ww = __dadd_rn(ww, pre_bicubic(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss)));
  for (int x = 0; x < ksize; ++x) {
    int v = 0;
    if (x < xmax) {
      double w = pre_bicubic(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      const double s = __dmul_rn(w, static_cast<double>(1 << PRE_BITS));
      v = w < 0 ? static_cast<int>(__dadd_rn(-0.5, s)) : static_cast<int>(__dadd_rn(0.5, s));
    }
    k[x] = v;
  }
  bounds[2 * xx] = xmin;
  bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ int pre_clip8(int v) {
  v >>= PRE_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Horizontal pass: img [H, W, 3] u8 (row stride in bytes) -> tmp [H, S, 3] u8.  Thread = one output pixel.
__global__ void pre_horizontal_kernel(const uint8_t* __restrict__ img, long long row_stride, int H, int S, int ksize,
                                      const int* __restrict__ bounds, const int* __restrict__ kk,
                                      uint8_t* __restrict__ tmp) {
  pdl_grid_wait();
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xx >= S) return;
  const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
  const int* k = kk + 1LL * xx * ksize;
  const uint8_t* src = img + y * row_stride + 3LL * xmin;
  int s0 = 1 << (PRE_BITS - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < xmax; ++x) {
    const int c = k[x];
    s0 += src[3 * x] * c;
    s1 += src[3 * x + 1] * c;
    s2 += src[3 * x + 2] * c;
  }
  uint8_t* d = tmp + (1LL * y * S + xx) * 3;
  d[0] = static_cast<uint8_t>(pre_clip8(s0));
  d[1] = static_cast<uint8_t>(pre_clip8(s1));
  d[2] = static_cast<uint8_t>(pre_clip8(s2));
}

// Vertical pass + rescale / normalise table: tmp [H, S, 3] u8 -> out [3, S, S] f32.  Thread = one output pixel.
__global__ void pre_vertical_kernel(const uint8_t* __restrict__ tmp, int S, int ksize, const int* __restrict__ bounds,
                                    const int* __restrict__ kk, const float* __restrict__ lut /*[3][256]*/,
                                    float* __restrict__ out) {
  pdl_grid_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y;
  if (x >= S) return;
  const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
  const int* k = kk + 1LL * yy * ksize;
  int s0 = 1 << (PRE_BITS - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < ymax; ++y) {
    const int c = k[y];
    const uint8_t* s = tmp + (1LL * (ymin + y) * S + x) * 3;
    s0 += s[0] * c;
    s1 += s[1] * c;
    s2 += s[2] * c;
  }
  const long long plane = 1LL * S * S, o = 1LL * yy * S + x;
  out[o] = __ldg(lut + pre_clip8(s0));
  out[plane + o] = __ldg(lut + 256 + pre_clip8(s1));
  out[2 * plane + o] = __ldg(lut + 512 + pre_clip8(s2));
}

static int pre_ksize(int in_size, int out_size) {
  const double scale = static_cast<double>(static_cast<float>(in_size)) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  return static_cast<int>(ceil(2.0 * filterscale)) * 2 + 1;
}

}  // namespace owl

using namespace owl;

extern "C" long long owl_preprocess_workspace_bytes(int H, int W, int out_size) {
  if (H <= 0 || W <= 0 || out_size <= 0) return -1;
  const long long kh = pre_ksize(W, out_size), kv = pre_ksize(H, out_size);
  const long long ints = 1LL * out_size * (kh + kv + 4);
  return ((ints * 4 + 255) & ~255LL) + 3LL * H * out_size;
}

extern "C" int owl_preprocess_image(const uint8_t* img_hwc, int H, int W, long long row_stride_bytes, const float* lut,
                                    float* out_chw, int out_size, void* workspace, long long workspace_bytes,
                                    void* stream) {
  OWL_CHECK_ARG(img_hwc && lut && out_chw && workspace, "preprocess_image: null argument");
  OWL_CHECK_ARG(H > 0 && W > 0 && out_size > 0 && row_stride_bytes >= 3LL * W, "preprocess_image: bad geometry");
  OWL_CHECK_ARG(workspace_bytes >= owl_preprocess_workspace_bytes(H, W, out_size),
                "preprocess_image: workspace of %lld bytes is too small", workspace_bytes);
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "preprocess_image: workspace must be 16-byte aligned");
  const int S = out_size, kh = pre_ksize(W, S), kv = pre_ksize(H, S);
  int* bounds_h = static_cast<int*>(workspace);
  int* bounds_v = bounds_h + 2 * S;
  int* kk_h = bounds_v + 2 * S;
  int* kk_v = kk_h + 1LL * S * kh;
  const long long ints = 1LL * S * (kh + kv + 4);
  uint8_t* tmp = static_cast<uint8_t*>(workspace) + ((ints * 4 + 255) & ~255LL);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OWL_LAUNCH(pre_coeff_kernel, (S + 127) / 128, 128, 0, s, W, S, kh, bounds_h, kk_h);
  OWL_LAUNCH(pre_coeff_kernel, (S + 127) / 128, 128, 0, s, H, S, kv, bounds_v, kk_v);
  OWL_LAUNCH(pre_horizontal_kernel, dim3((S + 127) / 128, H), 128, 0, s, img_hwc, row_stride_bytes, H, S, kh, bounds_h,
             kk_h, tmp);
  OWL_LAUNCH(pre_vertical_kernel, dim3((S + 127) / 128, S), 128, 0, s, tmp, S, kv, bounds_v, kk_v, lut, out_chw);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
