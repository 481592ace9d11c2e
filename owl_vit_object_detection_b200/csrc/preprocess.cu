// Image preprocessing on the device: what the reference does per item on the CPU in its DataLoader
// workers (src/dataset.py:64-71 -> HF OwlViTImageProcessor, transformers 4.30.2: PIL bicubic resize to 768 x 768,
// rescale by 1/255, CLIP mean / std, channels first) as three kernels per BATCH of ragged raw uint8 HWC images, so that the
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

// A batch of up to PRE_MAX_BATCH images of DIFFERENT sizes goes through three launches (coefficients of both axes of
// every image, horizontal pass, vertical pass): the per-image geometry travels by value in the kernel parameters.
constexpr int PRE_MAX_BATCH = OWL_PRE_MAX_BATCH;
struct PreItem {
  const uint8_t* px;      // [H, W, 3] u8, row stride in bytes
  long long stride;
  int* coeff;             // bounds_h [2S] | bounds_v [2S] | kk_h [S * kh] | kk_v [S * kv]
  uint8_t* tmp;           // [H, S, 3] u8 (result of the horizontal pass)
  int H, W, kh, kv;
};
struct PreBatch { PreItem it[PRE_MAX_BATCH]; };

// One thread per output coordinate of one axis of one image (blockIdx.y = axis, blockIdx.z = image):
// bounds[xx] = (first source index, tap count), kk[xx][0..ksize) int32.
__global__ void pre_coeff_kernel(const __grid_constant__ PreBatch batch, int out_size) {
  pdl_grid_wait();
  const PreItem& im = batch.it[blockIdx.z];
  const bool vertical = blockIdx.y != 0;
  const int in_size = vertical ? im.H : im.W, ksize = vertical ? im.kv : im.kh;
  int* bounds = im.coeff + (vertical ? 2 * out_size : 0);
  int* kk = im.coeff + 4 * out_size + (vertical ? 1LL * out_size * im.kh : 0);
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

// Horizontal pass: img [H, W, 3] u8 -> tmp [H, S, 3] u8.  Thread = one output pixel; blockIdx.y = source row (rows
// beyond this image's height leave), blockIdx.z = image.
__global__ void pre_horizontal_kernel(const __grid_constant__ PreBatch batch, int S) {
  pdl_grid_wait();
  const PreItem& im = batch.it[blockIdx.z];
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xx >= S || y >= im.H) return;
  const int* bounds = im.coeff;
  const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
  const int* k = im.coeff + 4 * S + 1LL * xx * im.kh;
  const uint8_t* src = im.px + y * im.stride + 3LL * xmin;
  int s0 = 1 << (PRE_BITS - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < xmax; ++x) {
    const int c = k[x];
    s0 += src[3 * x] * c;
    s1 += src[3 * x + 1] * c;
    s2 += src[3 * x + 2] * c;
  }
  uint8_t* d = im.tmp + (1LL * y * S + xx) * 3;
  d[0] = static_cast<uint8_t>(pre_clip8(s0));
  d[1] = static_cast<uint8_t>(pre_clip8(s1));
  d[2] = static_cast<uint8_t>(pre_clip8(s2));
}

// Vertical pass + rescale / normalise table: tmp [H, S, 3] u8 -> out[image] [3, S, S] f32.  Thread = one output pixel.
__global__ void pre_vertical_kernel(const __grid_constant__ PreBatch batch, int S, const float* __restrict__ lut /*[3][256]*/,
                                    float* __restrict__ out_all, long long out_image_stride) {
  pdl_grid_wait();
  const PreItem& im = batch.it[blockIdx.z];
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y;
  if (x >= S) return;
  const int* bounds = im.coeff + 2 * S;
  const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
  const int* k = im.coeff + 4 * S + 1LL * S * im.kh + 1LL * yy * im.kv;
  int s0 = 1 << (PRE_BITS - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < ymax; ++y) {
    const int c = k[y];
    const uint8_t* s = im.tmp + (1LL * (ymin + y) * S + x) * 3;
    s0 += s[0] * c;
    s1 += s[1] * c;
    s2 += s[2] * c;
  }
  float* out = out_all + blockIdx.z * out_image_stride;
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

static long long pre_image_bytes(int H, int W, int out_size) {
  const long long kh = pre_ksize(W, out_size), kv = pre_ksize(H, out_size);
  const long long ints = 1LL * out_size * (kh + kv + 4);
  return ((ints * 4 + 255) & ~255LL) + ((3LL * H * out_size + 255) & ~255LL);
}

extern "C" long long owl_preprocess_workspace_bytes(int H, int W, int out_size) {
  if (H <= 0 || W <= 0 || out_size <= 0) return -1;
  return pre_image_bytes(H, W, out_size);
}

extern "C" long long owl_preprocess_batch_workspace_bytes(const owl_pre_image* images_host, int n, int out_size) {
  if (!images_host || n <= 0 || out_size <= 0) return -1;
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    if (images_host[i].H <= 0 || images_host[i].W <= 0) return -1;
    total += pre_image_bytes(images_host[i].H, images_host[i].W, out_size);
  }
  return total;
}

extern "C" int owl_preprocess_batch(const owl_pre_image* images_host, int n, const float* lut, float* out_nchw,
                                    int out_size, void* workspace, long long workspace_bytes, void* stream) {
  OWL_CHECK_ARG(images_host && lut && out_nchw && workspace && n > 0 && out_size > 0, "preprocess_batch: null / empty argument");
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "preprocess_batch: workspace must be 16-byte aligned");
  OWL_CHECK_ARG(workspace_bytes >= owl_preprocess_batch_workspace_bytes(images_host, n, out_size),
                "preprocess_batch: workspace of %lld bytes is too small (or an image has an empty side)", workspace_bytes);
  const int S = out_size;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  for (int i0 = 0; i0 < n; i0 += PRE_MAX_BATCH) {
    const int nb = n - i0 < PRE_MAX_BATCH ? n - i0 : PRE_MAX_BATCH;
    PreBatch batch = {};
    int hmax = 0;
    for (int j = 0; j < nb; ++j) {
      const owl_pre_image& src = images_host[i0 + j];
      OWL_CHECK_ARG(src.pixels && src.row_stride_bytes >= 3LL * src.W, "preprocess_batch: image %d: bad geometry", i0 + j);
      PreItem& it = batch.it[j];
      it.px = src.pixels;
      it.stride = src.row_stride_bytes;
      it.H = src.H;
      it.W = src.W;
      it.kh = pre_ksize(src.W, S);
      it.kv = pre_ksize(src.H, S);
      const long long ints = 1LL * S * (it.kh + it.kv + 4);
      it.coeff = reinterpret_cast<int*>(ws);
      it.tmp = ws + ((ints * 4 + 255) & ~255LL);
      ws += pre_image_bytes(src.H, src.W, S);
      hmax = src.H > hmax ? src.H : hmax;
    }
    OWL_CHECK_ARG(hmax <= 65535 && S <= 65535, "preprocess_batch: image sides above 65535 are not supported");
    const unsigned gx = (S + 127) / 128;
    OWL_LAUNCH(pre_coeff_kernel, dim3(gx, 2, nb), 128, 0, s, batch, S);
    OWL_LAUNCH(pre_horizontal_kernel, dim3(gx, hmax, nb), 128, 0, s, batch, S);
    OWL_LAUNCH(pre_vertical_kernel, dim3(gx, S, nb), 128, 0, s, batch, S, lut, out_nchw + 3LL * S * S * i0,
               3LL * S * S);
  }
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_preprocess_image(const uint8_t* img_hwc, int H, int W, long long row_stride_bytes, const float* lut,
                                    float* out_chw, int out_size, void* workspace, long long workspace_bytes,
                                    void* stream) {
  OWL_CHECK_ARG(img_hwc && lut && out_chw && workspace, "preprocess_image: null argument");
  OWL_CHECK_ARG(H > 0 && W > 0 && out_size > 0 && row_stride_bytes >= 3LL * W, "preprocess_image: bad geometry");
  const owl_pre_image one = {img_hwc, H, W, row_stride_bytes};
  return owl_preprocess_batch(&one, 1, lut, out_chw, out_size, workspace, workspace_bytes, stream);
}
