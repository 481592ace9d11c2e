// HBM-bound forward kernels around the GEMMs: patch gather + cast, LayerNorm, the fused
// post-LN x CLS x LN tail, class-head normalisation, box-head tail, softmax rows, casts.
// One warp per row, 128-bit loads, fp32 statistics; all grids are sized from the row count.
#include "common.h"
#include <algorithm>
#include <cuda_fp16.h>

namespace owl {

constexpr int ROW_WARPS = 8;       // warps (rows) per CTA
constexpr int MAX_VEC = 8;         // row length <= 8 * 128 floats

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ------------------------------------------------------------------ patch gather (HF:336 conv k = s = patch)
// img [B,3,IS,IS] fp32 NCHW -> patches [B*g*g, ld] fp16, column = c*ps*ps + ky*ps + kx (conv-weight order).
template <int VEC>
__global__ void im2col_kernel(const float* __restrict__ img, __half* __restrict__ out, int B, int IS, int ps,
                              int ld) {
  pdl_grid_wait();
  const int g = IS / ps;
  const long long groups_per_row = IS / VEC;
  const long long total = 1LL * B * 3 * IS * groups_per_row;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int xg = static_cast<int>(i % groups_per_row);
    long long r = i / groups_per_row;
    const int y = static_cast<int>(r % IS);
    r /= IS;
    const int c = static_cast<int>(r % 3);
    const int b = static_cast<int>(r / 3);
    const int x = xg * VEC;
    const float* src = img + ((1LL * b * 3 + c) * IS + y) * IS + x;
    const int py = y / ps, ky = y - py * ps, px = x / ps, kx = x - px * ps;
    __half* dst = out + (1LL * b * g * g + py * g + px) * ld + (c * ps + ky) * ps + kx;
    if constexpr (VEC == 8) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
      __half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src));
      *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(v.x, v.y);
    }
  }
}

// ------------------------------------------------------------------ raw pixels -> patch rows (one pass)
// img [B,IS,IS,3] uint8 RGB (HWC, as decoded) -> patches [B*g*g, ld] fp16, same column order as im2col_kernel.
// lut [3][256] fp32 = the value byte v of channel c takes after the reference's rescale (/255) + CLIP normalise
// (src/dataset.py:64-71 -> HF OwlViTImageProcessor; built on the host, preprocess.py).  For an image that already
// has the model's resolution PIL's resize is the identity, so this kernel alone is the whole reference
// preprocessing followed by the fp16 cast of im2col_kernel: bit-identical patches from a quarter of the bytes.
// One thread = 8 consecutive pixels of a row: 24 contiguous bytes in, three 16-byte stores out (one per channel).
__global__ void u8_patches_kernel(const uint8_t* __restrict__ img, const float* __restrict__ lut,
                                  __half* __restrict__ out, int B, int IS, int ps, int ld) {
  __shared__ float slut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) slut[i] = lut[i];
  pdl_grid_wait();
  __syncthreads();
  const int g = IS / ps;
  const int groups_per_row = IS / 8;
  const long long total = 1LL * B * IS * groups_per_row;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int xg = static_cast<int>(i % groups_per_row);
    long long r = i / groups_per_row;
    const int y = static_cast<int>(r % IS);
    const int b = static_cast<int>(r / IS);
    const int x = xg * 8;
    const uint2* src = reinterpret_cast<const uint2*>(img + ((1LL * b * IS + y) * IS + x) * 3);   // 24 B, 8-aligned
    const uint2 w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    const uint32_t w[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
    const int py = y / ps, ky = y - py * ps, px = x / ps, kx = x - px * ps;
    __half* dst = out + (1LL * b * g * g + py * g + px) * ld + ky * ps + kx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t hh[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i0 = (2 * k) * 3 + c, i1 = (2 * k + 1) * 3 + c;          // byte index inside the 24
        const uint32_t v0 = (w[i0 >> 2] >> ((i0 & 3) * 8)) & 0xffu, v1 = (w[i1 >> 2] >> ((i1 & 3) * 8)) & 0xffu;
        const __half2 h = __floats2half2_rn(slut[c * 256 + v0], slut[c * 256 + v1]);
        hh[k] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(dst + 1LL * c * ps * ps) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm (HF:498,507,768; eps 1e-5)
// y[r] = LN(x[r]) * gamma + beta.  Row r is read at x + r * x_stride (so a strided subset of rows, e.g.
// the CLS rows, can be normalised).  If `cls_emb` is set, rows with r % tokens == 0 take their input
// from cls_emb + pos[0] instead (the CLS row of the embedding, HF:338-343).
// NV = D / 128 is a template parameter so that the row's loads are issued back to back with nothing between
// them (the SM issues in order: a consumer or a branch between two loads exposes a full memory latency per load).
template <typename OutT, int NV>
__global__ void layernorm_kernel(const float* __restrict__ x, long long x_stride, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, OutT* __restrict__ y, long long y_stride,
                                 int rows, float eps, const float* __restrict__ cls_emb,
                                 const float* __restrict__ pos0, int tokens) {
  constexpr int D = NV * 128;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_grid_wait();
  if (row >= rows) return;
  float4 v[NV], g[NV], be[NV];
  const bool is_cls = cls_emb != nullptr && (row % tokens) == 0;
  const float* xr = x + row * x_stride;
  if (!is_cls) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(cls_emb + i * 128 + lane * 4));
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos0 + i * 128 + lane * 4));
      v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {   // in flight behind the row itself; first used after both reductions
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4));
    be[i] = __ldg(reinterpret_cast<const float4*>(beta + i * 128 + lane * 4));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  OutT* yr = y + row * y_stride;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 o;
    o.x = (v[i].x - mean) * rstd * g[i].x + be[i].x;
    o.y = (v[i].y - mean) * rstd * g[i].y + be[i].y;
    o.z = (v[i].z - mean) * rstd * g[i].z + be[i].z;
    o.w = (v[i].w - mean) * rstd * g[i].w + be[i].w;
    store4(yr + i * 128 + lane * 4, o);
  }
}

// ------------------------------------------------------------------ reference src/models.py:80-86 fused
// feats[b,p] = LN2( LN1(x[b,1+p]) * ecls[b] )  with ecls[b] = LN1(x[b,0]) precomputed (fp32 [B,D]).
template <int NV>
__global__ void post_fuse_kernel(const float* __restrict__ x, const float* __restrict__ ecls,
                                 const float* __restrict__ g1, const float* __restrict__ b1,
                                 const float* __restrict__ g2, const float* __restrict__ b2,
                                 __half* __restrict__ feats, int B, int P, float eps) {
  constexpr int D = NV * 128;
  pdl_grid_wait();
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= B * P) return;
  const int lane = threadIdx.x & 31;
  const int b = row / P, p = row - b * P;
  const float* xr = x + (1LL * b * (P + 1) + 1 + p) * D;
  const float* cr = ecls + 1LL * b * D;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + bb * bb) + (c * c + d * d);
  }
  float rstd = rsqrtf(warp_sum(q) / D + eps);
  s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(g1 + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(b1 + c));
    const float4 cl = __ldg(reinterpret_cast<const float4*>(cr + c));
    v[i].x = ((v[i].x - mean) * rstd * g.x + be.x) * cl.x;
    v[i].y = ((v[i].y - mean) * rstd * g.y + be.y) * cl.y;
    v[i].z = ((v[i].z - mean) * rstd * g.z + be.z) * cl.z;
    v[i].w = ((v[i].w - mean) * rstd * g.w + be.w) * cl.w;
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  mean = warp_sum(s) / D;
  q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + bb * bb) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) / D + eps);
  __half* fr = feats + 1LL * row * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(g2 + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(b2 + c));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + be.x;
    o.y = (v[i].y - mean) * rstd * g.y + be.y;
    o.z = (v[i].z - mean) * rstd * g.z + be.z;
    o.w = (v[i].w - mean) * rstd * g.w + be.w;
    store4(fr + c, o);
  }
}

// ------------------------------------------------------------------ class head normalisations
// reference src/models.py:28-30  e / (||e|| + 1e-6)      (mode 0, image side)
// reference src/models.py:31-33  q / ||q|| + 1e-6        (mode 1, query side; precedence quirk Q1)
__global__ void rownorm_kernel(const float* __restrict__ e, __half* __restrict__ out, int rows, int E, int mode) {
  pdl_grid_wait();
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = E >> 7;
  const float* er = e + 1LL * row * E;
  float4 v[MAX_VEC];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i)
    if (i < nv) {
      v[i] = *reinterpret_cast<const float4*>(er + i * 128 + lane * 4);
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  const float nrm = sqrtf(warp_sum(q));
  __half* orow = out + 1LL * row * E;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i)
    if (i < nv) {
      float4 o;
      if (mode == 0) {
        const float inv = 1.0f / (nrm + 1e-6f);
        o = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
      } else {
        o = make_float4(v[i].x / nrm + 1e-6f, v[i].y / nrm + 1e-6f, v[i].z / nrm + 1e-6f, v[i].w / nrm + 1e-6f);
      }
      store4(orow + i * 128 + lane * 4, o);
    }
}

// ------------------------------------------------------------------ box head tail
// HF:1024 dense2 (D -> 4) + reference src/models.py:71-73: + box bias, sigmoid, cxcywh -> xyxy.
// h [M,D] fp16, w [4,D] fp32, bias [4], box_bias [P,4]; boxes [M,4] fp32, sig [M,4] fp32 (saved for bwd).
__global__ void box_tail_kernel(const __half* __restrict__ h, const float* __restrict__ w,
                                const float* __restrict__ bias, const float* __restrict__ box_bias,
                                float* __restrict__ boxes, float* __restrict__ sig, int M, int P, int D) {
  pdl_grid_wait();
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const __half* hr = h + 1LL * row * D;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(hr + c);
    const __half2* hp = reinterpret_cast<const __half2*>(&u);
    float x[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(hp[j]);
      x[2 * j] = f.x; x[2 * j + 1] = f.y;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + 1LL * o * D + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + 1LL * o * D + c + 4));
      acc[o] += x[0] * w0.x + x[1] * w0.y + x[2] * w0.z + x[3] * w0.w + x[4] * w1.x + x[5] * w1.y +
                x[6] * w1.z + x[7] * w1.w;
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = warp_sum(acc[o]);
  if (lane == 0) {
    const int p = row % P;
    float s[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float z = acc[o] + bias[o] + box_bias[p * 4 + o];
      s[o] = 1.0f / (1.0f + expf(-z));
    }
    *reinterpret_cast<float4*>(sig + 4LL * row) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(boxes + 4LL * row) =
        make_float4(s[0] - 0.5f * s[2], s[1] - 0.5f * s[3], s[0] + 0.5f * s[2], s[1] + 0.5f * s[3]);
  }
}

// ------------------------------------------------------------------ fp32 -> fp16 cast with scale
__global__ void cast_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n, float scale) {
  pdl_grid_wait();
  const long long i = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    store4(dst + i, make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale));
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2half_rn(src[j] * scale);
  }
}

}  // namespace owl

using namespace owl;

static inline int row_blocks(long long rows) { return static_cast<int>((rows + ROW_WARPS - 1) / ROW_WARPS); }

extern "C" int owl_im2col_f16(const float* img, void* out, int B, int image_size, int patch, long long ld,
                              void* stream) {
  OWL_CHECK_ARG(img && out && B > 0 && patch > 0 && image_size % patch == 0, "im2col: bad arguments");
  OWL_CHECK_ARG(image_size % 8 == 0 && patch % 2 == 0, "im2col: image_size %% 8 and patch %% 2 must be 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool v8 = patch % 8 == 0 && ld % 8 == 0;
  const long long total = 1LL * B * 3 * image_size * (image_size / (v8 ? 8 : 2));
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  if (v8) OWL_LAUNCH(im2col_kernel<8>, blocks, 256, 0, s, img, static_cast<__half*>(out), B, image_size, patch, (int)ld);
  else OWL_LAUNCH(im2col_kernel<2>, blocks, 256, 0, s, img, static_cast<__half*>(out), B, image_size, patch, (int)ld);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_u8_patches_f16(const unsigned char* img, const float* lut, void* out, int B, int image_size,
                                  int patch, long long ld, void* stream) {
  OWL_CHECK_ARG(img && lut && out && B > 0 && patch > 0 && image_size % patch == 0, "u8_patches: bad arguments");
  OWL_CHECK_ARG(patch % 8 == 0 && ld % 8 == 0, "u8_patches: patch %% 8 and ld %% 8 must be 0 (patch %d, ld %lld)", patch, ld);
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 7) == 0, "u8_patches: img must be 8-byte aligned");
  const long long total = 1LL * B * image_size * (image_size / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
  OWL_LAUNCH(u8_patches_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), img, lut, static_cast<__half*>(out), B,
             image_size, patch, static_cast<int>(ld));
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_layernorm(const float* x, long long x_stride, const float* gamma, const float* beta, void* y,
                             long long y_stride, int out_f16, int rows, int D, float eps, const float* cls_emb,
                             const float* pos0, int tokens, void* stream) {
  OWL_CHECK_ARG(x && gamma && beta && y && rows > 0, "layernorm: null / empty argument");
  OWL_CHECK_ARG(D % 128 == 0 && D <= 128 * MAX_VEC, "layernorm: D = %d must be a multiple of 128 and <= %d", D,
                128 * MAX_VEC);
  OWL_CHECK_ARG(!cls_emb || (pos0 && tokens > 0), "layernorm: cls_emb needs pos0 and tokens");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define OWL_LN_CASE(NV)                                                                                          \
  case NV:                                                                                                       \
    if (out_f16)                                                                                                 \
      OWL_LAUNCH((layernorm_kernel<__half, NV>), row_blocks(rows), ROW_WARPS * 32, 0, s, x, x_stride, gamma, beta, \
                 static_cast<__half*>(y), y_stride, rows, eps, cls_emb, pos0, tokens);                          \
    else                                                                                                         \
      OWL_LAUNCH((layernorm_kernel<float, NV>), row_blocks(rows), ROW_WARPS * 32, 0, s, x, x_stride, gamma, beta, \
                 static_cast<float*>(y), y_stride, rows, eps, cls_emb, pos0, tokens);                           \
    break;
  switch (D / 128) {
    OWL_LN_CASE(1) OWL_LN_CASE(2) OWL_LN_CASE(3) OWL_LN_CASE(4) OWL_LN_CASE(5) OWL_LN_CASE(6) OWL_LN_CASE(7) OWL_LN_CASE(8)
  }
#undef OWL_LN_CASE
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_post_fuse(const float* x, const float* ecls, const float* g1, const float* b1, const float* g2,
                             const float* b2, void* feats, int B, int P, int D, float eps, void* stream) {
  OWL_CHECK_ARG(x && ecls && g1 && b1 && g2 && b2 && feats && B > 0 && P > 0, "post_fuse: null / empty argument");
  OWL_CHECK_ARG(D % 128 == 0 && D <= 128 * MAX_VEC, "post_fuse: unsupported D = %d", D);
#define OWL_PF_CASE(NV)                                                                                         \
  case NV:                                                                                                      \
    OWL_LAUNCH(post_fuse_kernel<NV>, row_blocks(1LL * B * P), ROW_WARPS * 32, 0, static_cast<cudaStream_t>(stream), \
               x, ecls, g1, b1, g2, b2, static_cast<__half*>(feats), B, P, eps);                                \
    break;
  switch (D / 128) {
    OWL_PF_CASE(1) OWL_PF_CASE(2) OWL_PF_CASE(3) OWL_PF_CASE(4) OWL_PF_CASE(5) OWL_PF_CASE(6) OWL_PF_CASE(7) OWL_PF_CASE(8)
  }
#undef OWL_PF_CASE
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_rownorm_f16(const float* e, void* out, int rows, int E, int query_mode, void* stream) {
  OWL_CHECK_ARG(e && out && rows > 0, "rownorm: null / empty argument");
  OWL_CHECK_ARG(E % 128 == 0 && E <= 128 * MAX_VEC, "rownorm: unsupported E = %d", E);
  OWL_LAUNCH(rownorm_kernel, row_blocks(rows), ROW_WARPS * 32, 0, static_cast<cudaStream_t>(stream), 
      e, static_cast<__half*>(out), rows, E, query_mode);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_box_tail(const void* h, const float* w, const float* bias, const float* box_bias, float* boxes,
                            float* sig, int M, int P, int D, void* stream) {
  OWL_CHECK_ARG(h && w && bias && box_bias && boxes && sig && M > 0 && P > 0, "box_tail: null / empty argument");
  OWL_CHECK_ARG(D % 8 == 0, "box_tail: D %% 8 != 0");
  OWL_LAUNCH(box_tail_kernel, row_blocks(M), ROW_WARPS * 32, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(h), w, bias, box_bias, boxes, sig, M, P, D);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_cast_f16(const float* src, void* dst, long long n, float scale, void* stream) {
  OWL_CHECK_ARG(src && dst && n > 0, "cast_f16: null / empty argument");
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                "cast_f16: misaligned pointers");
  const long long threads = (n + 3) / 4;
  OWL_LAUNCH(cast_f16_kernel, static_cast<int>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), 
      src, static_cast<__half*>(dst), n, scale);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
