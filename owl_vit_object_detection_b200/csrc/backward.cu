// HBM-bound backward kernels around the dgrad / wgrad GEMMs (autograd of reference main.py:90 under the
// reference freeze rule, src/models.py:173-184): gradient scaling, max-pool / row-norm / box-tail / softmax /
// LayerNorm backward, bias gradients (column sums) and the fused AdamW step (reference main.py:56-60,91).
//
// Activation gradients travel between kernels in fp16 (they are tensor-core GEMM operands).  To keep them
// inside fp16's range they are multiplied by a power-of-two `S` chosen on the device from max|dL/d(sims,boxes)|
// (owl_grad_scale); every kernel that writes a PARAMETER gradient multiplies by 1/S, so the flat fp32 gradient
// buffer holds true gradients.  gscale[0] = S, gscale[1] = 1/S.
#include "common.h"
#include <algorithm>
#include <cuda_fp16.h>

namespace owl {

constexpr int BW_MAX_VEC = 8;  // row length <= 1024

__device__ __forceinline__ float bw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void bw_store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void bw_store4(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float4 f4_load_h(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------ gradient scale
__global__ void amax_kernel(const float* __restrict__ a, long long na, const float* __restrict__ b, long long nb,
                            unsigned int* __restrict__ bits) {
  pdl_grid_wait();
  float m = 0.f;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < na; i += 1LL * gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(a[i]));
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < nb; i += 1LL * gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(b[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // NaN / inf propagate as "very large": fmaxf drops NaN, so test explicitly
  if ((threadIdx.x & 31) == 0) atomicMax(bits, __float_as_uint(m));
}
__global__ void scale_from_amax_kernel(const unsigned int* __restrict__ bits, float target, float* __restrict__ gscale) {
  pdl_grid_wait();
  const float amax = __uint_as_float(*bits);
  float s = 1.0f;
  if (amax > 0.f && isfinite(amax)) {
    int e;
    frexpf(target / amax, &e);        // target / amax = f * 2^e, f in [0.5, 1)
    e = max(-24, min(24, e - 1));
    s = ldexpf(1.0f, e);
  }
  gscale[0] = s;
  gscale[1] = 1.0f / s;
}

// ------------------------------------------------------------------ class head: max-pool-3 backward
// dfull[m, 3c + j] = S * dsims[m, c] if j == argmax[m, c] else 0   (reference src/models.py:36)
__global__ void pool3_bwd_kernel(const float* __restrict__ dsims, const uint8_t* __restrict__ argmax,
                                 const float* __restrict__ gscale, __half* __restrict__ dfull, long long n) {
  pdl_grid_wait();
  const float S = gscale[0];
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) {
    const float g = dsims[i] * S;
    const int a = argmax[i];
    const __half z = __float2half(0.f), v = __float2half_rn(g);
    dfull[3 * i + 0] = a == 0 ? v : z;
    dfull[3 * i + 1] = a == 1 ? v : z;
    dfull[3 * i + 2] = a == 2 ? v : z;
  }
}

// ------------------------------------------------------------------ row-norm backward (reference src/models.py:28-33)
// mode 0: y = e / (|e| + eps)   -> de = dy / (n + eps) - e * (e . dy) / (n * (n + eps)^2)          (fp16 out)
// mode 1: y = q / |q| + eps     -> dq += (dy - qh * (qh . dy)) / |q|,  qh = q / |q|  (fp32, times gscale[1], ACCUMULATED)
template <typename OutT>
__global__ void rownorm_bwd_kernel(const float* __restrict__ e, const float* __restrict__ dy, OutT* __restrict__ out,
                                   int rows, int E, int mode, const float* __restrict__ gscale) {
  pdl_grid_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = E >> 7;
  float4 x[BW_MAX_VEC], g[BW_MAX_VEC];
  float q = 0.f, d = 0.f;
#pragma unroll
  for (int i = 0; i < BW_MAX_VEC; ++i)
    if (i < nv) {
      x[i] = *reinterpret_cast<const float4*>(e + 1LL * row * E + i * 128 + lane * 4);
      g[i] = *reinterpret_cast<const float4*>(dy + 1LL * row * E + i * 128 + lane * 4);
      q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
      d += (x[i].x * g[i].x + x[i].y * g[i].y) + (x[i].z * g[i].z + x[i].w * g[i].w);
    }
  const float n = sqrtf(bw_warp_sum(q));
  d = bw_warp_sum(d);
  float a, b;   // out = a * dy - b * x
  if (mode == 0) {
    const float ne = n + 1e-6f;
    a = 1.0f / ne;
    b = d / (n * ne * ne);
  } else {
    const float us = gscale[1];
    a = us / n;
    b = us * d / (n * n * n);
  }
#pragma unroll
  for (int i = 0; i < BW_MAX_VEC; ++i)
    if (i < nv) {
      float4 o;
      o.x = a * g[i].x - b * x[i].x; o.y = a * g[i].y - b * x[i].y;
      o.z = a * g[i].z - b * x[i].z; o.w = a * g[i].w - b * x[i].w;
      if constexpr (sizeof(OutT) == 4) {   // parameter gradient: accumulate like autograd does
        const float4 old = *reinterpret_cast<const float4*>(out + 1LL * row * E + i * 128 + lane * 4);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      bw_store4(out + 1LL * row * E + i * 128 + lane * 4, o);
    }
}

// ------------------------------------------------------------------ box head tail backward
// boxes = corners(sigmoid(z)), z = h1 W2^T + b2 + box_bias   (reference src/models.py:71-73, HF:1024)
//   dz[m,:]   = S * J^T dboxes[m,:] * s (1 - s)                      fp32 [M,4]
//   dpre1[m,:] = (dz[m,:] @ W2) * gelu'(pre1[m,:])                   fp16 [M,D]
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__global__ void box_tail_bwd_kernel(const float* __restrict__ dboxes, const float* __restrict__ sig,
                                    const float* __restrict__ w2, const __half* __restrict__ pre1,
                                    const float* __restrict__ gscale, float* __restrict__ dz,
                                    __half* __restrict__ dpre1, int M, int D) {
  pdl_grid_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float S = gscale[0];
  const float4 db = *reinterpret_cast<const float4*>(dboxes + 4LL * row);
  const float4 s = *reinterpret_cast<const float4*>(sig + 4LL * row);
  // x0 = cx - w/2, y0 = cy - h/2, x1 = cx + w/2, y1 = cy + h/2
  float4 z;
  z.x = S * (db.x + db.z) * s.x * (1.0f - s.x);
  z.y = S * (db.y + db.w) * s.y * (1.0f - s.y);
  z.z = S * 0.5f * (db.z - db.x) * s.z * (1.0f - s.z);
  z.w = S * 0.5f * (db.w - db.y) * s.w * (1.0f - s.w);
  if (lane == 0) *reinterpret_cast<float4*>(dz + 4LL * row) = z;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(w2 + c));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(w2 + D + c));
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(w2 + 2 * D + c));
    const float4 a3 = __ldg(reinterpret_cast<const float4*>(w2 + 3 * D + c));
    const float4 p = f4_load_h(pre1 + 1LL * row * D + c);
    float4 o;
    o.x = (z.x * a0.x + z.y * a1.x + z.z * a2.x + z.w * a3.x) * gelu_grad_f(p.x);
    o.y = (z.x * a0.y + z.y * a1.y + z.z * a2.y + z.w * a3.y) * gelu_grad_f(p.y);
    o.z = (z.x * a0.z + z.y * a1.z + z.z * a2.z + z.w * a3.z) * gelu_grad_f(p.z);
    o.w = (z.x * a0.w + z.y * a1.w + z.z * a2.w + z.w * a3.w) * gelu_grad_f(p.w);
    bw_store4(dpre1 + 1LL * row * D + c, o);
  }
}
// dW2[o, d] += (1/S) sum_m dz[m, o] h1[m, d];  db2[o] += (1/S) sum_m dz[m, o].   Block = 64 rows, thread = column.
__global__ void box_w2_grad_kernel(const float* __restrict__ dz, const __half* __restrict__ h1,
                                   const float* __restrict__ gscale, float* __restrict__ dw2,
                                   float* __restrict__ db2, int M, int D) {
  pdl_grid_wait();
  const int m0 = blockIdx.x * 64, m1 = min(M, m0 + 64);
  const float us = gscale[1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int m = m0; m < m1; ++m) {
      const float4 z = __ldg(reinterpret_cast<const float4*>(dz + 4LL * m));
      const float h = __half2float(h1[1LL * m * D + d]);
      a0 += z.x * h; a1 += z.y * h; a2 += z.z * h; a3 += z.w * h;
    }
    atomicAdd(dw2 + d, a0 * us);
    atomicAdd(dw2 + D + d, a1 * us);
    atomicAdd(dw2 + 2 * D + d, a2 * us);
    atomicAdd(dw2 + 3 * D + d, a3 * us);
  }
  if (threadIdx.x < 4) {
    float a = 0.f;
    for (int m = m0; m < m1; ++m) a += dz[4LL * m + threadIdx.x];
    atomicAdd(db2 + threadIdx.x, a * us);
  }
}

// ------------------------------------------------------------------ bias gradients: column sums
// out[n] += gscale[1] * sum_m x[m, n].   Block (32, 8) covers 64 columns x 256 rows.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long ld, int M, int N, const float* __restrict__ gscale,
                              float* __restrict__ out) {
  pdl_grid_wait();
  __shared__ float red[8][64];
  const int c = blockIdx.x * 64 + threadIdx.x * 2;
  const int m0 = blockIdx.y * 256;
  const int m1 = min(M, m0 + 256);
  float a = 0.f, b = 0.f;
  if (c < N) {
    // eight rows in flight per thread (the loads of a batch are issued before the first add)
    for (int m = m0 + threadIdx.y; m < m1; m += 64) {
      float2 f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int mm = min(m + 8 * k, M - 1);
        if constexpr (sizeof(T) == 2) f[k] = __half22float2(*reinterpret_cast<const __half2*>(x + mm * ld + c));
        else f[k] = *reinterpret_cast<const float2*>(x + mm * ld + c);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (m + 8 * k < m1) { a += f[k].x; b += f[k].y; }
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = a;
  red[threadIdx.y][threadIdx.x * 2 + 1] = b;
  __syncthreads();
  if (threadIdx.y == 0) {
    const float us = gscale ? gscale[1] : 1.0f;
    for (int j = 0; j < 2; ++j) {
      const int cc = threadIdx.x * 2 + j;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][cc];
      if (blockIdx.x * 64 + cc < N) atomicAdd(out + blockIdx.x * 64 + cc, s * us);
    }
  }
}

// Vectorised flavour (N, ld multiples of the 16-byte vector, aligned base): a thread owns VEC = 16 / sizeof(T)
// consecutive columns and reads 16 bytes per row with four rows in flight; a warp covers 32 * VEC columns of a row
// (512 bytes), the eight warps of a CTA take interleaved rows of the CTA's row range.  Optionally writes the fp16 copy
// of an fp32 input in the same pass (the cast the next GEMM's operand needs: one read of the gradient instead of two).
constexpr int CS_WARPS = 8;
template <typename T, bool CAST>
__global__ void __launch_bounds__(CS_WARPS * 32)
colsum_vec_kernel(const T* __restrict__ x, long long ld, int M, int N, int rows_per_cta,
                  const float* __restrict__ gscale, float* __restrict__ out, __half* __restrict__ y16, long long ldy) {
  constexpr int VEC = 16 / sizeof(T);
  pdl_grid_wait();
  __shared__ float red[CS_WARPS][32 * VEC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (blockIdx.x * 32 + lane) * VEC;
  const int m0 = blockIdx.y * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  if (c < N) {
    for (int m = m0 + warp; m < m1; m += 4 * CS_WARPS) {
      uint4 v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int mm = min(m + r * CS_WARPS, M - 1);
        v[r] = *reinterpret_cast<const uint4*>(x + mm * ld + c);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (m + r * CS_WARPS < m1) {
          if constexpr (sizeof(T) == 2) {
            const __half2* h = reinterpret_cast<const __half2*>(&v[r]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __half22float2(h[k]);
              acc[2 * k] += f.x;
              acc[2 * k + 1] += f.y;
            }
          } else {
            const float* f = reinterpret_cast<const float*>(&v[r]);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += f[k];
            if constexpr (CAST) {
              const __half2 lo = __floats2half2_rn(f[0], f[1]), hi = __floats2half2_rn(f[2], f[3]);
              uint2 o;
              o.x = *reinterpret_cast<const uint32_t*>(&lo);
              o.y = *reinterpret_cast<const uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(y16 + (m + r * CS_WARPS) * ldy + c) = o;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) red[warp][lane * VEC + k] = acc[k];
  __syncthreads();
  const float us = gscale ? gscale[1] : 1.0f;
  for (int j = threadIdx.x; j < 32 * VEC; j += CS_WARPS * 32) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < CS_WARPS; ++w) sum += red[w][j];
    const int cc = blockIdx.x * 32 * VEC + j;
    if (cc < N) atomicAdd(out + cc, sum * us);
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// y = xh * gamma + beta, xh = (x - mean) * rstd.
//   dx = rstd * (g - mean(g) - xh * mean(g * xh)),  g = dy * gamma          (+ dx_add when given)
//   dgamma += us * sum_rows dy * xh;  dbeta += us * sum_rows dy
// CTA = 4 warps, each warp walks rows r0 + w, r0 + w + 4, ... of its chunk and keeps private parameter-gradient
// accumulators in shared memory; one set of global atomics per CTA.
constexpr int LNB_WARPS = 4;

template <int NV>
__global__ void __launch_bounds__(LNB_WARPS * 32)
ln_bwd_kernel(const float* __restrict__ x, long long x_stride, const float* __restrict__ dy, long long dy_stride,
              const float* __restrict__ gamma, const float* __restrict__ dx_add, float* __restrict__ dx,
              long long dx_stride, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, float eps,
              int rows_per_cta, const float* __restrict__ gscale) {
  constexpr int D = NV * 128;
  pdl_grid_wait();
  extern __shared__ float lnb_sm[];  // [LNB_WARPS][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* accg = lnb_sm + (warp * 2) * D;
  float* accb = accg + D;
  for (int i = lane; i < 2 * D; i += 32) accg[i] = 0.f;
  float4 gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gm[i] = __ldg(reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4));
  __syncwarp();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  for (int row = r0 + warp; row < r1; row += LNB_WARPS) {
    // all of the row's loads are issued before the first use (in-order issue would expose a latency per load)
    float4 xv[NV], gv[NV], av[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const float4*>(x + row * x_stride + i * 128 + lane * 4);
#pragma unroll
    for (int i = 0; i < NV; ++i) gv[i] = *reinterpret_cast<const float4*>(dy + row * dy_stride + i * 128 + lane * 4);
    if (dx_add) {
#pragma unroll
      for (int i = 0; i < NV; ++i) av[i] = *reinterpret_cast<const float4*>(dx_add + row * dx_stride + i * 128 + lane * 4);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    const float mean = bw_warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
    const float rstd = rsqrtf(bw_warp_sum(q) / D + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;   // xh
      float4 a = *reinterpret_cast<float4*>(accg + c), b = *reinterpret_cast<float4*>(accb + c);
      a.x += gv[i].x * xv[i].x; a.y += gv[i].y * xv[i].y; a.z += gv[i].z * xv[i].z; a.w += gv[i].w * xv[i].w;
      b.x += gv[i].x; b.y += gv[i].y; b.z += gv[i].z; b.w += gv[i].w;
      *reinterpret_cast<float4*>(accg + c) = a;
      *reinterpret_cast<float4*>(accb + c) = b;
      gv[i].x *= gm[i].x; gv[i].y *= gm[i].y; gv[i].z *= gm[i].z; gv[i].w *= gm[i].w;   // g = dy * gamma
      sg += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      sgx += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
    }
    if (dx) {
      const float mg = bw_warp_sum(sg) / D, mgx = bw_warp_sum(sgx) / D;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = i * 128 + lane * 4;
        float4 o;
        o.x = rstd * (gv[i].x - mg - xv[i].x * mgx); o.y = rstd * (gv[i].y - mg - xv[i].y * mgx);
        o.z = rstd * (gv[i].z - mg - xv[i].z * mgx); o.w = rstd * (gv[i].w - mg - xv[i].w * mgx);
        if (dx_add) { o.x += av[i].x; o.y += av[i].y; o.z += av[i].z; o.w += av[i].w; }
        *reinterpret_cast<float4*>(dx + row * dx_stride + c) = o;
      }
    }
  }
  __syncthreads();
  const float us = gscale ? gscale[1] : 1.0f;
  for (int i = threadIdx.x; i < D; i += LNB_WARPS * 32) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) { a += lnb_sm[(w * 2) * D + i]; b += lnb_sm[(w * 2 + 1) * D + i]; }
    atomicAdd(dgamma + i, a * us);
    atomicAdd(dbeta + i, b * us);
  }
}

// ------------------------------------------------------------------ backward of reference src/models.py:80-86
// feats = LN2(LN1(x_p) * cl),  cl = LN1(x_cls).  Given dfeats (fp32, scaled): dx for the patch rows,
// dcl[b] (atomic, scaled), and the four LayerNorm parameter gradients (atomic, un-scaled).
template <int NV>
__global__ void __launch_bounds__(LNB_WARPS * 32)
post_fuse_bwd_kernel(const float* __restrict__ x, const float* __restrict__ ecls, const float* __restrict__ g1,
                     const float* __restrict__ b1, const float* __restrict__ g2, const float* __restrict__ dfeats,
                     float* __restrict__ dx, float* __restrict__ dcl, float* __restrict__ dg1, float* __restrict__ db1,
                     float* __restrict__ dg2, float* __restrict__ db2, int P, float eps, int rows_per_cta,
                     const float* __restrict__ gscale) {
  constexpr int D = NV * 128;
  pdl_grid_wait();
  extern __shared__ float pf_sm[];  // [LNB_WARPS][5][D]: dg2, db2, dcl, dg1, db1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  float* acc = pf_sm + warp * 5 * D;
  for (int i = lane; i < 5 * D; i += 32) acc[i] = 0.f;
  __syncwarp();
  const float* cr = ecls + 1LL * b * D;
  const int p0 = blockIdx.x * rows_per_cta, p1 = min(P, p0 + rows_per_cta);
  for (int p = p0 + warp; p < p1; p += LNB_WARPS) {
    const float* xr = x + (1LL * b * (P + 1) + 1 + p) * D;
    const float* dyr = dfeats + (1LL * b * P + p) * D;
    // the row's loads are issued before the first use (in-order issue would expose a latency per load)
    float4 xh[NV], uh[NV], g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) xh[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
#pragma unroll
    for (int i = 0; i < NV; ++i) g[i] = *reinterpret_cast<const float4*>(dyr + i * 128 + lane * 4);   // dy for now
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (xh[i].x + xh[i].y) + (xh[i].z + xh[i].w);
    float mean = bw_warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xh[i].x -= mean; xh[i].y -= mean; xh[i].z -= mean; xh[i].w -= mean;
      q += (xh[i].x * xh[i].x + xh[i].y * xh[i].y) + (xh[i].z * xh[i].z + xh[i].w * xh[i].w);
    }
    const float rstd1 = rsqrtf(bw_warp_sum(q) / D + eps);
    s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      xh[i].x *= rstd1; xh[i].y *= rstd1; xh[i].z *= rstd1; xh[i].w *= rstd1;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(g1 + c));
      const float4 be = __ldg(reinterpret_cast<const float4*>(b1 + c));
      const float4 cl = __ldg(reinterpret_cast<const float4*>(cr + c));
      uh[i].x = (xh[i].x * gm.x + be.x) * cl.x; uh[i].y = (xh[i].y * gm.y + be.y) * cl.y;
      uh[i].z = (xh[i].z * gm.z + be.z) * cl.z; uh[i].w = (xh[i].w * gm.w + be.w) * cl.w;
      s += (uh[i].x + uh[i].y) + (uh[i].z + uh[i].w);
    }
    mean = bw_warp_sum(s) / D;
    q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uh[i].x -= mean; uh[i].y -= mean; uh[i].z -= mean; uh[i].w -= mean;
      q += (uh[i].x * uh[i].x + uh[i].y * uh[i].y) + (uh[i].z * uh[i].z + uh[i].w * uh[i].w);
    }
    const float rstd2 = rsqrtf(bw_warp_sum(q) / D + eps);
    float sg = 0.f, sgu = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      uh[i].x *= rstd2; uh[i].y *= rstd2; uh[i].z *= rstd2; uh[i].w *= rstd2;
      const float4 dy = g[i];
      float4 a = *reinterpret_cast<float4*>(acc + c), bb = *reinterpret_cast<float4*>(acc + D + c);
      a.x += dy.x * uh[i].x; a.y += dy.y * uh[i].y; a.z += dy.z * uh[i].z; a.w += dy.w * uh[i].w;
      bb.x += dy.x; bb.y += dy.y; bb.z += dy.z; bb.w += dy.w;
      *reinterpret_cast<float4*>(acc + c) = a;
      *reinterpret_cast<float4*>(acc + D + c) = bb;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(g2 + c));
      g[i].x = dy.x * gm.x; g[i].y = dy.y * gm.y; g[i].z = dy.z * gm.z; g[i].w = dy.w * gm.w;
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgu += (g[i].x * uh[i].x + g[i].y * uh[i].y) + (g[i].z * uh[i].z + g[i].w * uh[i].w);
    }
    float mg = bw_warp_sum(sg) / D, mgx = bw_warp_sum(sgu) / D;
    sg = 0.f; sgu = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      // du
      float4 du;
      du.x = rstd2 * (g[i].x - mg - uh[i].x * mgx); du.y = rstd2 * (g[i].y - mg - uh[i].y * mgx);
      du.z = rstd2 * (g[i].z - mg - uh[i].z * mgx); du.w = rstd2 * (g[i].w - mg - uh[i].w * mgx);
      const float4 gm = __ldg(reinterpret_cast<const float4*>(g1 + c));
      const float4 be = __ldg(reinterpret_cast<const float4*>(b1 + c));
      const float4 cl = __ldg(reinterpret_cast<const float4*>(cr + c));
      float4 a = *reinterpret_cast<float4*>(acc + 2 * D + c);          // dcl += du * t
      a.x += du.x * (xh[i].x * gm.x + be.x); a.y += du.y * (xh[i].y * gm.y + be.y);
      a.z += du.z * (xh[i].z * gm.z + be.z); a.w += du.w * (xh[i].w * gm.w + be.w);
      *reinterpret_cast<float4*>(acc + 2 * D + c) = a;
      float4 dt = make_float4(du.x * cl.x, du.y * cl.y, du.z * cl.z, du.w * cl.w);
      float4 ag = *reinterpret_cast<float4*>(acc + 3 * D + c), ab = *reinterpret_cast<float4*>(acc + 4 * D + c);
      ag.x += dt.x * xh[i].x; ag.y += dt.y * xh[i].y; ag.z += dt.z * xh[i].z; ag.w += dt.w * xh[i].w;
      ab.x += dt.x; ab.y += dt.y; ab.z += dt.z; ab.w += dt.w;
      *reinterpret_cast<float4*>(acc + 3 * D + c) = ag;
      *reinterpret_cast<float4*>(acc + 4 * D + c) = ab;
      g[i].x = dt.x * gm.x; g[i].y = dt.y * gm.y; g[i].z = dt.z * gm.z; g[i].w = dt.w * gm.w;   // dxh
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgu += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    mg = bw_warp_sum(sg) / D; mgx = bw_warp_sum(sgu) / D;
    float* dxr = dx + (1LL * b * (P + 1) + 1 + p) * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd1 * (g[i].x - mg - xh[i].x * mgx); o.y = rstd1 * (g[i].y - mg - xh[i].y * mgx);
      o.z = rstd1 * (g[i].z - mg - xh[i].z * mgx); o.w = rstd1 * (g[i].w - mg - xh[i].w * mgx);
      *reinterpret_cast<float4*>(dxr + i * 128 + lane * 4) = o;
    }
  }
  __syncthreads();
  const float us = gscale[1];
  for (int i = threadIdx.x; i < D; i += LNB_WARPS * 32) {
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w)
#pragma unroll
      for (int k = 0; k < 5; ++k) a[k] += pf_sm[(w * 5 + k) * D + i];
    atomicAdd(dg2 + i, a[0] * us);
    atomicAdd(db2 + i, a[1] * us);
    atomicAdd(dcl + 1LL * b * D + i, a[2]);
    atomicAdd(dg1 + i, a[3] * us);
    atomicAdd(db1 + i, a[4] * us);
  }
}

// ------------------------------------------------------------------ fused AdamW (torch.optim.AdamW semantics)
// p *= 1 - lr * wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
// also refreshes the fp16 shadow used by the GEMMs.  grad_mul folds the 1/world of the all-reduce average.
// state[0] = step count (as float), state[1] = 1 - beta1^step, state[2] = sqrt(1 - beta2^step): advanced on the
// device so that a captured CUDA graph of the step replays with the right bias correction.
__global__ void adamw_advance_kernel(float* __restrict__ state, float beta1, float beta2) {
  pdl_grid_wait();
  const float step = state[0] + 1.0f;
  state[0] = step;
  state[1] = 1.0f - powf(beta1, step);
  state[2] = sqrtf(1.0f - powf(beta2, step));
}
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, __half* __restrict__ p16, long long n, float lr, float beta1,
                             float beta2, float eps, float wd, const float* __restrict__ state, float grad_mul) {
  pdl_grid_wait();
  const long long i = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float bc1 = state[1], bc2_sqrt = state[2];
  float4 pp = *reinterpret_cast<float4*>(p + i);
  const float4 gg = *reinterpret_cast<const float4*>(g + i);
  float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
  float* pa = reinterpret_cast<float*>(&pp);
  const float* ga = reinterpret_cast<const float*>(&gg);
  float* ma = reinterpret_cast<float*>(&mm);
  float* va = reinterpret_cast<float*>(&vv);
  const float step = lr / bc1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gk = ga[k] * grad_mul;
    pa[k] *= 1.0f - lr * wd;
    ma[k] = beta1 * ma[k] + (1.0f - beta1) * gk;
    va[k] = beta2 * va[k] + (1.0f - beta2) * gk * gk;
    pa[k] -= step * ma[k] / (sqrtf(va[k]) / bc2_sqrt + eps);
  }
  *reinterpret_cast<float4*>(p + i) = pp;
  *reinterpret_cast<float4*>(m + i) = mm;
  *reinterpret_cast<float4*>(v + i) = vv;
  if (p16) bw_store4(p16 + i, pp);
}

}  // namespace owl

using namespace owl;

extern "C" int owl_grad_scale(const float* a, long long na, const float* b, long long nb, float target, float* gscale,
                              void* stream) {
  OWL_CHECK_ARG(a && gscale && na > 0 && target > 0.f, "grad_scale: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned int* bits = reinterpret_cast<unsigned int*>(gscale + 2);
  OWL_CUDA(cudaMemsetAsync(bits, 0, sizeof(unsigned int), s));
  const int blocks = static_cast<int>(std::min<long long>((std::max(na, nb) + 255) / 256, 148LL * 8));
  OWL_LAUNCH(amax_kernel, blocks, 256, 0, s, a, na, b, b ? nb : 0, bits);
  OWL_LAUNCH(scale_from_amax_kernel, 1, 1, 0, s, bits, target, gscale);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_pool3_bwd(const float* dsims, const uint8_t* argmax, const float* gscale, void* dfull_f16,
                             long long n, void* stream) {
  OWL_CHECK_ARG(dsims && argmax && gscale && dfull_f16 && n > 0, "pool3_bwd: bad arguments");
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148LL * 16));
  OWL_LAUNCH(pool3_bwd_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), dsims, argmax, gscale,
                                                                          static_cast<__half*>(dfull_f16), n);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_rownorm_bwd(const float* e, const float* dy, void* out, int rows, int E, int query_mode,
                               const float* gscale, void* stream) {
  OWL_CHECK_ARG(e && dy && out && rows > 0 && gscale, "rownorm_bwd: bad arguments");
  OWL_CHECK_ARG(E % 128 == 0 && E <= 128 * BW_MAX_VEC, "rownorm_bwd: unsupported E = %d", E);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (query_mode)
    OWL_LAUNCH(rownorm_bwd_kernel<float>, (rows + 7) / 8, 256, 0, s, e, dy, static_cast<float*>(out), rows, E, 1, gscale);
  else
    OWL_LAUNCH(rownorm_bwd_kernel<__half>, (rows + 7) / 8, 256, 0, s, e, dy, static_cast<__half*>(out), rows, E, 0, gscale);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_box_tail_bwd(const float* dboxes, const float* sig, const float* w2, const void* pre1_f16,
                                const void* h1_f16, const float* gscale, float* dz, void* dpre1_f16, float* dw2,
                                float* db2, int M, int D, void* stream) {
  OWL_CHECK_ARG(dboxes && sig && w2 && pre1_f16 && h1_f16 && gscale && dz && dpre1_f16 && dw2 && db2 && M > 0,
                "box_tail_bwd: bad arguments");
  OWL_CHECK_ARG(D % 4 == 0, "box_tail_bwd: D %% 4 != 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OWL_LAUNCH(box_tail_bwd_kernel, (M + 7) / 8, 256, 0, s, dboxes, sig, w2, static_cast<const __half*>(pre1_f16), gscale, dz,
                                                  static_cast<__half*>(dpre1_f16), M, D);
  OWL_LAUNCH(box_w2_grad_kernel, (M + 63) / 64, 256, 0, s, dz, static_cast<const __half*>(h1_f16), gscale, dw2, db2, M, D);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_colsum(const void* x, int is_f16, long long ld, int M, int N, const float* gscale, float* out,
                          void* cast_out_f16, void* stream) {
  OWL_CHECK_ARG(x && out && M > 0 && N > 0 && N % 2 == 0 && ld % 2 == 0, "colsum: bad arguments (N, ld even)");
  OWL_CHECK_ARG(!cast_out_f16 || !is_f16, "colsum: the fused fp16 copy is for fp32 input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vec = is_f16 ? 8 : 4;
  const bool aligned = N % vec == 0 && ld % vec == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                       (!cast_out_f16 || (reinterpret_cast<uintptr_t>(cast_out_f16) & 7) == 0);
  if (aligned) {
    // rows per CTA: about four CTAs per SM in total, a multiple of the 32 rows one pass of the CTA covers
    const int col_blocks = (N + 32 * vec - 1) / (32 * vec);
    int row_blocks = (4 * num_sms() + col_blocks - 1) / col_blocks;
    int rows_per_cta = (M + row_blocks - 1) / row_blocks;
    rows_per_cta = ((rows_per_cta + 4 * CS_WARPS - 1) / (4 * CS_WARPS)) * (4 * CS_WARPS);
    // up to 256 rows: ONE row block, i.e. one atomic per column and a run-to-run deterministic sum (a bias gradient
    // that is mathematically zero - k_proj.bias - is pure rounding noise whose sign decides a +-lr Adam step)
    if (M <= 256) rows_per_cta = 256;
    row_blocks = (M + rows_per_cta - 1) / rows_per_cta;
    dim3 grid(col_blocks, row_blocks);
    if (is_f16)
      OWL_LAUNCH((colsum_vec_kernel<__half, false>), grid, CS_WARPS * 32, 0, s, static_cast<const __half*>(x), ld, M, N,
                 rows_per_cta, gscale, out, static_cast<__half*>(nullptr), 0LL);
    else if (cast_out_f16)
      OWL_LAUNCH((colsum_vec_kernel<float, true>), grid, CS_WARPS * 32, 0, s, static_cast<const float*>(x), ld, M, N,
                 rows_per_cta, gscale, out, static_cast<__half*>(cast_out_f16), static_cast<long long>(N));
    else
      OWL_LAUNCH((colsum_vec_kernel<float, false>), grid, CS_WARPS * 32, 0, s, static_cast<const float*>(x), ld, M, N,
                 rows_per_cta, gscale, out, static_cast<__half*>(nullptr), 0LL);
    OWL_CUDA(cudaGetLastError());
    return OWL_OK;
  }
  OWL_CHECK_ARG(!cast_out_f16, "colsum: the fused fp16 copy needs N and ld to be multiples of 4 and a 16-byte aligned input");
  dim3 grid((N + 63) / 64, (M + 255) / 256), block(32, 8);
  if (is_f16) OWL_LAUNCH(colsum_kernel<__half>, grid, block, 0, s, static_cast<const __half*>(x), ld, M, N, gscale, out);
  else OWL_LAUNCH(colsum_kernel<float>, grid, block, 0, s, static_cast<const float*>(x), ld, M, N, gscale, out);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_layernorm_bwd(const float* x, long long x_stride, const float* dy, long long dy_stride,
                                 const float* gamma, const float* dx_add, float* dx, long long dx_stride,
                                 float* dgamma, float* dbeta, int rows, int D, float eps, const float* gscale,
                                 void* stream) {
  OWL_CHECK_ARG(x && dy && gamma && dgamma && dbeta && rows > 0, "layernorm_bwd: bad arguments");
  OWL_CHECK_ARG(D % 128 == 0 && D <= 128 * BW_MAX_VEC, "layernorm_bwd: unsupported D = %d", D);
  OWL_CHECK_ARG(!dx_add || dx, "layernorm_bwd: dx_add needs dx");
  const int rows_per_cta = rows >= 148 * 16 ? 16 : (rows >= 148 * 4 ? 8 : 4);
  const size_t smem = sizeof(float) * LNB_WARPS * 2 * D;
#define OWL_LNB_CASE(NV)                                                                                              \
  case NV:                                                                                                            \
    OWL_LAUNCH(ln_bwd_kernel<NV>, (rows + rows_per_cta - 1) / rows_per_cta, LNB_WARPS * 32, smem,                     \
               static_cast<cudaStream_t>(stream), x, x_stride, dy, dy_stride, gamma, dx_add, dx, dx_stride, dgamma,  \
               dbeta, rows, eps, rows_per_cta, gscale);                                                               \
    break;
  switch (D / 128) {
    OWL_LNB_CASE(1) OWL_LNB_CASE(2) OWL_LNB_CASE(3) OWL_LNB_CASE(4) OWL_LNB_CASE(5) OWL_LNB_CASE(6) OWL_LNB_CASE(7)
    OWL_LNB_CASE(8)
  }
#undef OWL_LNB_CASE
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_post_fuse_bwd(const float* x, const float* ecls, const float* g1, const float* b1, const float* g2,
                                 const float* dfeats, float* dx, float* dcl, float* dg1, float* db1, float* dg2,
                                 float* db2, int B, int P, int D, float eps, const float* gscale, void* stream) {
  OWL_CHECK_ARG(x && ecls && g1 && b1 && g2 && dfeats && dx && dcl && dg1 && db1 && dg2 && db2 && gscale && B > 0 && P > 0,
                "post_fuse_bwd: bad arguments");
  OWL_CHECK_ARG(D % 128 == 0 && D <= 128 * BW_MAX_VEC, "post_fuse_bwd: unsupported D = %d", D);
  const int rows_per_cta = 24;
  const size_t smem = sizeof(float) * LNB_WARPS * 5 * D;
  dim3 grid((P + rows_per_cta - 1) / rows_per_cta, B);
#define OWL_PFB_CASE(NV)                                                                                             \
  case NV: {                                                                                                         \
    static SmemOptIn optin;                                                                                          \
    OWL_CUDA(ensure_smem(optin, post_fuse_bwd_kernel<NV>, smem));                                                    \
    OWL_LAUNCH(post_fuse_bwd_kernel<NV>, grid, LNB_WARPS * 32, smem, static_cast<cudaStream_t>(stream), x, ecls, g1, \
               b1, g2, dfeats, dx, dcl, dg1, db1, dg2, db2, P, eps, rows_per_cta, gscale);                          \
  } break;
  switch (D / 128) {
    OWL_PFB_CASE(1) OWL_PFB_CASE(2) OWL_PFB_CASE(3) OWL_PFB_CASE(4) OWL_PFB_CASE(5) OWL_PFB_CASE(6) OWL_PFB_CASE(7)
    OWL_PFB_CASE(8)
  }
#undef OWL_PFB_CASE
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_f16,
                         long long n, float lr, float beta1, float beta2, float eps, float weight_decay, float* state,
                         float grad_mul, void* stream) {
  OWL_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state && n > 0 && n % 4 == 0, "adamw: bad arguments");
  const long long threads = n / 4;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OWL_LAUNCH(adamw_advance_kernel, 1, 1, 0, s, state, beta1, beta2);
  OWL_LAUNCH(adamw_kernel, static_cast<unsigned>((threads + 255) / 256), 256, 0, s, 
      params, grads, exp_avg, exp_avg_sq, static_cast<__half*>(params_f16), n, lr, beta1, beta2, eps, weight_decay,
      state, grad_mul);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
