// Query-bank initialisation on the device: the pieces of the OWL-ViT TEXT tower that are not already
// covered by the GEMM / LayerNorm kernels of the vision path.  reference src/models.py:155-169 runs the HuggingFace
// text model once over three prompts per class (240 prompts x 16 tokens for COCO) and keeps `text_embeds` as the
// learned query bank; here that is
//
//   owl_text_embed      HF:370-373   token + position embedding                        -> x [N*S, D] fp32
//   (owl_layernorm, owl_gemm: LN1, fused q|k|v projection, out-proj + residual, LN2, MLP with quick_gelu; the same
//    kernels as the vision encoder, HF:490-511)
//   owl_text_attn       HF:379-404 with the causal + padding mask of HF:661-666        -> ctx [N*S, D] fp16
//   owl_text_pool_ln    HF:677-684   final LayerNorm of the end-of-text row (argmax id)  -> pooled [N, D] fp16
//   (owl_gemm: text_projection, HF:978)
//   owl_l2norm_rows     HF:984       text_embeds / ||text_embeds||                      -> [N, E] fp32
//
// It runs once per training run on 3840 token rows: these kernels are written for clarity, not for a roofline.
#include "common.h"
#include <cuda_fp16.h>
#include <math_constants.h>

namespace owl {

constexpr int TEXT_MAX_S = 32;        // tokens per prompt the attention kernel supports (CLIP text: 16)
constexpr int TEXT_DH = 64;           // head dim
constexpr int TEXT_ATTN_WARPS = 4;    // (prompt, head) pairs per CTA

// x[row, :] = tok_emb[ids[row], :] + pos_emb[row % S, :]
__global__ void text_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok_emb,
                                  const float* __restrict__ pos_emb, float* __restrict__ x, int rows, int S, int D,
                                  int vocab, int* __restrict__ status) {
  pdl_grid_wait();
  const int row = blockIdx.x;
  if (row >= rows) return;
  long long id = ids[row];
  if (id < 0 || id >= vocab) {           // nn.Embedding raises IndexError: flag it, stay in bounds
    if (threadIdx.x == 0 && status) atomicOr(status, 8);
    id = 0;
  }
  const float4* t = reinterpret_cast<const float4*>(tok_emb + id * D);
  const float4* p = reinterpret_cast<const float4*>(pos_emb + static_cast<long long>(row % S) * D);
  float4* o = reinterpret_cast<float4*>(x + static_cast<long long>(row) * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 a = t[i], b = p[i];
    o[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// One warp per (prompt n, head h); lane i owns query token i (S <= 32).  Scores of a 16-token prompt are a 16 x 16
// triangle: K and V rows sit in shared memory (every lane reads the same row: broadcast), the query row, the scores
// and the output row of a lane stay in registers.  Key j is visible to query i iff j <= i and mask[n, j] != 0.
__global__ void __launch_bounds__(TEXT_ATTN_WARPS * 32)
text_attn_kernel(const __half* __restrict__ qkv, const int* __restrict__ mask, __half* __restrict__ ctx, int N, int S,
                 int H, float scale) {
  pdl_grid_wait();
  __shared__ __align__(16) __half sK[TEXT_ATTN_WARPS][TEXT_MAX_S][TEXT_DH];
  __shared__ __align__(16) __half sV[TEXT_ATTN_WARPS][TEXT_MAX_S][TEXT_DH];
  __shared__ int sM[TEXT_ATTN_WARPS][TEXT_MAX_S];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * TEXT_ATTN_WARPS + warp;
  if (pair >= N * H) return;               // whole warp leaves together; only __syncwarp below
  const int n = pair / H, h = pair - n * H;
  const int D = H * TEXT_DH;
  const long long row0 = static_cast<long long>(n) * S;
  // stage K and V of this (prompt, head): S rows x 64 halves = 8 x 16 bytes per row
  for (int c = lane; c < S * 8; c += 32) {
    const int j = c >> 3, part = c & 7;
    const __half* src = qkv + (row0 + j) * 3 * D + h * TEXT_DH + part * 8;
    *reinterpret_cast<uint4*>(&sK[warp][j][part * 8]) = *reinterpret_cast<const uint4*>(src + D);
    *reinterpret_cast<uint4*>(&sV[warp][j][part * 8]) = *reinterpret_cast<const uint4*>(src + 2 * D);
  }
  if (lane < S) sM[warp][lane] = mask ? mask[row0 + lane] : 1;
  __syncwarp();
  if (lane >= S) return;
  const int i = lane;
  __half2 q[TEXT_DH / 2];
  {
    const uint4* src = reinterpret_cast<const uint4*>(qkv + (row0 + i) * 3 * D + h * TEXT_DH);
#pragma unroll
    for (int part = 0; part < 8; ++part) {
      const uint4 v = src[part];
      q[part * 4 + 0] = *reinterpret_cast<const __half2*>(&v.x);
      q[part * 4 + 1] = *reinterpret_cast<const __half2*>(&v.y);
      q[part * 4 + 2] = *reinterpret_cast<const __half2*>(&v.z);
      q[part * 4 + 3] = *reinterpret_cast<const __half2*>(&v.w);
    }
  }
  float s[TEXT_MAX_S];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < TEXT_MAX_S; ++j) {
    s[j] = -CUDART_INF_F;
    if (j < S && j <= i && sM[warp][j] != 0) {
      float acc = 0.f;
      const __half2* k2 = reinterpret_cast<const __half2*>(&sK[warp][j][0]);
#pragma unroll
      for (int d = 0; d < TEXT_DH / 2; ++d) {
        const float2 a = __half22float2(q[d]), b = __half22float2(k2[d]);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
      }
      s[j] = acc * scale;
      m = fmaxf(m, s[j]);
    }
  }
  float o[TEXT_DH];
#pragma unroll
  for (int d = 0; d < TEXT_DH; ++d) o[d] = 0.f;
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < TEXT_MAX_S; ++j) {
    if (j < S && s[j] > -CUDART_INF_F) {
      const float p = __expf(s[j] - m);
      l += p;
      const __half2* v2 = reinterpret_cast<const __half2*>(&sV[warp][j][0]);
#pragma unroll
      for (int d = 0; d < TEXT_DH / 2; ++d) {
        const float2 b = __half22float2(v2[d]);
        o[2 * d] = fmaf(p, b.x, o[2 * d]);
        o[2 * d + 1] = fmaf(p, b.y, o[2 * d + 1]);
      }
    }
  }
  const float inv = l > 0.f ? 1.0f / l : 0.f;      // a query without a visible key (cannot happen for real prompts: 0
  __half* dst = ctx + (row0 + i) * D + h * TEXT_DH;  // rows of padding only) yields a zero row instead of NaN
#pragma unroll
  for (int part = 0; part < 8; ++part) {
    uint4 v;
    __half2 t;
    t = __floats2half2_rn(o[part * 8 + 0] * inv, o[part * 8 + 1] * inv); v.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(o[part * 8 + 2] * inv, o[part * 8 + 3] * inv); v.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(o[part * 8 + 4] * inv, o[part * 8 + 5] * inv); v.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2half2_rn(o[part * 8 + 6] * inv, o[part * 8 + 7] * inv); v.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(dst + part * 8) = v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per prompt: the end-of-text token is the FIRST position holding the largest id (torch.argmax, HF:683); its
// row of the residual stream goes through the final LayerNorm (two-pass variance like torch) into fp16.
__global__ void text_pool_ln_kernel(const float* __restrict__ x, const long long* __restrict__ ids,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    __half* __restrict__ out, int N, int S, int D, float eps) {
  pdl_grid_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= N) return;
  long long best = -0x7fffffffffffffffLL - 1;
  int pos = 0;
  for (int j = 0; j < S; ++j) {             // every lane scans the (<= 32) ids: no exchange needed
    const long long v = ids[static_cast<long long>(n) * S + j];
    if (v > best) { best = v; pos = j; }
  }
  const float* row = x + (static_cast<long long>(n) * S + pos) * D;
  float sum = 0.f;
  for (int d = lane; d < D; d += 32) sum += row[d];
  const float mean = warp_sum(sum) / D;
  float sq = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float c = row[d] - mean;
    sq = fmaf(c, c, sq);
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + eps);
  for (int d = lane; d < D; d += 32)
    out[static_cast<long long>(n) * D + d] = __float2half_rn((row[d] - mean) * rstd * gamma[d] + beta[d]);
}

// out[r, :] = in[r, :] / ||in[r, :]||_2   (one warp per row, fp32)
__global__ void l2norm_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int D) {
  pdl_grid_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= rows) return;
  const float* src = in + static_cast<long long>(r) * D;
  float sq = 0.f;
  for (int d = lane; d < D; d += 32) sq = fmaf(src[d], src[d], sq);
  const float nrm = sqrtf(warp_sum(sq));
  for (int d = lane; d < D; d += 32) out[static_cast<long long>(r) * D + d] = src[d] / nrm;
}

}  // namespace owl

using namespace owl;

extern "C" int owl_text_embed(const long long* ids, const float* tok_emb, const float* pos_emb, float* x, int rows,
                              int S, int D, int vocab, int* status, void* stream) {
  OWL_CHECK_ARG(ids && tok_emb && pos_emb && x && rows > 0 && S > 0 && vocab > 0, "text_embed: null / empty argument");
  OWL_CHECK_ARG(D > 0 && D % 4 == 0, "text_embed: D = %d must be a multiple of 4", D);
  OWL_LAUNCH(text_embed_kernel, static_cast<unsigned>(rows), 128, 0, static_cast<cudaStream_t>(stream), ids, tok_emb,
             pos_emb, x, rows, S, D, vocab, status);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_text_attn(const void* qkv_f16, const int* mask, void* ctx_f16, int N, int S, int H, int head_dim,
                             float scale, void* stream) {
  OWL_CHECK_ARG(qkv_f16 && ctx_f16 && N > 0 && H > 0, "text_attn: null / empty argument");
  OWL_CHECK_ARG(S > 0 && S <= TEXT_MAX_S, "text_attn: %d tokens per prompt (at most %d are built)", S, TEXT_MAX_S);
  OWL_CHECK_ARG(head_dim == TEXT_DH, "text_attn: head_dim %d is not built (only 64)", head_dim);
  const long long pairs = 1LL * N * H;
  OWL_LAUNCH(text_attn_kernel, static_cast<unsigned>((pairs + TEXT_ATTN_WARPS - 1) / TEXT_ATTN_WARPS),
             TEXT_ATTN_WARPS * 32, 0, static_cast<cudaStream_t>(stream), static_cast<const __half*>(qkv_f16), mask,
             static_cast<__half*>(ctx_f16), N, S, H, scale);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_text_pool_ln(const float* x, const long long* ids, const float* gamma, const float* beta,
                                void* out_f16, int N, int S, int D, float eps, void* stream) {
  OWL_CHECK_ARG(x && ids && gamma && beta && out_f16 && N > 0 && S > 0 && D > 0, "text_pool_ln: null / empty argument");
  OWL_LAUNCH(text_pool_ln_kernel, static_cast<unsigned>((N + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream), x, ids,
             gamma, beta, static_cast<__half*>(out_f16), N, S, D, eps);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_l2norm_rows(const float* in, float* out, int rows, int D, void* stream) {
  OWL_CHECK_ARG(in && out && rows > 0 && D > 0, "l2norm_rows: null / empty argument");
  OWL_LAUNCH(l2norm_rows_kernel, static_cast<unsigned>((rows + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream), in,
             out, rows, D);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
