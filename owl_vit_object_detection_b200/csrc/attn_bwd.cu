// Fused attention backward for the trainable encoder layer (autograd of HF:393-404), head_dim 64, on tcgen05:
// scores, probabilities and their gradients never touch HBM.
//
//   grid           = (key blocks of 128, heads, images); 160 threads; one CTA per SM (448 TMEM columns)
//   warp 0         = control (elected lane): TMA producer (K_j, V_j once; Q_i, dO_i through a 2-slot ring) and MMA
//                    issuer; per query tile i:
//                        S  = Q_i K_j^T        dP = dO_i V_j^T                  (128 x 128 x 64 each, fp32 in TMEM)
//                        dV += P^T dO_i        dK += dS^T Q_i                   (128 x 64 x 128, operands from smem)
//                        dQ_i = dS K_j                                          (128 x 64 x 128)
//   warps 1..4     = one thread per query row: P = exp(scale S - lse), dS = P (scale dP - delta) from TMEM, written
//                    as fp16 MMA operands into shared memory (128-byte swizzle); then dQ_i from TMEM -> fp32 atomics
//                    (16-byte reds) into dq32; at the end dV_j / dK_j from TMEM -> fp16 rows of dqkv
//   TMEM (512 col) = S @0 (128), dP @128 (128), dV @256 (64), dK @320 (64), dQ @384 (64)
//
// lse [B,H,S] is the natural-log log-sum-exp saved by the forward, delta [B,H,S] = scale * rowsum(dctx . ctx)
// (owl_attn_delta).  Rows / keys past the sequence end are zero-filled by TMA and masked to P = dS = 0.
// Replaces five batched K = 64 GEMMs whose epilogues wrote and re-read P and dS through HBM (~290 us -> see DESIGN.md).
#include "common.h"
#include "ptx.cuh"
#include <algorithm>

namespace owl {

constexpr int AB_T = 128;                 // queries per tile = keys per block
constexpr int AB_DH = 64;
constexpr int AB_THREADS = 160;
constexpr int AB_TILE_BYTES = AB_T * AB_DH * 2;          // 16 KB: one [128][64] fp16 operand tile
constexpr int AB_SMEM = 2 * AB_TILE_BYTES /*K, V*/ + 4 * AB_TILE_BYTES /*Q, dO x 2 slots*/ + 4 * AB_TILE_BYTES /*P, dS*/ +
                        1024 + 256;
constexpr uint32_t AB_TM_S = 0, AB_TM_DP = 128, AB_TM_DV = 256, AB_TM_DK = 320, AB_TM_DQ = 384;

__device__ __forceinline__ float ab_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ab_sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t ab_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const float* __restrict__ lse, const float* __restrict__ delta, __half* __restrict__ dqkv,
                float* __restrict__ dq32, int S, int D, float scale) {
  extern __shared__ uint8_t ab_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE_BYTES;
  uint8_t* sQ = sV + AB_TILE_BYTES;            // [2 slots]
  uint8_t* sDO = sQ + 2 * AB_TILE_BYTES;       // [2 slots]
  uint8_t* sP = sDO + 2 * AB_TILE_BYTES;       // [2 key panels of 64][128 queries][64 keys]
  uint8_t* sDS = sP + 2 * AB_TILE_BYTES;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(sDS + 2 * AB_TILE_BYTES);
  uint64_t* q_full = kv_full + 1;      // [2]
  uint64_t* q_empty = q_full + 2;      // [2]
  uint64_t* s_full = q_empty + 2;
  uint64_t* p_full = s_full + 1;
  uint64_t* dq_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z, H = gridDim.y;
  const int nq = (S + AB_T - 1) / AB_T;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(dq_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_wait();

  if (warp == 0) {
    // ------------------------------------------------ control warp
    constexpr uint32_t IDESC_S = make_idesc_f16(AB_T, AB_T, false, false);     // Q K^T, dO V^T
    constexpr uint32_t IDESC_KV = make_idesc_f16(AB_T, AB_DH, true, true);     // P^T dO, dS^T Q  (both operands MN-major)
    constexpr uint32_t IDESC_DQ = make_idesc_f16(AB_T, AB_DH, false, true);    // dS K
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    auto load_q = [&](int i) {
      const int sl = i & 1;
      mbar_wait(&q_empty[sl], ((i >> 1) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&q_full[sl], 2 * AB_TILE_BYTES);
        tma_load_3d(sQ + sl * AB_TILE_BYTES, &tmQKV, &q_full[sl], h * AB_DH, i * AB_T, b);
        tma_load_3d(sDO + sl * AB_TILE_BYTES, &tmDO, &q_full[sl], h * AB_DH, i * AB_T, b);
      }
      __syncwarp();
    };
    auto issue_s = [&](int i) {     // S = Q_i K^T, dP = dO_i V^T
      const int sl = i & 1;
      mbar_wait(&q_full[sl], (i >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t dQd = make_sdesc_sw128(smem_u32(sQ + sl * AB_TILE_BYTES), 0, 1024);
        const uint64_t dOd = make_sdesc_sw128(smem_u32(sDO + sl * AB_TILE_BYTES), 0, 1024);
        const uint64_t dKd = make_sdesc_sw128(smem_u32(sK), 0, 1024);
        const uint64_t dVd = make_sdesc_sw128(smem_u32(sV), 0, 1024);
#pragma unroll
        for (int k = 0; k < AB_DH / 16; ++k) umma_f16(tm + AB_TM_S, dQd + 2 * k, dKd + 2 * k, IDESC_S, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AB_DH / 16; ++k) umma_f16(tm + AB_TM_DP, dOd + 2 * k, dVd + 2 * k, IDESC_S, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(kv_full, 2 * AB_TILE_BYTES);
      tma_load_3d(sK, &tmQKV, kv_full, D + h * AB_DH, j0, b);
      tma_load_3d(sV, &tmQKV, kv_full, 2 * D + h * AB_DH, j0, b);
    }
    __syncwarp();
    load_q(0);
    if (nq > 1) load_q(1);
    mbar_wait(kv_full, 0);
    issue_s(0);
    for (int i = 0; i < nq; ++i) {
      const int sl = i & 1;
      mbar_wait(p_full, i & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        // MN-major A: [K = queries][64 keys] panels 16 KB apart (LBO), 8-row groups 1024 B apart (SBO), K step = 2048 B
        const uint64_t aP = make_sdesc_sw128(smem_u32(sP), 2 * AB_TILE_BYTES / 2, 1024);
        const uint64_t aDS = make_sdesc_sw128(smem_u32(sDS), 2 * AB_TILE_BYTES / 2, 1024);
        const uint64_t bDO = make_sdesc_sw128(smem_u32(sDO + sl * AB_TILE_BYTES), 8192, 1024);
        const uint64_t bQ = make_sdesc_sw128(smem_u32(sQ + sl * AB_TILE_BYTES), 8192, 1024);
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_f16(tm + AB_TM_DV, aP + (2048 >> 4) * k, bDO + (2048 >> 4) * k, IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k)
          umma_f16(tm + AB_TM_DK, aDS + (2048 >> 4) * k, bQ + (2048 >> 4) * k, IDESC_KV, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&q_empty[sl]);
        // dQ_i = dS K_j: A = dS K-major (two 64-key panels of [128 q][64 keys]), B = K_j MN-major (N = head dim)
        const uint64_t bK = make_sdesc_sw128(smem_u32(sK), 8192, 1024);
#pragma unroll
        for (int k = 0; k < AB_T / 16; ++k) {
          const uint64_t aK = make_sdesc_sw128(smem_u32(sDS) + (k >> 2) * AB_TILE_BYTES + (k & 3) * 32, 0, 1024);
          umma_f16(tm + AB_TM_DQ, aK, bK + (2048 >> 4) * k, IDESC_DQ, k > 0 ? 1u : 0u);
        }
        umma_commit(dq_full);
      }
      __syncwarp();
      if (i + 1 < nq) issue_s(i + 1);
      if (i + 2 < nq) load_q(i + 2);
    }
  } else {
    // ------------------------------------------------ one thread per query row (and, at the end, per key row)
    const int r = (warp & 3) * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const float scale_log2 = scale * 1.4426950408889634f;
    const uint32_t sp_row = smem_u32(sP) + r * 128, sds_row = smem_u32(sDS) + r * 128;
    for (int i = 0; i < nq; ++i) {
      const int q = i * AB_T + r;
      const bool q_ok = q < S;
      const long long li = (static_cast<long long>(b) * H + h) * S + (q_ok ? q : 0);
      const float lse2 = __ldg(lse + li) * 1.4426950408889634f;
      const float dl = __ldg(delta + li);
      mbar_wait(s_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32(trow + AB_TM_S + c * 32, sv);
        tmem_ld32(trow + AB_TM_DP + c * 32, dv);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float p0 = ab_exp2(fmaf(__uint_as_float(sv[2 * e]), scale_log2, -lse2));
          float p1 = ab_exp2(fmaf(__uint_as_float(sv[2 * e + 1]), scale_log2, -lse2));
          const int key = j0 + c * 32 + 2 * e;
          p0 = (q_ok && key < S) ? p0 : 0.f;
          p1 = (q_ok && key + 1 < S) ? p1 : 0.f;
          const float d0 = p0 * fmaf(__uint_as_float(dv[2 * e]), scale, -dl);
          const float d1 = p1 * fmaf(__uint_as_float(dv[2 * e + 1]), scale, -dl);
          pp[e] = ab_pack(p0, p1);
          dd[e] = ab_pack(d0, d1);
        }
        // 32 keys = 64 bytes = four 16-byte chunks of this row in key panel c / 2 (128-byte swizzle: chunk ^ (row & 7))
        const uint32_t pan = (c >> 1) * AB_TILE_BYTES;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t off = pan + ((((c & 1) * 4 + u) ^ (r & 7)) << 4);
          ab_sts128(sp_row + off, make_uint4(pp[4 * u], pp[4 * u + 1], pp[4 * u + 2], pp[4 * u + 3]));
          ab_sts128(sds_row + off, make_uint4(dd[4 * u], dd[4 * u + 1], dd[4 * u + 2], dd[4 * u + 3]));
        }
      }
      fence_proxy_async_smem();      // the operand tiles were written through the generic proxy
      tc_fence_before();
      mbar_arrive(p_full);
      // dQ_i (this key block's contribution) -> fp32 atomics
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      float* dst = dq32 + (static_cast<long long>(b) * S + q) * D + h * AB_DH;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(trow + AB_TM_DQ + c * 32, o);
        tmem_ld_wait();
        if (q_ok) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            atomicAdd(reinterpret_cast<float4*>(dst + c * 32 + 4 * u),
                      make_float4(__uint_as_float(o[4 * u]), __uint_as_float(o[4 * u + 1]), __uint_as_float(o[4 * u + 2]),
                                  __uint_as_float(o[4 * u + 3])));
        }
      }
      tc_fence_before();
    }
    // dV_j, dK_j: thread = key row (dq_full of the last tile completed after every MMA of this CTA)
    const int key = j0 + r;
    __half* row = dqkv + (static_cast<long long>(b) * S + min(key, S - 1)) * 3 * D + h * AB_DH;
#pragma unroll
    for (int w = 0; w < 2; ++w) {          // 0: dK -> column block D, 1: dV -> column block 2 D
      __half* dst = row + (w + 1) * D;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(trow + (w ? AB_TM_DV : AB_TM_DK) + c * 32, o);     // warp-collective: every lane takes part
        tmem_ld_wait();
        if (key < S) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 v;
            v.x = ab_pack(__uint_as_float(o[8 * u]), __uint_as_float(o[8 * u + 1]));
            v.y = ab_pack(__uint_as_float(o[8 * u + 2]), __uint_as_float(o[8 * u + 3]));
            v.z = ab_pack(__uint_as_float(o[8 * u + 4]), __uint_as_float(o[8 * u + 5]));
            v.w = ab_pack(__uint_as_float(o[8 * u + 6]), __uint_as_float(o[8 * u + 7]));
            *reinterpret_cast<uint4*>(dst + c * 32 + u * 8) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dst[r][0..cols) (fp16, row stride ld_dst) = fp16(src[r][0..cols)) (fp32, contiguous rows)
__global__ void cast_rows_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long rows, int cols,
                                     long long ld_dst) {
  pdl_grid_wait();
  const long long n4 = rows * (cols >> 2);
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n4; i += 1LL * gridDim.x * blockDim.x) {
    const long long r = i / (cols >> 2);
    const int c = static_cast<int>(i - r * (cols >> 2)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * cols + c);
    uint2 u;
    u.x = ab_pack(v.x, v.y);
    u.y = ab_pack(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = u;
  }
}

}  // namespace owl

using namespace owl;

extern "C" int owl_attn_bwd(const void* qkv_f16, const void* dctx_f16, const float* lse, const float* delta,
                            void* dqkv_f16, float* dq32, int B, int S, int H, int head_dim, float scale, void* stream) {
  OWL_CHECK_ARG(qkv_f16 && dctx_f16 && lse && delta && dqkv_f16 && dq32 && B > 0 && S > 0 && H > 0, "attn_bwd: bad arguments");
  OWL_CHECK_ARG(head_dim == AB_DH, "attn_bwd: head_dim %d is not built (only 64)", head_dim);
  const int D = H * head_dim;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tmQKV, tmDO;
  int rc = make_tensor_map_f16(&tmQKV, qkv_f16, 3ull * D, static_cast<uint64_t>(S), static_cast<uint64_t>(B), 3ull * D,
                               static_cast<uint64_t>(S) * 3ull * D, AB_DH, AB_T);
  if (rc) return rc;
  rc = make_tensor_map_f16(&tmDO, dctx_f16, static_cast<uint64_t>(D), static_cast<uint64_t>(S), static_cast<uint64_t>(B),
                           static_cast<uint64_t>(D), static_cast<uint64_t>(S) * D, AB_DH, AB_T);
  if (rc) return rc;
  static SmemOptIn optin;
  OWL_CUDA(ensure_smem(optin, attn_bwd_kernel, AB_SMEM));
  OWL_CUDA(cudaMemsetAsync(dq32, 0, sizeof(float) * static_cast<size_t>(B) * S * D, s));
  dim3 grid((S + AB_T - 1) / AB_T, H, B);
  OWL_LAUNCH(attn_bwd_kernel, grid, AB_THREADS, AB_SMEM, s, tmQKV, tmDO, lse, delta, static_cast<__half*>(dqkv_f16), dq32,
             S, D, scale);
  const long long rows = 1LL * B * S;
  OWL_LAUNCH(cast_rows_f16_kernel, static_cast<unsigned>(std::min<long long>((rows * (D / 4) + 255) / 256, 148LL * 16)), 256, 0,
             s, dq32, static_cast<__half*>(dqkv_f16), rows, D, 3LL * D);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
