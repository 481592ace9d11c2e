// Fused attention forward for the ViT encoder (HF:379-404: softmax(q k^T / sqrt(dh)) v, no mask, no dropout)
// on tcgen05 tensor cores: scores never leave the SM.
//
//   CTA            = one (image, head, 128-query tile); 192 threads
//   warp 0         = TMA producer: Q tile once, then K/V blocks of 208 keys through a 3-stage ring
//   warp 1         = MMA issuer:   S_j = Q K_j^T  (128 x 208 x 64, operands in swizzled smem, fp32 in TMEM)
//                                  O  += P_j V_j  (128 x 64 x 208, A = P_j read straight from TMEM, B = V_j smem)
//   warps 2..5     = softmax: thread t owns query row t (TMEM lane t): running max / sum in fp32, exp2 on the
//                    MUFU, P_j written back over S_j in TMEM as packed fp16, O rescaled in TMEM when the max grows;
//                    final O / l written to ctx as fp16
//   TMEM (512 col) = S/P buffer 0 @0, S/P buffer 1 @224, O @448 (64 columns)
//
// 577 tokens = 4.5 query tiles and 2.77 key blocks: out-of-range rows are zero-filled by TMA (per-image bounds in
// the tensor map), out-of-range key columns are masked to -inf before the softmax.
#include "common.h"
#include "ptx.cuh"

namespace owl {

constexpr int FA_BM = 128;       // queries per CTA
constexpr int FA_BN = 208;       // keys per block (UMMA N, multiple of 16; 3 blocks cover 577 tokens)
constexpr int FA_DH = 64;        // head dim
constexpr int FA_STAGES = 3;
constexpr int FA_THREADS = 192;
constexpr int FA_Q_BYTES = FA_BM * FA_DH * 2;       // 16 KB
constexpr int FA_KV_BYTES = FA_BN * FA_DH * 2;      // 26 KB (26 * 1024: keeps every block 1024-byte aligned)
constexpr int FA_SMEM = FA_Q_BYTES + 2 * FA_STAGES * FA_KV_BYTES + 1024 + 256;
constexpr uint32_t FA_TMEM_S0 = 0, FA_TMEM_S1 = 224, FA_TMEM_O = 448;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x 8 columns of packed fp16 pairs per K step of 16
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      __half* __restrict__ ctx, int S, int D, float scale_log2) {
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_Q_BYTES;
  uint8_t* sV = sK + FA_STAGES * FA_KV_BYTES;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(sV + FA_STAGES * FA_KV_BYTES);
  uint64_t* kv_empty = kv_full + FA_STAGES;
  uint64_t* s_full = kv_empty + FA_STAGES;   // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_full = p_full + 2;             // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (S + FA_BN - 1) / FA_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int s = 0; s < FA_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_wait();   // set-up above overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % FA_STAGES;
        const uint32_t ph = (j / FA_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * FA_KV_BYTES + (j == 0 ? FA_Q_BYTES : 0));
        if (j == 0) tma_load_3d(sQ, &tmQ, &kv_full[st], h * FA_DH, q0, b);
        tma_load_3d(sK + st * FA_KV_BYTES, &tmKV, &kv_full[st], D + h * FA_DH, j * FA_BN, b);
        tma_load_3d(sV + st * FA_KV_BYTES, &tmKV, &kv_full[st], 2 * D + h * FA_DH, j * FA_BN, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC_S = make_idesc_f16(FA_BM, FA_BN, false, false);
    constexpr uint32_t IDESC_O = make_idesc_f16(FA_BM, FA_DH, false, true);
    const uint32_t aQ = smem_u32(sQ);
    auto issue_s = [&](int j) {
      const int st = j % FA_STAGES;
      mbar_wait(&kv_full[st], (j / FA_STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d = tmem_base + ((j & 1) ? FA_TMEM_S1 : FA_TMEM_S0);
        const uint32_t bK = smem_u32(sK + st * FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_DH / 16; ++k)
          umma_f16(d, make_sdesc_sw128(aQ + k * 32, 0, 1024), make_sdesc_sw128(bK + k * 32, 0, 1024), IDESC_S,
                   k > 0 ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < n_blocks; ++j) {
      if (j + 1 < n_blocks) issue_s(j + 1);
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const int st = j % FA_STAGES;
        const uint32_t aP = tmem_base + ((j & 1) ? FA_TMEM_S1 : FA_TMEM_S0);
        const uint32_t bV = smem_u32(sV + st * FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)
          umma_f16_ts(tmem_base + FA_TMEM_O, aP + k * 8, make_sdesc_sw128(bV + k * 2048, 8192, 1024), IDESC_O,
                      (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ softmax / correction / epilogue: one query row per thread
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_blocks; ++j) {
      const uint32_t sbuf = tmem_base + lane_addr + ((j & 1) ? FA_TMEM_S1 : FA_TMEM_S0);
      const int valid = min(FA_BN, S - j * FA_BN);   // key columns of this block that exist
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const bool full = valid == FA_BN;   // warp-uniform; only the last block of an image is partial
      // pass A: block maximum (four 16-column loads in flight per wait)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < FA_BN; c0 += 64) {
        if (c0 >= valid) break;
        uint32_t r[64];
        const int nchunk = (FA_BN - c0) >= 64 ? 4 : (FA_BN - c0) / 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nchunk) tmem_ld16(sbuf + c0 + q * 16, r + q * 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          if (i < nchunk * 16) {
            float a = __uint_as_float(r[i]), b2 = __uint_as_float(r[i + 1]);
            if (!full) {
              a = (c0 + i < valid) ? a : -INFINITY;
              b2 = (c0 + i + 1 < valid) ? b2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, a);
            mx1 = fmaxf(mx1, b2);
          }
        }
      }
      const float m_new = fmaxf(m_run, fmaxf(mx0, mx1));
      const float mc = m_new * scale_log2;
      // pass B: p = exp2(s * c - m * c) as packed fp16, written over the start of the same buffer (the write
      // position 8c trails the read position 16c, and every thread only touches its own TMEM lane)
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < FA_BN / 32; ++c) {
        uint32_t r[32], pk[16];
        tmem_ld16(sbuf + c * 32, r);
        tmem_ld16(sbuf + c * 32 + 16, r + 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), scale_log2, -mc));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -mc));
          if (!full) {
            p0 = (c * 32 + 2 * i < valid) ? p0 : 0.f;
            p1 = (c * 32 + 2 * i + 1 < valid) ? p1 : 0.f;
          }
          sum0 += p0;
          sum1 += p1;
          const __half2 hp = __floats2half2_rn(p0, p1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        tmem_st16(sbuf + c * 16, pk);
      }
      {
        constexpr int c0 = (FA_BN / 32) * 32;   // the last 16 columns (208 = 6 * 32 + 16)
        static_assert(FA_BN - c0 == 16, "tail handling assumes BN % 32 == 16");
        uint32_t r[16], pk[8];
        tmem_ld16(sbuf + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), scale_log2, -mc));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -mc));
          if (!full) {
            p0 = (c0 + 2 * i < valid) ? p0 : 0.f;
            p1 = (c0 + 2 * i + 1 < valid) ? p1 : 0.f;
          }
          sum0 += p0;
          sum1 += p1;
          const __half2 hp = __floats2half2_rn(p0, p1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        tmem_st8(sbuf + c0 / 2, pk);
      }
      const float sum = sum0 + sum1;
      if (j > 0) {
        // O was accumulated against the old maximum: rescale it once P V_{j-1} has retired
        const float alpha = fast_exp2((m_run - m_new) * scale_log2);
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
          const uint32_t obuf = tmem_base + lane_addr + FA_TMEM_O;
#pragma unroll
          for (int c = 0; c < FA_DH / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(obuf + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(obuf + c * 16, r);
          }
        }
        l_run = l_run * alpha + sum;
      } else {
        l_run = sum;
      }
      m_run = m_new;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
    }
    // epilogue: ctx[b, q0 + row, h * 64 ..] = O / l
    mbar_wait(o_full, (n_blocks - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + row;
    __half* dst = ctx + (static_cast<long long>(b) * S + q) * D + h * FA_DH;
    const uint32_t obuf = tmem_base + lane_addr + FA_TMEM_O;
#pragma unroll
    for (int c = 0; c < FA_DH / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(obuf + c * 16, r);
      tmem_ld_wait();
      if (q < S) {
        uint4 o0, o1;
        __half2 t;
        t = __floats2half2_rn(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l); o0.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l); o0.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l); o0.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l); o0.w = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(r[8]) * inv_l, __uint_as_float(r[9]) * inv_l); o1.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(r[10]) * inv_l, __uint_as_float(r[11]) * inv_l); o1.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(r[12]) * inv_l, __uint_as_float(r[13]) * inv_l); o1.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(r[14]) * inv_l, __uint_as_float(r[15]) * inv_l); o1.w = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint4*>(dst + c * 16) = o0;
        *reinterpret_cast<uint4*>(dst + c * 16 + 8) = o1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// 3-D fp16 tensor map over the packed QKV buffer: {3*D columns, S rows, B images}
int make_qkv_map(CUtensorMap* out, const void* qkv, int B, int S, int D, uint32_t box_rows) {
  return make_tensor_map_f16(out, qkv, 3ull * D, static_cast<uint64_t>(S), static_cast<uint64_t>(B), 3ull * D,
                             static_cast<uint64_t>(S) * 3ull * D, FA_DH, box_rows);
}

}  // namespace owl

using namespace owl;

extern "C" int owl_flash_attn_fwd(const void* qkv_f16, void* ctx_f16, int B, int S, int H, int head_dim, float scale,
                                  void* stream) {
  OWL_CHECK_ARG(qkv_f16 && ctx_f16 && B > 0 && S > 0 && H > 0, "flash_attn_fwd: bad arguments");
  OWL_CHECK_ARG(head_dim == FA_DH, "flash_attn_fwd: head_dim %d is not built (only 64)", head_dim);
  const int D = H * head_dim;
  CUtensorMap tmQ, tmKV;
  int rc = make_qkv_map(&tmQ, qkv_f16, B, S, D, FA_BM);
  if (rc) return rc;
  rc = make_qkv_map(&tmKV, qkv_f16, B, S, D, FA_BN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    OWL_CUDA(cudaFuncSetAttribute(flash_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    configured = true;
  }
  dim3 grid((S + FA_BM - 1) / FA_BM, H, B);
  OWL_LAUNCH(flash_attn_fwd_kernel, grid, FA_THREADS, FA_SMEM, static_cast<cudaStream_t>(stream), 
      tmQ, tmKV, static_cast<__half*>(ctx_f16), S, D, scale * 1.4426950408889634f);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
