// Fused attention forward for the ViT encoder (HF:379-404: softmax(q k^T / sqrt(dh)) v, no mask, no dropout)
// on tcgen05 tensor cores: scores never leave the SM.
//
//   CTA            = one (image, head, 128-query tile); 320 threads
//   warp 0         = TMA producer: Q tile once, then K/V blocks of 208 keys through a 3-stage ring
//   warp 1         = MMA issuer:   S_j = Q K_j^T  (128 x 208 x 64, operands in swizzled smem, fp32 in TMEM)
//                                  O  += P_j V_j  (128 x 64 x 208, A = P_j read straight from TMEM, B = V_j smem)
//   warps 2..9     = softmax: two threads per query row (TMEM lane = row; warps 2..5 take key columns 0..111 of a
//                    block, warps 6..9 columns 112..207, so every SM sub-partition interleaves two softmax warps):
//                    running max / sum in fp32 (row max exchanged through shared memory), exp2 on the MUFU, P_j
//                    written back over S_j in TMEM as packed fp16, O rescaled in TMEM when the max grows; final
//                    O / l written to ctx as fp16
//   TMEM (512 col) = S/P buffer 0 @0, S/P buffer 1 @224, O @448 (64 columns)
//
// 577 tokens = 4.5 query tiles and 2.77 key blocks: out-of-range rows are zero-filled by TMA (per-image bounds in
// the tensor map), out-of-range key columns are masked to -inf before the softmax.
#include "common.h"
#include "ptx.cuh"
#include <stdlib.h>

namespace owl {

constexpr int FA_BM = 128;       // queries per CTA
constexpr int FA_DH = 64;        // head dim
constexpr int FA_THREADS = 320;
constexpr int FA_Q_BYTES = FA_BM * FA_DH * 2;       // 16 KB

// Two tilings of the same kernel:
//   Wide   : 208 keys per block (3 blocks = 624 >= 577 tokens), two S buffers so that S_{j+1} is computed while the
//            softmax of block j runs, 3-stage K/V ring, all 512 TMEM columns, one CTA per SM.
//   Narrow : 128 keys per block (5 blocks), one S buffer, 2-stage ring, 256 TMEM columns and ~84 KB of shared
//            memory, so that TWO CTAs share an SM: their TMA / MMA / softmax phases interleave, which hides the
//            per-tile start-up and tail that the serial chain of a single CTA leaves exposed (measured timeline:
//            1.5 us start-up + 3 x 2.4 us softmax + 1.6 us tail per tile in the Wide tiling).
template <int BN_, int NSBUF_, int STAGES_, int SPLIT_, int MINCTAS_>
struct FaCfg {
  static constexpr int BN = BN_;           // keys per block (UMMA N, multiple of 16)
  static constexpr int NSBUF = NSBUF_;     // S/P buffers in TMEM
  static constexpr int STAGES = STAGES_;   // K/V ring depth
  static constexpr int SPLIT = SPLIT_;     // key columns [0, SPLIT) -> softmax half 0, [SPLIT, BN) -> half 1
  static constexpr int MINCTAS = MINCTAS_;
  static constexpr int KV_BYTES = BN * FA_DH * 2;   // multiple of 1024: every block stays swizzle-aligned
  static constexpr int SMEM = FA_Q_BYTES + 2 * STAGES * KV_BYTES + 1024 + 256 + 2 * 2 * 128 * 4;
  static constexpr uint32_t TMEM_COLS = NSBUF == 2 ? 512 : 256;
  static constexpr uint32_t TMEM_S0 = 0, TMEM_S1 = 224, TMEM_O = NSBUF == 2 ? 448 : 128;
  static_assert(BN % 16 == 0 && SPLIT % 16 == 0 && KV_BYTES % 1024 == 0, "tile shape");
  // TMEM column (relative to the S buffer) where the packed fp16 probabilities of key chunk c live
  static __host__ __device__ constexpr int p_col(int c) { return c < SPLIT / 16 ? 8 * c : SPLIT + 8 * (c - SPLIT / 16); }
};
using FaWide = FaCfg<208, 2, 3, 112, 1>;
using FaNarrow = FaCfg<128, 1, 2, 64, 2>;

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

template <class Cfg>
__global__ void __launch_bounds__(FA_THREADS, Cfg::MINCTAS)
flash_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      __half* __restrict__ ctx, int S, int D, float scale_log2, long long* __restrict__ dbg) {
  // dbg (development only): when non-null, CTA (0,0,0) and the last CTA record %globaltimer at phase boundaries
  auto stamp = [&](int slot) {
    if (dbg != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[slot] = t;
    }
  };
  constexpr int FA_BN = Cfg::BN, FA_STAGES = Cfg::STAGES, FA_KV_BYTES = Cfg::KV_BYTES, FA_SPLIT = Cfg::SPLIT;
  constexpr int NSBUF = Cfg::NSBUF;
  constexpr uint32_t FA_TMEM_S0 = Cfg::TMEM_S0, FA_TMEM_S1 = Cfg::TMEM_S1, FA_TMEM_O = Cfg::TMEM_O;
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
  float* xch = reinterpret_cast<float*>(kv_full) + 64;   // [2 block parities][2 halves][128] row max (and final row sum) exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (S + FA_BN - 1) / FA_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int s = 0; s < FA_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 256); }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 64) stamp(0);
  pdl_grid_wait();   // set-up above overlaps the previous kernel's tail
  if (threadIdx.x == 64) stamp(1);

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
        const uint32_t d = tmem_base + ((j % NSBUF) ? FA_TMEM_S1 : FA_TMEM_S0);
        const uint32_t bK = smem_u32(sK + st * FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_DH / 16; ++k)
          umma_f16(d, make_sdesc_sw128(aQ + k * 32, 0, 1024), make_sdesc_sw128(bK + k * 32, 0, 1024), IDESC_S,
                   k > 0 ? 1u : 0u);
        umma_commit(&s_full[j % NSBUF]);
      }
      __syncwarp();
    };
    if (NSBUF == 2) issue_s(0);
    for (int j = 0; j < n_blocks; ++j) {
      // two buffers: S_{j+1} runs ahead of the softmax of block j; one buffer: S_j follows P.V_{j-1} in issue order
      if (NSBUF == 2) { if (j + 1 < n_blocks) issue_s(j + 1); } else { issue_s(j); }
      mbar_wait(&p_full[j % NSBUF], (j / NSBUF) & 1);
      tc_fence_after();
      if (lane == 0) {
        const int st = j % FA_STAGES;
        const uint32_t aP = tmem_base + ((j % NSBUF) ? FA_TMEM_S1 : FA_TMEM_S0);
        const uint32_t bV = smem_u32(sV + st * FA_KV_BYTES);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)
          umma_f16_ts(tmem_base + FA_TMEM_O, aP + Cfg::p_col(k), make_sdesc_sw128(bV + k * 2048, 8192, 1024), IDESC_O,
                      (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ softmax / correction / epilogue: two threads per query row
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                 // 0: key columns [0, 112), 1: [112, 208) of every block
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    constexpr int C0 = FA_SPLIT / 16, C1 = FA_BN / 16; // chunk ranges: half 0 = [0, C0), half 1 = [C0, C1)
    const int cb = half ? C0 : 0, ce = half ? C1 : C0;
    auto sync_softmax = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_blocks; ++j) {
      const uint32_t sbuf = tmem_base + lane_addr + ((j % NSBUF) ? FA_TMEM_S1 : FA_TMEM_S0);
      const int valid = min(FA_BN, S - j * FA_BN);   // key columns of this block that exist
      const bool full = valid == FA_BN;               // only the last block of an image is partial
      mbar_wait(&s_full[j % NSBUF], (j / NSBUF) & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && j < 6) stamp(2 + 4 * j);
      // pass A: maximum over this thread's columns, then over the row
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
      for (int c0 = cb; c0 < ce; c0 += 2) {
        if (c0 * 16 >= valid) break;
        uint32_t r[32];
        const int nchunk = min(2, ce - c0);
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (q < nchunk) tmem_ld16(sbuf + (c0 + q) * 16, r + q * 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if (i < nchunk * 16) {
            float a = __uint_as_float(r[i]), b2 = __uint_as_float(r[i + 1]);
            if (!full) {
              a = (c0 * 16 + i < valid) ? a : -INFINITY;
              b2 = (c0 * 16 + i + 1 < valid) ? b2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, a);
            mx1 = fmaxf(mx1, b2);
          }
        }
      }
      float* xj = xch + (j & 1) * 256;   // double-buffered: block j + 1 must not overwrite what a slow partner still reads
      xj[half * 128 + row] = fmaxf(mx0, mx1);
      sync_softmax();
      const float m_new = fmaxf(m_run, fmaxf(xj[row], xj[128 + row]));
      const float mc = m_new * scale_log2;
      if (threadIdx.x == 64 && j < 6) stamp(3 + 4 * j);
      // pass B: p = exp2(s * c - m * c) as packed fp16, written IN PLACE over the columns this thread has already
      // consumed: key chunk c (16 fp32 columns at 16c) becomes 8 packed columns at fa_p_col(c), which always trails
      // the thread's own read position and never leaves its half.  The P.V MMAs address each chunk individually, so
      // P does not have to be contiguous.
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
      for (int c = cb; c < ce; ++c) {
        uint32_t r[16], pk[8];
        tmem_ld16(sbuf + c * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), scale_log2, -mc));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), scale_log2, -mc));
          if (!full) {
            p0 = (c * 16 + 2 * i < valid) ? p0 : 0.f;
            p1 = (c * 16 + 2 * i + 1 < valid) ? p1 : 0.f;
          }
          sum0 += p0;
          sum1 += p1;
          const __half2 hp = __floats2half2_rn(p0, p1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        tmem_st8(sbuf + Cfg::p_col(c), pk);
      }
      const float sum = sum0 + sum1;
      if (threadIdx.x == 64 && j < 6) stamp(4 + 4 * j);
      if (j > 0) {
        // O was accumulated against the old maximum: rescale it once P V_{j-1} has retired (32 columns per half)
        const float alpha = fast_exp2((m_run - m_new) * scale_log2);
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
          const uint32_t obuf = tmem_base + lane_addr + FA_TMEM_O + half * 32;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
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
      mbar_arrive(&p_full[j % NSBUF]);
      if (threadIdx.x == 64 && j < 6) stamp(5 + 4 * j);
    }
    // epilogue: ctx[b, q0 + row, h * 64 + half * 32 ..] = O / l   (l = sum of both halves' partial sums)
    mbar_wait(o_full, (n_blocks - 1) & 1);
    tc_fence_after();
    if (threadIdx.x == 64) stamp(30);
    sync_softmax();                 // everyone is past the last read of the row-max exchange
    xch[half * 128 + row] = l_run;
    sync_softmax();
    const float inv_l = 1.0f / (xch[row] + xch[128 + row]);
    const int q = q0 + row;
    __half* dst = ctx + (static_cast<long long>(b) * S + q) * D + h * FA_DH + half * 32;
    const uint32_t obuf = tmem_base + lane_addr + FA_TMEM_O + half * 32;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
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

  if (threadIdx.x == 64) stamp(31);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// 3-D fp16 tensor map over the packed QKV buffer: {3*D columns, S rows, B images}
int make_qkv_map(CUtensorMap* out, const void* qkv, int B, int S, int D, uint32_t box_rows) {
  return make_tensor_map_f16(out, qkv, 3ull * D, static_cast<uint64_t>(S), static_cast<uint64_t>(B), 3ull * D,
                             static_cast<uint64_t>(S) * 3ull * D, FA_DH, box_rows);
}

}  // namespace owl

using namespace owl;

static long long* g_fa_dbg = nullptr;
extern "C" void owl_flash_attn_debug(long long* dbg) { g_fa_dbg = dbg; }   // development hook, not in the public header

extern "C" int owl_flash_attn_fwd(const void* qkv_f16, void* ctx_f16, int B, int S, int H, int head_dim, float scale,
                                  void* stream) {
  OWL_CHECK_ARG(qkv_f16 && ctx_f16 && B > 0 && S > 0 && H > 0, "flash_attn_fwd: bad arguments");
  OWL_CHECK_ARG(head_dim == FA_DH, "flash_attn_fwd: head_dim %d is not built (only 64)", head_dim);
  const int D = H * head_dim;
  // tiling: Narrow (two CTAs per SM) unless OWL_FA_WIDE=1
  static int wide = -1;
  if (wide < 0) { const char* e = getenv("OWL_FA_WIDE"); wide = (e && e[0] == '1') ? 1 : 0; }
  CUtensorMap tmQ, tmKV;
  int rc = make_qkv_map(&tmQ, qkv_f16, B, S, D, FA_BM);
  if (rc) return rc;
  rc = make_qkv_map(&tmKV, qkv_f16, B, S, D, wide ? FaWide::BN : FaNarrow::BN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    OWL_CUDA(cudaFuncSetAttribute(flash_attn_fwd_kernel<FaWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, FaWide::SMEM));
    OWL_CUDA(cudaFuncSetAttribute(flash_attn_fwd_kernel<FaNarrow>, cudaFuncAttributeMaxDynamicSharedMemorySize, FaNarrow::SMEM));
    configured = true;
  }
  dim3 grid((S + FA_BM - 1) / FA_BM, H, B);
  const float sl2 = scale * 1.4426950408889634f;
  if (wide)
    OWL_LAUNCH(flash_attn_fwd_kernel<FaWide>, grid, FA_THREADS, FaWide::SMEM, static_cast<cudaStream_t>(stream), tmQ, tmKV,
               static_cast<__half*>(ctx_f16), S, D, sl2, g_fa_dbg);
  else
    OWL_LAUNCH(flash_attn_fwd_kernel<FaNarrow>, grid, FA_THREADS, FaNarrow::SMEM, static_cast<cudaStream_t>(stream), tmQ,
               tmKV, static_cast<__half*>(ctx_f16), S, D, sl2, g_fa_dbg);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
