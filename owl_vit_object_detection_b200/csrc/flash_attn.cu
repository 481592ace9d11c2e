// Fused attention forward for the ViT encoder (HF:379-404: softmax(q k^T / sqrt(dh)) v, no mask, no dropout)
// on tcgen05 tensor cores: scores never leave the SM.
//
//   CTA            = one (image, head, 128-query tile); 288 threads (112 registers each); TWO CTAs per SM
//   warp 0         = control (one elected lane): TMA producer (Q tile once, K/V blocks of 128 keys through a
//                    2-stage ring) and MMA issuer:
//                                  S_j = Q K_j^T  (128 x 128 x 64, operands in swizzled smem, fp32 in TMEM)
//                                  O  += P_j V_j  (128 x 64 x 128, A = P_j read straight from TMEM, B = V_j smem)
//   warps 1..8     = softmax: two threads per query row (TMEM lane = row; warps 1..4 take key columns 0..63 of a
//                    block, warps 5..8 columns 64..127, so every SM sub-partition interleaves two softmax warps)
//   TMEM (256 col) = S @0 (128 fp32 columns), O @128 (64), P @192 (128 keys as packed fp16 = 64 columns)
//
// What bounds this kernel at head_dim 64 is not the tensor pipe: per 128 x 128 block the two MMAs take ~512 clk,
// while reading S out of TMEM once (64 KB at 64 B/clk/SM) and the 16384 exp2 on the 16-lane MUFU take ~1024 clk
// each.  So S is read ONCE per block: every block after the first is exponentiated optimistically against the
// running maximum while its own maximum is tracked on the side; only if some row's maximum grew by more than 2^8
// (exponent domain) is the block redone against the new maximum and O rescaled (S is still intact because P has
// its own TMEM columns).  Probabilities therefore stay <= 256 (fp16-exact range), sums are fp32, and the final
// O / l removes the common factor.
//
// 577 tokens = 4.5 query tiles and 4.5 key blocks: out-of-range rows are zero-filled by TMA (per-image bounds in
// the tensor map), out-of-range key columns are masked to -inf; the last block only computes the 16-key chunks
// that hold a key.
#include "common.h"
#include "ptx.cuh"
#include <stdlib.h>

namespace owl {

constexpr int FA_BM = 128;       // queries per CTA
constexpr int FA_DH = 64;        // head dim
constexpr int FA_THREADS = 288;   // control warp + 8 softmax warps
constexpr int FA_Q_BYTES = FA_BM * FA_DH * 2;       // 16 KB

template <int BN_, int STAGES_, int MINCTAS_>
struct FaCfg {
  static constexpr int BN = BN_;           // keys per block (UMMA N, multiple of 16)
  static constexpr int STAGES = STAGES_;   // K/V ring depth
  static constexpr int SPLIT = BN_ / 2;    // key columns [0, SPLIT) -> softmax half 0, [SPLIT, BN) -> half 1
  static constexpr int MINCTAS = MINCTAS_;
  static constexpr int KV_BYTES = BN * FA_DH * 2;   // multiple of 1024: every block stays swizzle-aligned
  static constexpr int SMEM = FA_Q_BYTES + 2 * STAGES * KV_BYTES + 1024 + 256 + 2 * 2 * 128 * 4;
  static constexpr uint32_t TMEM_COLS = 256;
  static constexpr uint32_t TMEM_S = 0, TMEM_O = BN, TMEM_P = BN + FA_DH;
  static_assert(BN == 128 && TMEM_P + BN / 2 <= TMEM_COLS && KV_BYTES % 1024 == 0, "tile shape");
};
using FaNarrow = FaCfg<128, 2, 2>;

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
  constexpr uint32_t FA_TMEM_S = Cfg::TMEM_S, FA_TMEM_O = Cfg::TMEM_O, FA_TMEM_P = Cfg::TMEM_P;
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_Q_BYTES;
  uint8_t* sV = sK + FA_STAGES * FA_KV_BYTES;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(sV + FA_STAGES * FA_KV_BYTES);
  uint64_t* kv_empty = kv_full + FA_STAGES;
  uint64_t* s_full = kv_empty + FA_STAGES;   // [1] (slot [1] unused)
  uint64_t* p_full = s_full + 2;             // [1]
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
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
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
  if (dbg != nullptr && threadIdx.x == 64) {
    const long long cta = blockIdx.x + gridDim.x * (blockIdx.y + 1LL * gridDim.y * blockIdx.z);
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[64 + 3 * cta] = t;
  }

  if (warp == 0) {
    // ------------------------------------------------ control warp: TMA producer + MMA issuer (one elected lane).
    // One warp for both roles keeps the CTA at 9 warps, i.e. 112 registers per thread with two CTAs per SM, which
    // is what lets a softmax thread hold its 64 scores of a block in registers.
    constexpr uint32_t IDESC_S = make_idesc_f16(FA_BM, FA_BN, false, false);
    constexpr uint32_t IDESC_O = make_idesc_f16(FA_BM, FA_DH, false, true);
    const uint32_t aQ = smem_u32(sQ);
    auto load_kv = [&](int j) {   // K/V block j -> ring slot j % STAGES (+ the Q tile with block 0)
      const int st = j % FA_STAGES;
      mbar_arrive_expect_tx(&kv_full[st], 2 * FA_KV_BYTES + (j == 0 ? FA_Q_BYTES : 0));
      if (j == 0) tma_load_3d(sQ, &tmQ, &kv_full[st], h * FA_DH, q0, b);
      tma_load_3d(sK + st * FA_KV_BYTES, &tmKV, &kv_full[st], D + h * FA_DH, j * FA_BN, b);
      tma_load_3d(sV + st * FA_KV_BYTES, &tmKV, &kv_full[st], 2 * D + h * FA_DH, j * FA_BN, b);
    };
    auto issue_s = [&](int j) {   // S_j = Q K_j^T
      const int st = j % FA_STAGES;
      mbar_wait(&kv_full[st], (j / FA_STAGES) & 1);
      tc_fence_after();
      const uint32_t bK = smem_u32(sK + st * FA_KV_BYTES);
#pragma unroll
      for (int k = 0; k < FA_DH / 16; ++k)
        umma_f16(tmem_base + FA_TMEM_S, make_sdesc_sw128(aQ + k * 32, 0, 1024), make_sdesc_sw128(bK + k * 32, 0, 1024),
                 IDESC_S, k > 0 ? 1u : 0u);
      umma_commit(s_full);
    };
    if (lane == 0) {
      for (int j = 0; j < FA_STAGES && j < n_blocks; ++j) load_kv(j);
      issue_s(0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % FA_STAGES;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t bV = smem_u32(sV + st * FA_KV_BYTES);
        // chunks of 16 keys; the last block of an image stops at the chunk that still holds a key
        const int nk = (min(FA_BN, S - j * FA_BN) + 15) / 16;
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)
          if (k < nk)
            umma_f16_ts(tmem_base + FA_TMEM_O, tmem_base + FA_TMEM_P + 8 * k, make_sdesc_sw128(bV + k * 2048, 8192, 1024),
                        IDESC_O, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
        // S_{j+1} follows P V_j in issue order; nothing reads S_j any more (the softmax threads arrived on p_full)
        if (j + 1 < n_blocks) issue_s(j + 1);
        // refill the slot P V_j is draining; the load has the whole softmax of block j + 1 to land
        if (j + FA_STAGES < n_blocks) {
          mbar_wait(&kv_empty[st], (j / FA_STAGES) & 1);
          load_kv(j + FA_STAGES);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ softmax / correction / epilogue: two threads per query row
    const int quad = warp & 3;                        // TMEM lane quadrant a warp may touch = warp id % 4
    const int half = (warp - 1) >> 2;                 // 0: key columns [0, 64), 1: [64, 128) of every block
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    static_assert(FA_SPLIT == 64, "softmax code below: 64 key columns per thread, two groups of 32");
    auto sync_softmax = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    constexpr float kLazy = 8.0f;
    const uint32_t sbuf = tmem_base + lane_addr + FA_TMEM_S + half * FA_SPLIT;
    const uint32_t pbuf = tmem_base + lane_addr + FA_TMEM_P + half * (FA_SPLIT / 2);
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_blocks; ++j) {
      const int valid = min(FA_BN, S - j * FA_BN);    // key columns of this block that exist
      const int my_valid = valid - half * FA_SPLIT;   // ... among this thread's 64 (<= 0: none)
      float* xj = xch + (j & 1) * 256;   // double-buffered: block j + 1 must not overwrite what a slow partner still reads
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && j < 6) stamp(2 + 4 * j);
      if (j == 0) {
        // first block: the maximum has to be known before anything can be exponentiated
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g * 32 < my_valid) {
            uint32_t r[32];
            tmem_ld32(sbuf + g * 32, r);
            tmem_ld_wait();
            if (my_valid >= g * 32 + 32) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(r[i]));     mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (g * 32 + i < my_valid) mx0 = fmaxf(mx0, __uint_as_float(r[i]));
            }
          }
        }
        xj[half * 128 + row] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        sync_softmax();
        m_run = fmaxf(xj[row], xj[128 + row]);        // finite: key 0 always exists
        if (threadIdx.x == 64) stamp(3);
      }
      // Blocks after the first: optimistic single pass against the running maximum; the block's own maximum is
      // checked afterwards and, rarely, the pass is repeated against the new maximum (attempt 1).
      float m_new = m_run, sum = 0.f;
      bool regrown = false;
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        // one pass over this thread's (up to) 64 scores, 16 at a time with the next chunk's TMEM load in flight
        // behind the current chunk's arithmetic: P = exp2(s * c - mc) as packed fp16 (key chunk of 16 -> 8 columns)
        const float mc = m_new * scale_log2;
        float sum0 = 0.f, sum1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY;
        uint32_t ra[16], rb[16];
        if (my_valid > 0) {
          tmem_ld16(sbuf, ra);
          tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 16 < my_valid) {       // warp-uniform; chunks without keys are skipped by the P.V MMAs as well
            uint32_t* rc = (c & 1) ? rb : ra;
            uint32_t* rn = (c & 1) ? ra : rb;
            if (c + 1 < 4 && (c + 1) * 16 < my_valid) tmem_ld16(sbuf + (c + 1) * 16, rn);
            if (my_valid < c * 16 + 16) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c * 16 + i >= my_valid) rc[i] = 0xff800000u;   // -inf -> p = 0
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float s0 = __uint_as_float(rc[2 * i]), s1 = __uint_as_float(rc[2 * i + 1]);
              mx0 = fmaxf(mx0, s0);
              mx1 = fmaxf(mx1, s1);
              const float p0 = fast_exp2(fmaf(s0, scale_log2, -mc));
              const float p1 = fast_exp2(fmaf(s1, scale_log2, -mc));
              sum0 += p0;
              sum1 += p1;
              const __half2 hp = __floats2half2_rn(p0, p1);
              pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
            }
            tmem_st8(pbuf + c * 8, pk);
            tmem_ld_wait();              // chunk c + 1 has landed behind the arithmetic above
          }
        }
        sum = sum0 + sum1;
        if (j == 0 || attempt == 1) break;
        xj[half * 128 + row] = fmaxf(mx0, mx1);
        sync_softmax();
        const float m_blk = fmaxf(xj[row], xj[128 + row]);
        const bool grow = (m_blk - m_run) * scale_log2 > kLazy;   // same verdict in both threads of a row
        if (threadIdx.x == 64 && j < 6) stamp(3 + 4 * j);
        if (!__any_sync(0xffffffffu, grow)) break;
        // rows that did not grow reproduce the same values in the second attempt
        m_new = grow ? m_blk : m_run;
        regrown = true;
        tmem_st_wait();
      }
      if (regrown) {
        // O was accumulated against the old maximum: rescale it once P V_{j-1} has retired (32 columns per half)
        const float alpha = fast_exp2((m_run - m_new) * scale_log2);   // 1 for rows that did not grow
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        const uint32_t obuf = tmem_base + lane_addr + FA_TMEM_O + half * 32;
        uint32_t o[32];
        tmem_ld32(obuf, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(obuf, o);
        l_run *= alpha;
        m_run = m_new;
      }
      l_run += sum;
      if (threadIdx.x == 64 && j < 6) stamp(4 + 4 * j);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
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
  if (dbg != nullptr && threadIdx.x == 64) {   // development only: per-CTA [start, end, sm] after the 64 phase stamps
    const long long cta = blockIdx.x + gridDim.x * (blockIdx.y + 1LL * gridDim.y * blockIdx.z);
    long long t;
    uint32_t sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    dbg[64 + 3 * cta + 1] = t;
    dbg[64 + 3 * cta + 2] = sm;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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
  CUtensorMap tmQ, tmKV;
  int rc = make_qkv_map(&tmQ, qkv_f16, B, S, D, FA_BM);
  if (rc) return rc;
  rc = make_qkv_map(&tmKV, qkv_f16, B, S, D, FaNarrow::BN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    OWL_CUDA(cudaFuncSetAttribute(flash_attn_fwd_kernel<FaNarrow>, cudaFuncAttributeMaxDynamicSharedMemorySize, FaNarrow::SMEM));
    configured = true;
  }
  dim3 grid((S + FA_BM - 1) / FA_BM, H, B);
  const float sl2 = scale * 1.4426950408889634f;
  OWL_LAUNCH(flash_attn_fwd_kernel<FaNarrow>, grid, FA_THREADS, FaNarrow::SMEM, static_cast<cudaStream_t>(stream), tmQ,
               tmKV, static_cast<__half*>(ctx_f16), S, D, sl2, g_fa_dbg);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
