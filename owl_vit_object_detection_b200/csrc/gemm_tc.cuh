// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[g][m][n] = sum_k A[g][m][k] * B[g][n][k]          (fp16 operands, fp32 accumulate in TMEM)
//
// * operands are staged global -> shared by TMA (3-D tiled tensor maps, 128-byte swizzle, OOB = 0),
//   through a multi-stage mbarrier ring;
// * one elected thread issues tcgen05.mma (128 x BN x 16 per instruction) into a double-buffered
//   TMEM accumulator (2 x BN columns);
// * four epilogue warps drain TMEM with tcgen05.ld and apply a fused epilogue functor
//   (bias / activation / residual / position embedding / max-pool / split-K reduction);
// * each operand may be K-major (reduction dim contiguous in memory) or MN-major (row index of the
//   GEMM contiguous in memory) - the second form is what dgrad (B = W as stored) and wgrad
//   (A = dY^T, B = X^T as stored) need, so no transposed copies are ever made;
// * a "batch" dimension g = (outer, head) addresses per-head attention matrices inside the packed
//   [tokens, 3*hidden] QKV buffer through the tensor map's third coordinate / a column offset.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
#pragma once
#include "ptx.cuh"

namespace owl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 fp16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_BUDGET = 200 * 1024;

struct GemmShape {
  int M, N, K;       // per-batch logical sizes
  int G, H;          // G batches; g -> (outer = g / H, head = g % H)
  int a_col_off;     // added to A's inner (memory-column) coordinate per head
  int a_sb, a_sh;    // A's 3rd tensor-map coordinate = outer * a_sb + head * a_sh
  int b_col_off, b_sb, b_sh;
  int split_k;       // K is cut into split_k slices (epilogue must then reduce atomically)
};

__host__ __device__ constexpr int gemm_stage_bytes(int BN) { return (GEMM_BM + BN) * GEMM_BK * 2; }
__host__ __device__ constexpr int gemm_num_stages(int BN) {
  return GEMM_SMEM_BUDGET / gemm_stage_bytes(BN) > 8 ? 8 : GEMM_SMEM_BUDGET / gemm_stage_bytes(BN);
}
__host__ __device__ constexpr int gemm_smem_bytes(int BN) {
  return gemm_num_stages(BN) * gemm_stage_bytes(BN) + 1024 /*align slack*/ + 256 /*barriers*/;
}
__host__ __device__ constexpr int gemm_tmem_cols(int BN) {
  return 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
}

// Descriptor of the k-th 16-wide K slice of an operand tile that starts at `base` in shared memory.
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t operand_desc(uint32_t base, int k16) {
  if constexpr (!MN_MAJOR) {
    // [rows][64 k] rows of 128 B, 8-row groups 1024 B apart; a K step of 16 elements = 32 B.
    return make_sdesc_sw128(base + k16 * 32, 0, 1024);
  } else {
    // chunks of 64 MN-elements: [64 k][64 mn] (8 KB each, LBO); 8-k groups 1024 B apart (SBO);
    // a K step of 16 = 16 rows of 128 B.
    return make_sdesc_sw128(base + k16 * 2048, 8192, 1024);
  }
}

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmShape gs, const typename Epi::Params ep) {
  constexpr int STAGES = gemm_num_stages(BN);
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = gemm_tmem_cols(BN);
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(!B_MN || BN % 64 == 0, "MN-major B tiles are loaded in 64-wide chunks");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int mb = (gs.M + GEMM_BM - 1) / GEMM_BM;
  const int nb = (gs.N + BN - 1) / BN;
  const int kb_total = (gs.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per = (kb_total + gs.split_k - 1) / gs.split_k;
  const int tiles_per_g = mb * nb * gs.split_k;
  const int num_tiles = tiles_per_g * gs.G;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int g = t / tiles_per_g;
        int r = t - g * tiles_per_g;
        const int ks = r / (mb * nb);
        r -= ks * (mb * nb);
        const int m_blk = r / nb, n_blk = r - (r / nb) * nb;
        const int outer = g / gs.H, head = g - outer * gs.H;
        const int a_c = head * gs.a_col_off, a_b = outer * gs.a_sb + head * gs.a_sh;
        const int b_c = head * gs.b_col_off, b_b = outer * gs.b_sb + head * gs.b_sh;
        const int m0 = m_blk * GEMM_BM, n0 = n_blk * BN;
        const int kb0 = ks * kb_per;
        const int kb1 = min(kb0 + kb_per, kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if constexpr (!A_MN) {
            tma_load_3d(sa, &tmA, &full_bar[stage], kb * GEMM_BK + a_c, m0, a_b);
          } else {
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c)
              tma_load_3d(sa + c * 8192, &tmA, &full_bar[stage], m0 + c * 64 + a_c, kb * GEMM_BK, a_b);
          }
          if constexpr (!B_MN) {
            tma_load_3d(sb, &tmB, &full_bar[stage], kb * GEMM_BK + b_c, n0, b_b);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_3d(sb + c * 8192, &tmB, &full_bar[stage], n0 + c * 64 + b_c, kb * GEMM_BK, b_b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC = make_idesc_f16(GEMM_BM, BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int r = t % tiles_per_g;
      const int ks = r / (mb * nb);
      const int kb0 = ks * kb_per;
      const int kb1 = min(kb0 + kb_per, kb_total);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            umma_f16(d_tmem, operand_desc<A_MN>(sa, k), operand_desc<B_MN>(sb, k), IDESC,
                     (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) umma_commit(&tfull_bar[as]);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ epilogue warps
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int g = t / tiles_per_g;
      int r = t - g * tiles_per_g;
      const int ks = r / (mb * nb);
      r -= ks * (mb * nb);
      const int m_blk = r / nb, n_blk = r - (r / nb) * nb;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quad * 32) << 16);
      const int kb0 = ks * kb_per;
      // a split whose K range is empty contributes nothing (can happen when split_k does not divide)
      const bool has_k = kb0 < kb_total;
      Epi::template run<BN>(ep, taddr, g, m_blk * GEMM_BM + row, n_blk * BN, gs.M, gs.N, has_k);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =================================================================== epilogues
__device__ __forceinline__ float act_qgelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float act_qgelu_grad(float x) {
  const float s = 1.0f / (1.0f + __expf(-1.702f * x));
  return s + 1.702f * x * s * (1.0f - s);
}
__device__ __forceinline__ float act_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float act_gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

enum : int { ACT_NONE = 0, ACT_QGELU = 1, ACT_GELU = 2, ACT_QGELU_GRAD = 3, ACT_GELU_GRAD = 4 };

// fp16 output:  v = alpha*acc + bias[n];  optionally save v (pre-activation) to `pre_out`;
//               then v = act(v)  or  v *= act'(dact_src[m][n]).
struct EpiF16 {
  struct Params {
    __half* out;
    __half* pre_out;
    const float* bias;
    const __half* dact_src;
    int ldo, ld_pre, ld_dact;
    long long o_sb, o_sh;  // element offsets per outer batch / head
    int H;
    int act;
    float alpha;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, uint32_t taddr, int g, int m, int n0, int M, int N,
                                             bool has_k) {
    const int outer = g / p.H, head = g - outer * p.H;
    const long long obase = outer * p.o_sb + head * p.o_sh;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n = n0 + c * 32;
      if (n >= N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      if (m >= M) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = has_k ? __uint_as_float(r[i]) * p.alpha : 0.0f;
      const int nv = min(32, N - n);
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) v[i] += __ldg(p.bias + n + i);
      }
      if (p.pre_out) {
        __half* po = p.pre_out + (long long)m * p.ld_pre + n;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) po[i] = __float2half_rn(v[i]);
      }
      if (p.act == ACT_QGELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = act_qgelu(v[i]);
      } else if (p.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = act_gelu(v[i]);
      } else if (p.act == ACT_QGELU_GRAD || p.act == ACT_GELU_GRAD) {
        const __half* ds = p.dact_src + (long long)m * p.ld_dact + n;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < nv) {
            const float x = __half2float(ds[i]);
            v[i] *= (p.act == ACT_QGELU_GRAD) ? act_qgelu_grad(x) : act_gelu_grad(x);
          }
        }
      }
      __half* o = p.out + obase + (long long)m * p.ldo + n;
      if (nv == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 pk;
          __half2 h0 = __floats2half2_rn(v[8 * i + 0], v[8 * i + 1]);
          __half2 h1 = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
          __half2 h2 = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]);
          __half2 h3 = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          reinterpret_cast<uint4*>(o)[i] = pk;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) o[i] = __float2half_rn(v[i]);
      }
    }
  }
};

// fp32 output:  v = alpha*acc + bias[n] + resid[m][n] + pos[(m % rows_per_img + 1)][n]
//   mode 0: out = v      mode 1: out += v      mode 2: atomicAdd(out, v)  (split-K / shared outputs)
//   rows_per_img > 0 remaps output row m -> m + m / rows_per_img + 1 (patch rows -> token rows, CLS first).
struct EpiF32 {
  struct Params {
    float* out;
    const float* bias;
    const float* resid;
    const float* pos;
    int ldo, ldr;
    long long o_sb, o_sh;
    int H;
    int mode;
    int rows_per_img;
    float alpha;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, uint32_t taddr, int g, int m, int n0, int M, int N,
                                             bool has_k) {
    const int outer = g / p.H, head = g - outer * p.H;
    const long long obase = outer * p.o_sb + head * p.o_sh;
    int mo = m;
    int prow = 0;
    if (p.rows_per_img > 0) {
      const int img = m / p.rows_per_img;
      prow = m - img * p.rows_per_img + 1;
      mo = m + img + 1;
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n = n0 + c * 32;
      if (n >= N) break;
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      if (m >= M) continue;
      if (!has_k && p.mode != 0) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = has_k ? __uint_as_float(r[i]) * p.alpha : 0.0f;
      const int nv = min(32, N - n);
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) v[i] += __ldg(p.bias + n + i);
      }
      if (p.resid) {
        const float* rs = p.resid + (long long)mo * p.ldr + n;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) v[i] += rs[i];
      }
      if (p.pos) {
        const float* ps = p.pos + (long long)prow * p.ldo + n;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) v[i] += __ldg(ps + i);
      }
      float* o = p.out + obase + (long long)mo * p.ldo + n;
      if (p.mode == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) atomicAdd(o + i, v[i]);
      } else {
        if (p.mode == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nv) v[i] += o[i];
        }
        if (nv == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nv) o[i] = v[i];
        }
      }
    }
  }
};

// Class-head tail (reference src/models.py:35-36): sims[m][c] = max_{j<3} acc[m][3c + j];
// also records which prompt variant won (for the backward scatter).
struct EpiPool3 {
  struct Params {
    float* sims;        // [M, C]
    uint8_t* argmax;    // [M, C]
    int C;
  };
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, uint32_t taddr, int g, int m, int n0, int M, int N,
                                             bool has_k) {
    (void)g; (void)has_k;
#pragma unroll 1
    for (int c = 0; c * 96 < BN; ++c) {
      const int n = n0 + c * 96;
      if (n >= N) break;
      uint32_t r[96];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (c * 96 + j * 32 < BN) tmem_ld32(taddr + c * 96 + j * 32, r + 32 * j);
      tmem_ld_wait();
      if (m >= M) continue;
      const int cls0 = n / 3;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = n + 3 * i;
        if (col + 2 < N && c * 96 + 3 * i + 2 < BN) {
          const float a = __uint_as_float(r[3 * i]), b = __uint_as_float(r[3 * i + 1]),
                      d = __uint_as_float(r[3 * i + 2]);
          // torch max_pool1d keeps the first maximum
          float best = a;
          int arg = 0;
          if (b > best) { best = b; arg = 1; }
          if (d > best) { best = d; arg = 2; }
          p.sims[(long long)m * p.C + cls0 + i] = best;
          p.argmax[(long long)m * p.C + cls0 + i] = static_cast<uint8_t>(arg);
        }
      }
    }
  }
};

}  // namespace owl
