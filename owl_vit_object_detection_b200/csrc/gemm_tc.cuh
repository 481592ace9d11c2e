// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[g][m][n] = sum_k A[g][m][k] * B[g][n][k]          (fp16 operands, fp32 accumulate in TMEM)
//
// * operands are staged global -> shared by TMA (3-D tiled tensor maps, 128-byte swizzle, OOB = 0),
//   through a multi-stage mbarrier ring;
// * one elected thread issues tcgen05.mma (128 x BN x 16 per instruction) into a double-buffered
//   TMEM accumulator (2 x BN columns);
// * four epilogue warps drain TMEM with tcgen05.ld and apply a fused epilogue functor
//   (bias / activation / residual / position embedding / max-pool / split-K reduction); every global
//   read and write of the epilogue is transposed through a per-warp swizzled shared-memory tile so that
//   each warp-level access covers whole 128-byte lines;
// * each operand may be K-major (reduction dim contiguous in memory) or MN-major (row index of the
//   GEMM contiguous in memory) - the second form is what dgrad (B = W as stored) and wgrad
//   (A = dY^T, B = X^T as stored) need, so no transposed copies are ever made;
// * a "batch" dimension g = (outer, head) addresses per-head attention matrices inside the packed
//   [tokens, 3*hidden] QKV buffer through the tensor map's third coordinate / a column offset.
//
// Warp roles: warps 0 .. E-1 = epilogue (E = gemm_epi_warps(BN); TMEM lane quadrant = warp_idx % 4, column part =
// warp_idx / 4, <= 64 columns each), warp E = TMA producer, warp E+1 = MMA issuer + TMEM owner.
#pragma once
#include "ptx.cuh"

namespace owl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 fp16 = one 128-byte swizzle row
constexpr int GEMM_SMEM_LIMIT = 227 * 1024;   // opt-in shared memory per CTA on sm_100a

// Epilogue warps of a tile: every warp drains <= 64 accumulator columns of one TMEM lane quadrant, so a 256-wide tile
// has sixteen of them (four per scheduler), a 192-wide one twelve.  With eight warps per tile (two per scheduler) the
// epilogue of the K = 768 layer GEMMs - a serial chain of tcgen05.ld / MUFU / shared-memory transposition / global
// store latencies - took ~11k clk per 128 x 256 tile against ~6k clk of MMAs, i.e. the tensor pipe idled 45 % of the
// time waiting for a free accumulator (ncu: sm__pipe_tensor_cycles_active 55 %).
__host__ __device__ constexpr int gemm_epi_warps(int BN) { return BN >= 128 ? 4 * (BN / 64) : 8; }
__host__ __device__ constexpr int gemm_threads(int BN) { return 64 + 32 * gemm_epi_warps(BN); }

struct GemmShape {
  int M, N, K;       // per-batch logical sizes
  int G, H;          // G batches; g -> (outer = g / H, head = g % H)
  int a_col_off;     // added to A's inner (memory-column) coordinate per head
  int a_sb, a_sh;    // A's 3rd tensor-map coordinate = outer * a_sb + head * a_sh
  int b_col_off, b_sb, b_sh;
  int split_k;       // K is cut into split_k slices (epilogue must then reduce atomically)
  int debug;         // dev only (OWL_GEMM_DEBUG): 1 = epilogue does nothing, 2 = no TMA loads (MMAs run on stale smem),
                     // 4 = no loads and no MMAs (epilogue only)
};

constexpr int GEMM_EPI_STAGE_BYTES = 4096;  // per epilogue warp: 32 rows x 128 B transposition buffer
constexpr int GEMM_EPI_BIAS_BYTES = 256;    // per epilogue warp: the bias of its <= 64 columns (broadcast reads)
// BN = N tile of the (cluster) tile, CM = CTAs of the cluster along M (each stages BN / CM rows of B)
__host__ __device__ constexpr int gemm_stage_bytes(int BN, int CM = 1) { return (GEMM_BM + BN / CM) * GEMM_BK * 2; }
// MINB = CTAs per SM the kernel is built for.  2 is the flavour for short-K, epilogue-dominated problems (the
// attention-backward GEMMs, K = 64: one k-block per tile): a 64 KB operand ring and <= 102 registers let two CTAs
// share an SM.
__host__ __device__ constexpr int gemm_num_stages(int BN, int CM = 1, int MINB = 1) {
  const int budget = MINB == 2 ? 64 * 1024
                               : GEMM_SMEM_LIMIT - gemm_epi_warps(BN) * (GEMM_EPI_STAGE_BYTES + GEMM_EPI_BIAS_BYTES) - 1024 - 512;
  const int n = budget / gemm_stage_bytes(BN, CM);
  return n > 8 ? 8 : n;
}
__host__ __device__ constexpr int gemm_smem_bytes(int BN, int CM = 1, int MINB = 1) {
  return gemm_num_stages(BN, CM, MINB) * gemm_stage_bytes(BN, CM) +
         gemm_epi_warps(BN) * (GEMM_EPI_STAGE_BYTES + GEMM_EPI_BIAS_BYTES) + 1024 /*align slack*/ + 512 /*barriers*/;
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

// CM = 2 selects the CTA-pair flavour (tcgen05 cta_group::2): the two CTAs of a cluster own two consecutive 128-row
// M tiles of the SAME N tile.  One thread of the leader CTA issues M = 256 MMAs that read each CTA's A tile and each
// CTA's HALF of the B tile from that CTA's own shared memory, so per CTA both the TMA fill and the tensor-core
// operand fetch of B are halved.  That matters because a single-CTA 128x256 tile needs 96 B/clk of operand reads
// plus 96 B/clk of TMA writes against the 128 B/clk shared-memory port (measured: 1.04 PFLOP/s = 63 % of cuBLAS),
// while the pair needs 64 + 64.
//   full_bar   leader only: its own arrive.expect_tx(2 x stage bytes) + complete_tx from both CTAs' TMA loads
//   empty_bar  per CTA, count 1: the leader's tcgen05.commit is multicast to both CTAs
//   tfull_bar  per CTA, count 1: same multicast commit after the last k-block of a tile
//   tempty_bar leader only, count CM x epilogue warps: the peer's epilogue warps arrive remotely
template <int BN, bool A_MN, bool B_MN, class Epi, int CM = 1, int MINB = 1>
__global__ void __launch_bounds__(gemm_threads(BN), MINB)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmShape gs, const __grid_constant__ typename Epi::Params ep) {
  static_assert(CM == 1 || CM == 2, "1 = single CTA, 2 = CTA pair (cta_group::2)");
  static_assert(CM == 1 || !A_MN, "the pair flavour is built for K-major A");
  static_assert(!B_MN || (BN / 64) % CM == 0, "MN-major B: whole 64-wide chunks per CTA");
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int BN_CTA = BN / CM;                       // B rows held by one CTA
  constexpr int B_BYTES = BN_CTA * GEMM_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = gemm_num_stages(BN, CM, MINB);
  constexpr int EPI_WARPS = gemm_epi_warps(BN);
  constexpr int PARTS = EPI_WARPS / 4;                  // column parts of a tile, one epilogue warp per (quadrant, part)
  static_assert(STAGES >= 2, "operand ring too small");
  static_assert(MINB == 1 || (CM == 1 && 2 * gemm_tmem_cols(BN) <= 512), "two CTAs per SM: both accumulators must fit TMEM");
  constexpr uint32_t TMEM_COLS = gemm_tmem_cols(BN);
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(!B_MN || BN % 64 == 0, "MN-major B tiles are loaded in 64-wide chunks");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;   // 4 KB per epilogue warp
  uint8_t* epi_bias = epi_stage + EPI_WARPS * GEMM_EPI_STAGE_BYTES;   // 256 B per epilogue warp
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_bias + EPI_WARPS * GEMM_EPI_BIAS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint64_t* epi_bar = tempty_bar + 2;         // [EPI_WARPS] per epilogue warp: TMA loads into its staging tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Warp roles: the epilogue warps come first (warp w drains TMEM lane quadrant w % 4, column part w / 4), the two
  // single-thread control warps LAST: the sub-partition arbiter prefers the highest warp id among eligible warps, and
  // the MMA issuer's handful of instructions per k-block must never queue behind four busy epilogue warps.
  constexpr int WARP_TMA = EPI_WARPS, WARP_MMA = EPI_WARPS + 1;

  const int mb = (gs.M + GEMM_BM - 1) / GEMM_BM;
  const int nb = (gs.N + BN - 1) / BN;
  const int kb_total = (gs.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per = (kb_total + gs.split_k - 1) / gs.split_k;
  const int mbc = (mb + CM - 1) / CM;                       // M tiles in units of clusters
  const int tiles_per_g = mbc * nb * gs.split_k;            // cluster tiles per batch
  const int num_tiles = tiles_per_g * gs.G;
  const int crank = CM > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = crank == 0;
  const int cluster_id = blockIdx.x / CM, num_clusters = gridDim.x / CM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS * CM);
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&epi_bar[w], 1);
    fence_barrier_init();
  }
  if constexpr (CM > 1) cluster_sync_all();   // both CTAs are resident with initialised barriers before pair ops
  if (warp == WARP_MMA) {
    if constexpr (CM == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    else { tmem_alloc_2sm(tmem_slot, TMEM_COLS); tmem_relinquish_2sm(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CM > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_wait();   // everything above overlapped the previous kernel's tail; global memory is touched from here on

  if (warp == WARP_TMA) {
    // ------------------------------------------------ TMA producer (every CTA fills its own shared memory)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int g = t / tiles_per_g;
        int r = t - g * tiles_per_g;
        const int ks = r / (mbc * nb);
        r -= ks * (mbc * nb);
        const int m_blk = (r / nb) * CM + crank, n_blk = r - (r / nb) * nb;
        const int outer = g / gs.H, head = g - outer * gs.H;
        const int a_c = head * gs.a_col_off, a_b = outer * gs.a_sb + head * gs.a_sh;
        const int b_c = head * gs.b_col_off, b_b = outer * gs.b_sb + head * gs.b_sh;
        const int m0 = m_blk * GEMM_BM;
        const int n0 = n_blk * BN + crank * BN_CTA;          // this CTA's slice of the B tile
        const int kb0 = ks * kb_per;
        const int kb1 = (gs.debug & 4) ? kb0 : min(kb0 + kb_per, kb_total);   // dev: 4 = epilogue only
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (gs.debug & 2) {   // dev: hand the slot over without loading anything
            if (CM == 1 || leader) mbar_arrive(&full_bar[stage]);
          } else if constexpr (CM == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
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
          } else {
            // the leader's barrier collects the bytes of both CTAs
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            tma_load_3d_2sm(sa, &tmA, &full_bar[stage], kb * GEMM_BK + a_c, m0, a_b);
            if constexpr (!B_MN) {
              tma_load_3d_2sm(sb, &tmB, &full_bar[stage], kb * GEMM_BK + b_c, n0, b_b);
            } else {
#pragma unroll
              for (int c = 0; c < BN_CTA / 64; ++c)
                tma_load_3d_2sm(sb + c * 8192, &tmB, &full_bar[stage], n0 + c * 64 + b_c, kb * GEMM_BK, b_b);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ------------------------------------------------ MMA issuer (pair flavour: leader CTA only)
    if (CM == 1 || leader) {
      constexpr uint32_t IDESC = make_idesc_f16(GEMM_BM * CM, BN, A_MN, B_MN);
      // The whole warp walks the loop with warp-uniform state; one elected lane issues.  Descriptors are built once
      // and advanced by adding to their 14-bit (address >> 4) field: the issue loop must stay far below the
      // 512 clk a k-block of MMAs takes, or the issuing thread itself becomes the bottleneck.
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t adesc0 = operand_desc<A_MN>(smem_base, 0);
      const uint64_t bdesc0 = operand_desc<B_MN>(smem_base + A_BYTES, 0);
      constexpr uint64_t KSTEP_A = (A_MN ? 2048 : 32) >> 4, KSTEP_B = (B_MN ? 2048 : 32) >> 4;
      constexpr uint64_t STAGE_STEP = STAGE_BYTES >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int r = t % tiles_per_g;
        const int ks = r / (mbc * nb);
        const int kb0 = ks * kb_per;
        const int kb1 = (gs.debug & 4) ? kb0 : min(kb0 + kb_per, kb_total);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t ad = adesc0 + stage * STAGE_STEP, bd = bdesc0 + stage * STAGE_STEP;
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              if constexpr (CM == 1)
                umma_f16(d_tmem, ad + k * KSTEP_A, bd + k * KSTEP_B, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
              else
                umma_f16_2sm(d_tmem, ad + k * KSTEP_A, bd + k * KSTEP_B, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (of both CTAs in the pair flavour) when these MMAs retire
            if constexpr (CM == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_2sm(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) {
          if constexpr (CM == 1) umma_commit(&tfull_bar[as]);
          else umma_commit_2sm(&tfull_bar[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps
    const int quad = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = warp >> 2;           // which part of the tile's columns it drains
    constexpr int NH = BN / PARTS;
    const uint32_t stage_buf = smem_u32(epi_stage + warp * GEMM_EPI_STAGE_BYTES);
    const uint32_t bias_buf = smem_u32(epi_bias + warp * GEMM_EPI_BIAS_BYTES);
    uint64_t* ebar = &epi_bar[warp];
    uint32_t ephase = 0;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const int g = t / tiles_per_g;
      int r = t - g * tiles_per_g;
      const int ks = r / (mbc * nb);
      r -= ks * (mbc * nb);
      const int m_blk = (r / nb) * CM + crank, n_blk = r - (r / nb) * nb;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * GEMM_BM + quad * 32;
      const int ncol0 = n_blk * BN + (Epi::kSplitColumns ? half * NH : 0);
      // global operands of the epilogue (bias slice, first residual chunk) are fetched BEFORE waiting for the
      // accumulator, so their latency hides behind the tile's MMAs
      typename Epi::Pre pre;
      Epi::template prologue<Epi::kSplitColumns ? NH : BN>(ep, pre, g, row0, lane, ncol0, gs.M, gs.N, stage_buf, ebar);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quad * 32) << 16);
      const int kb0 = ks * kb_per;
      // a split whose K range is empty contributes nothing (can happen when split_k does not divide)
      const bool has_k = kb0 < kb_total;
      if (gs.debug & 1) {
        // dev: accumulator handed back untouched
      } else if constexpr (Epi::kSplitColumns) {
        Epi::template run<NH>(ep, pre, taddr + half * NH, stage_buf, bias_buf, g, row0, lane, ncol0, gs.M, gs.N, has_k,
                              ebar, ephase);
      } else if (half == 0) {
        Epi::template run<BN>(ep, pre, taddr, stage_buf, bias_buf, g, row0, lane, ncol0, gs.M, gs.N, has_k, ebar, ephase);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CM == 1) mbar_arrive(&tempty_bar[as]);
        else mbar_arrive_cluster(&tempty_bar[as], 0);   // the leader's MMA thread owns the accumulator hand-off
      }
    }
    tma_store_wait_all();   // bulk-tensor stores issued by this thread (Epi with a TMA store path) are complete
  }

  tc_fence_before();
  __syncthreads();
  // no CTA may exit (or free TMEM) while its partner can still touch its shared memory, barriers or TMEM
  if constexpr (CM > 1) cluster_sync_all();
  if (warp == WARP_MMA) {
    tc_fence_after();
    if constexpr (CM == 1) tmem_dealloc(tmem_base, TMEM_COLS);
    else tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// =================================================================== epilogues
// activation math: MUFU-based (ex2 / rcp), no IEEE-division slow paths
// quick_gelu (transformers activations.py:122-123): x * sigmoid(1.702 x) = h + h * tanh(0.851 x), h = x / 2.
// ONE MUFU op per element (tanh.approx, max rel. error 2^-11 = half an fp16 ulp) instead of two (ex2 + rcp): the
// 16-lane MUFU pipe is what bounded the fc1 epilogue (54 -> 43 us at batch 16).
__device__ __forceinline__ float act_qgelu(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
__device__ __forceinline__ float act_qgelu_grad(float x) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-1.702f * x));
  return s + 1.702f * x * s * (1.0f - s);
}
__device__ __forceinline__ float act_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float act_gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// ACT_EXP_ROW:   v = exp(v - rowvec[g][m])                    (probabilities recomputed from the saved log-sum-exp)
// ACT_SMAX_GRAD: v = act_src[g][m][n] * (v - rowvec[g][m])      (softmax backward, rowvec = alpha * sum_j P dP)
enum : int { ACT_NONE = 0, ACT_QGELU = 1, ACT_GELU = 2, ACT_QGELU_GRAD = 3, ACT_GELU_GRAD = 4, ACT_EXP_ROW = 5,
             ACT_SMAX_GRAD = 6 };

template <int ACT>
__device__ __forceinline__ float apply_act(float v, float src) {
  if constexpr (ACT == ACT_QGELU) return act_qgelu(v);
  else if constexpr (ACT == ACT_GELU) return act_gelu(v);
  else if constexpr (ACT == ACT_QGELU_GRAD) return v * act_qgelu_grad(src);
  else if constexpr (ACT == ACT_GELU_GRAD) return v * act_gelu_grad(src);
  else if constexpr (ACT == ACT_SMAX_GRAD) return v * src;
  else return v;
}

// ---- per-warp transposition tile: 32 rows x 128 bytes, 16-byte chunks XOR-swizzled with (row & 7).
// "row phase": thread t owns row t (what tcgen05.ld 32x32b gives).  "line phase": lanes 8r..8r+7 own the
// eight 16-byte chunks of one row, four rows per access, so global accesses cover whole 128-byte lines.
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// bias slice of one epilogue warp: lane l holds columns ncol0 + 4l .. 4l+3 (zero outside [0, N) or the slice)
template <int W>
__device__ __forceinline__ float4 load_bias_slice(const float* bias, int ncol0, int N, int lane) {
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  const int c = ncol0 + 4 * lane;
  if (bias != nullptr && 4 * lane < W && c < N) {
    if (c + 4 <= N) {
      b = ldg_f4(bias + c);
    } else {
      b.x = __ldg(bias + c);
      if (c + 1 < N) b.y = __ldg(bias + c + 1);
      if (c + 2 < N) b.z = __ldg(bias + c + 2);
    }
  }
  return b;
}
// v[0..3] += bias columns 4 * src_lane .. + 3 (all lanes must call)
__device__ __forceinline__ void add_bias4(float* v, const float4& b, int src_lane) {
  v[0] += __shfl_sync(0xffffffffu, b.x, src_lane);
  v[1] += __shfl_sync(0xffffffffu, b.y, src_lane);
  v[2] += __shfl_sync(0xffffffffu, b.z, src_lane);
  v[3] += __shfl_sync(0xffffffffu, b.w, src_lane);
}

// line phase, fp16 tile of 32 rows x 64 columns: global -> staging (all loads issued before the first store, see
// EpiF32::load_addend)
__device__ __forceinline__ void tile_load_f16(const __half* src, long long ld, int row0, int n, int M, int N,
                                              uint32_t stage, int lane, bool vec_ok) {
  // N here is min(matrix N, end of this warp's column slice)
  const int sr = lane >> 3, sc = lane & 7;
  if (vec_ok && n + 64 <= N) {   // warp-uniform
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = min(row0 + i * 4 + sr, M - 1);
      v[i] = *reinterpret_cast<const uint4*>(src + (long long)m * ld + n + sc * 8);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sts128(stage_addr(stage, i * 4 + sr, sc), v[i]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + sr, m = row0 + row, col = n + sc * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m < M && col < N) {
      const __half* s = src + (long long)m * ld + col;
      if (vec_ok && col + 8 <= N) {
        v = *reinterpret_cast<const uint4*>(s);
      } else {
        __half t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = (col + j < N) ? s[j] : __float2half(0.f);
        v = *reinterpret_cast<uint4*>(t);
      }
    }
    sts128(stage_addr(stage, row, sc), v);
  }
}
// line phase, fp16 tile: staging -> global
__device__ __forceinline__ void tile_store_f16(__half* dst, long long ld, int row0, int n, int M, int N,
                                               uint32_t stage, int lane, bool vec_ok) {
  const int sr = lane >> 3, sc = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + sr, m = row0 + row, col = n + sc * 8;
    if (m < M && col < N) {
      uint4 v = lds128(stage_addr(stage, row, sc));
      __half* d = dst + (long long)m * ld + col;
      if (vec_ok && col + 8 <= N) {
        *reinterpret_cast<uint4*>(d) = v;
      } else {
        const __half* t = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (col + j < N) d[j] = t[j];
      }
    }
  }
}

// fp16 output:  v = alpha*acc + bias[n];  optionally save v (pre-activation) to `pre_out`;
//               then v = act(v)  or  v *= act'(dact_src[m][n]).
struct EpiF16Params {
  __half* out;
  __half* pre_out;
  const float* bias;
  const __half* dact_src;
  const float* rowvec;   // ACT_EXP_ROW / ACT_SMAX_GRAD: one fp32 per output row, [G][rowvec_stride]
  long long rowvec_stride;
  long long d_sb, d_sh;  // element offsets of dact_src per outer batch / head
  int ldo, ld_pre, ld_dact;
  long long o_sb, o_sh;  // element offsets per outer batch / head
  int H;
  int vec_ok;            // all of out / pre_out / dact_src rows are 16-byte aligned
  float alpha;
  const float* alpha_dev;  // optional device scalar multiplied into alpha
  int use_tma;             // out (and pre_out) are written with bulk-tensor stores through tm_out / tm_pre
  CUtensorMap tm_out;      // [M, N] fp16, box 64 columns x 32 rows, 128-byte swizzle (= the staging tile's layout)
  CUtensorMap tm_pre;
};

template <int ACT>
struct EpiF16 {
  using Params = EpiF16Params;
  static constexpr bool kSplitColumns = true;
  struct Pre { float4 bias; };
  template <int BN>
  static __device__ __forceinline__ void prologue(const Params& p, Pre& pre, int g, int row0, int lane, int n0, int M,
                                                  int N, uint32_t, uint64_t*) {
    (void)g; (void)row0; (void)M;
    pre.bias = load_bias_slice<BN>(p.bias, n0, N, lane);
  }
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, const Pre& pre, uint32_t taddr, uint32_t stage,
                                             uint32_t bias_smem, int g, int row0, int lane, int n0, int M, int Nfull,
                                             bool has_k, uint64_t*, uint32_t&) {
    const int N = min(Nfull, n0 + BN);  // this warp's column slice ends here
    const int outer = g / p.H, head = g - outer * p.H;
    __half* out = p.out + outer * p.o_sb + head * p.o_sh;
    const bool vec_ok = p.vec_ok != 0;
    const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
    constexpr bool kGrad = (ACT == ACT_QGELU_GRAD || ACT == ACT_GELU_GRAD || ACT == ACT_SMAX_GRAD);
    constexpr bool kRowVec = (ACT == ACT_EXP_ROW || ACT == ACT_SMAX_GRAD);
    if (row0 >= M) return;  // warp-uniform: nothing of this warp's 32 rows is inside the matrix
    // Fast path of the layer GEMMs: the warp's slice is whole 64-column chunks, the staging tile (32 rows x 128 B,
    // 16-byte chunks XOR-swizzled with the row) is exactly a 128B-swizzled TMA box, so one elected lane stores it with
    // a bulk-tensor copy (rows / columns outside the matrix are clipped by the TMA unit) instead of eight LDS + STG
    // with per-row address arithmetic and bounds checks per lane; the bias of the slice sits in shared memory
    // (written once per tile from the prologue's coalesced load) and is read back as broadcasts.
    const bool tma = (BN % 64 == 0) && p.use_tma != 0;
    const bool smem_bias = (BN % 64 == 0) && p.bias != nullptr;
    if (smem_bias) {
      if (lane < BN / 4)
        sts128(bias_smem + lane * 16, make_uint4(__float_as_uint(pre.bias.x), __float_as_uint(pre.bias.y),
                                                 __float_as_uint(pre.bias.z), __float_as_uint(pre.bias.w)));
      __syncwarp();
    }
    const __half* dsrc = kGrad ? p.dact_src + outer * p.d_sb + head * p.d_sh : nullptr;
    float rv = 0.f;         // row phase: thread t owns row t of the warp's 32
    if constexpr (kRowVec) rv = __ldg(p.rowvec + g * p.rowvec_stride + min(row0 + lane, M - 1));
#pragma unroll 1
    for (int c = 0; c < (BN + 63) / 64; ++c) {
      const int n = n0 + c * 64;
      if (n >= N) break;  // warp-uniform
      if constexpr (kGrad) {
        tile_load_f16(dsrc, p.ld_dact, row0, n, M, N, stage, lane, vec_ok);
        __syncwarp();
      }
      // pass 0 (only with pre_out): pre-activation values; pass 1: activated values
#pragma unroll 1
      for (int pass = (p.pre_out ? 0 : 1); pass < 2; ++pass) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (n + h * 32 >= N || c * 64 + h * 32 >= BN) break;
          uint32_t r[32];
          tmem_ld32(taddr + c * 64 + h * 32, r);
          tmem_ld_wait();
          // The epilogue is a serial chain on two warps per scheduler: every instruction here is exposed latency.
          // (has_k is always true for this epilogue: split-K needs the fp32 atomic one.)
          float v[32];
          if (alpha == 1.0f) {      // warp-uniform
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
          }
          if (p.bias) {
            if (smem_bias) {
              // all lanes read the same eight 16-byte words of the slice's bias from shared memory (broadcasts)
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const uint4 bq = lds128(bias_smem + (c * 16 + h * 8 + q) * 16);
                v[4 * q] += __uint_as_float(bq.x); v[4 * q + 1] += __uint_as_float(bq.y);
                v[4 * q + 2] += __uint_as_float(bq.z); v[4 * q + 3] += __uint_as_float(bq.w);
              }
            } else if (n + h * 32 + 32 <= N) {
              // all lanes read the same eight 16-byte words: one broadcast transaction each (no shuffles)
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + n + h * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 bq = __ldg(b4 + q);
                v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) add_bias4(v + 4 * q, pre.bias, c * 16 + h * 8 + q);
            }
          }
          if constexpr (kRowVec) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (ACT == ACT_EXP_ROW) ? __expf(v[i] - rv) : v[i] - rv;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t sa = stage_addr(stage, lane, h * 4 + q);
            if (pass == 1) {
              if constexpr (kGrad) {
                const uint4 s = lds128(sa);
                const __half2* sh = reinterpret_cast<const __half2*>(&s);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __half22float2(sh[j]);
                  v[8 * q + 2 * j] = apply_act<ACT>(v[8 * q + 2 * j], f.x);
                  v[8 * q + 2 * j + 1] = apply_act<ACT>(v[8 * q + 2 * j + 1], f.y);
                }
              } else if constexpr (ACT != ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[8 * q + j] = apply_act<ACT>(v[8 * q + j], 0.f);
              }
            }
            uint4 pk;
            pk.x = pack_h2(v[8 * q + 0], v[8 * q + 1]);
            pk.y = pack_h2(v[8 * q + 2], v[8 * q + 3]);
            pk.z = pack_h2(v[8 * q + 4], v[8 * q + 5]);
            pk.w = pack_h2(v[8 * q + 6], v[8 * q + 7]);
            sts128(sa, pk);
          }
        }
        if (tma) {
          // generic-proxy writes of the staging tile -> visible to the async proxy, then one lane stores the box
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(pass == 0 ? &p.tm_pre : &p.tm_out, stage, n, row0, 0);
            tma_store_commit();
            tma_store_wait_read();     // the tile may be overwritten by the next pass / chunk / output tile
          }
          __syncwarp();
        } else {
          __syncwarp();
          if (pass == 0) tile_store_f16(p.pre_out, p.ld_pre, row0, n, M, N, stage, lane, vec_ok);
          else tile_store_f16(out, p.ldo, row0, n, M, N, stage, lane, vec_ok);
          __syncwarp();
        }
      }
    }
  }
};

// fp32 output:  v = alpha*acc + bias[n] + resid[m][n] + pos[(m % rows_per_img + 1)][n]
//   mode 0: out = v      mode 1: out += v      mode 2: atomicAdd(out, v)  (split-K / shared outputs)
//   rows_per_img > 0 remaps output row m -> m + m / rows_per_img + 1 (patch rows -> token rows, CLS first).
struct EpiF32 {
  static constexpr bool kSplitColumns = true;
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
    int vec_ok;            // out / resid / pos rows are 16-byte aligned
    float alpha;
    const float* alpha_dev;
    int use_tma;             // mode 0, one [M, N] matrix, 16-byte aligned rows: TMA residual loads + TMA stores
    CUtensorMap tm_out;      // [M, N] fp32, box 32 columns x 32 rows, 128-byte swizzle (= the staging tile's layout)
    CUtensorMap tm_resid;
  };
  struct Pre { float4 bias; float4 add[8]; float4 b4; };
  // TMA path of the layer GEMMs (out-proj, fc2, fp32 dgrads): the 32 x 32 residual tile of a chunk is loaded by the TMA
  // unit straight into the warp's swizzled staging tile (the first chunk's while the tile's MMAs still run), every
  // thread adds its accumulator row + the bias (shared-memory broadcasts) IN PLACE - thread t owns row t, so there is
  // no transposition, no line phase, no per-row global address arithmetic - and one lane stores the tile with a
  // bulk-tensor copy (rows past M are zero-filled on the way in and clipped on the way out).
  template <int BN>
  static __device__ __forceinline__ bool tma_path(const Params& p, int n0, int N) {
    return p.use_tma != 0 && n0 + BN <= N;
  }
  // line phase gather of resid + pos + old output for the 32 x 32 tile at column n (registers only).
  // All loads of a chunk are issued back to back BEFORE anything consumes them: the SM issues in order, so a
  // consumer (or a data-dependent branch) between two loads would expose one full memory latency per load
  // (measured: 4 us per chunk, +23 us on the out-proj GEMM).  Rows past M are clamped to a valid row and zeroed
  // afterwards, so the common path has no branch at all.
  static __device__ __forceinline__ void load_addend(const Params& p, const float* out, int row0, int n, int M, int N,
                                                     int lane, float4* acc4) {
    const bool vec_ok = p.vec_ok != 0;
    const int sr = lane >> 3, sc = lane & 7;
    if (vec_ok && n + 32 <= N) {   // warp-uniform: whole 16-byte columns inside the matrix
      const int col = n + sc * 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.rows_per_img == 0 && !p.pos) {   // residual stream / accumulate: rows map one to one
        if (p.resid) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            acc4[i] = *reinterpret_cast<const float4*>(p.resid + (long long)min(row0 + i * 4 + sr, M - 1) * p.ldr + col);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc4[i] = z;
        }
        if (p.mode == 1) {
          float4 r1[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            r1[i] = *reinterpret_cast<const float4*>(out + (long long)min(row0 + i * 4 + sr, M - 1) * p.ldo + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc4[i].x += r1[i].x; acc4[i].y += r1[i].y; acc4[i].z += r1[i].z; acc4[i].w += r1[i].w; }
        }
        return;   // rows >= M hold a copy of row M - 1: harmless, the store phase never writes them
      }
      // patch-embedding flavour: output / residual rows are remapped (CLS first), position embedding added
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = min(row0 + i * 4 + sr, M - 1);
        int mo = m, prow = 0;
        if (p.rows_per_img > 0) {
          const int img = m / p.rows_per_img;
          prow = m - img * p.rows_per_img + 1;
          mo = m + img + 1;
        }
        float4 a = z;
        if (p.pos) a = ldg_f4(p.pos + (long long)prow * p.ldo + col);
        if (p.resid) { const float4 t = *reinterpret_cast<const float4*>(p.resid + (long long)mo * p.ldr + col); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
        if (p.mode == 1) { const float4 t = *reinterpret_cast<const float4*>(out + (long long)mo * p.ldo + col); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
        acc4[i] = a;
      }
      return;   // rows >= M hold a copy of row M - 1: harmless, the store phase never writes them
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 4 + sr, m = row0 + row, col = n + sc * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && col < N) {
        int mo = m, prow = 0;
        if (p.rows_per_img > 0) {
          const int img = m / p.rows_per_img;
          prow = m - img * p.rows_per_img + 1;
          mo = m + img + 1;
        }
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j) {
          if (col + j < N) {
            if (p.resid) e[j] += p.resid[(long long)mo * p.ldr + col + j];
            if (p.pos) e[j] += __ldg(p.pos + (long long)prow * p.ldo + col + j);
            if (p.mode == 1) e[j] += out[(long long)mo * p.ldo + col + j];
          }
        }
        a = make_float4(e[0], e[1], e[2], e[3]);
      }
      acc4[i] = a;
    }
  }
  static __device__ __forceinline__ bool has_addend(const Params& p) {
    return p.resid != nullptr || p.pos != nullptr || p.mode == 1;
  }
  // Lean path (the layer GEMMs): aligned rows, plain row mapping, store or accumulate, every column chunk of the
  // warp's slice fully inside the matrix.  Bias and addend are then applied in the LINE phase, where a lane owns
  // four fixed columns of eight rows: the bias is one float4 per chunk (no shuffles), the addend never goes through
  // shared memory, and a store costs a handful of instructions.  The generic path below spends ~85 instructions per
  // 16-byte store on its bounds / mode / row-remap handling, which with two epilogue warps per scheduler is pure
  // exposed latency (measured: 5.3 us per 128 x 192 tile without any global traffic).
  template <int BN>
  static __device__ __forceinline__ bool lean(const Params& p, int n0, int N) {
    return p.vec_ok != 0 && p.rows_per_img == 0 && p.pos == nullptr && n0 + BN <= N &&
           (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  }
  // line-phase operands of the 32 x 32 chunk at column n: addend rows (clamped to M - 1) and the bias float4
  static __device__ __forceinline__ void lean_fetch(const Params& p, const float* out, int row0, int n, int M, int lane,
                                                    float4* add, float4& b4) {
    const int sr = lane >> 3, col = n + (lane & 7) * 4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.resid) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        add[i] = *reinterpret_cast<const float4*>(p.resid + (long long)min(row0 + i * 4 + sr, M - 1) * p.ldr + col);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) add[i] = z;
    }
    if (p.mode == 1) {
      float4 r1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        r1[i] = *reinterpret_cast<const float4*>(out + (long long)min(row0 + i * 4 + sr, M - 1) * p.ldo + col);
#pragma unroll
      for (int i = 0; i < 8; ++i) { add[i].x += r1[i].x; add[i].y += r1[i].y; add[i].z += r1[i].z; add[i].w += r1[i].w; }
    }
    b4 = p.bias ? ldg_f4(p.bias + col) : z;
  }
  template <int BN>
  static __device__ __forceinline__ void prologue(const Params& p, Pre& pre, int g, int row0, int lane, int n0, int M,
                                                  int N, uint32_t stage, uint64_t* ebar) {
    if (tma_path<BN>(p, n0, N)) {
      pre.bias = load_bias_slice<BN>(p.bias, n0, N, lane);
      if (row0 < M && p.resid != nullptr && lane == 0) {
        tma_store_wait_read();                 // the previous tile's last store has finished reading the staging tile
        mbar_arrive_expect_tx(ebar, 4096);
        tma_load_3d_s(stage, &p.tm_resid, ebar, n0, row0, 0);
      }
      return;
    }
    if (lean<BN>(p, n0, N)) {
      if (row0 < M) {
        const int outer = g / p.H, head = g - outer * p.H;
        lean_fetch(p, p.out + outer * p.o_sb + head * p.o_sh, row0, n0, M, lane, pre.add, pre.b4);
      }
      return;
    }
    pre.bias = load_bias_slice<BN>(p.bias, n0, N, lane);
    if (has_addend(p) && row0 < M && n0 < N) {
      const int outer = g / p.H, head = g - outer * p.H;
      load_addend(p, p.out + outer * p.o_sb + head * p.o_sh, row0, n0, M, N, lane, pre.add);
    }
  }
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, Pre& pre, uint32_t taddr, uint32_t stage,
                                             uint32_t bias_smem, int g, int row0, int lane, int n0, int M, int N,
                                             bool has_k, uint64_t* ebar, uint32_t& ephase) {
    const int outer = g / p.H, head = g - outer * p.H;
    float* out = p.out + outer * p.o_sb + head * p.o_sh;
    const bool vec_ok = p.vec_ok != 0;
    const float alpha = p.alpha_dev ? p.alpha * __ldg(p.alpha_dev) : p.alpha;
    if (row0 >= M) return;
    if (!has_k && p.mode != 0) return;  // an empty K slice adds nothing
    const bool has_add = has_addend(p);
    const int sr = lane >> 3, sc = lane & 7;
    if (tma_path<BN>(p, n0, N)) {
      const bool with_resid = p.resid != nullptr;
      if (p.bias) {     // the slice's bias -> shared memory, read back as broadcasts
        if (lane < BN / 4)
          sts128(bias_smem + lane * 16, make_uint4(__float_as_uint(pre.bias.x), __float_as_uint(pre.bias.y),
                                                   __float_as_uint(pre.bias.z), __float_as_uint(pre.bias.w)));
        __syncwarp();
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (c > 0 && with_resid && lane == 0) {     // chunk 0's residual tile was requested in the prologue
          tma_store_wait_read();
          mbar_arrive_expect_tx(ebar, 4096);
          tma_load_3d_s(stage, &p.tm_resid, ebar, n0 + c * 32, row0, 0);
        }
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        if (with_resid) {
          mbar_wait(ebar, ephase);
          ephase ^= 1;
        } else {
          if (lane == 0) tma_store_wait_read();     // the staging tile is free again
          __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t sa = stage_addr(stage, lane, q);
          // same summation order as the line-phase path: (alpha * acc + bias) + residual
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_k)
            o = make_float4(__uint_as_float(r[4 * q]) * alpha, __uint_as_float(r[4 * q + 1]) * alpha,
                            __uint_as_float(r[4 * q + 2]) * alpha, __uint_as_float(r[4 * q + 3]) * alpha);
          if (p.bias) {
            const uint4 bq = lds128(bias_smem + (c * 8 + q) * 16);
            o.x += __uint_as_float(bq.x); o.y += __uint_as_float(bq.y); o.z += __uint_as_float(bq.z); o.w += __uint_as_float(bq.w);
          }
          if (with_resid) {
            const uint4 u = lds128(sa);
            o.x += __uint_as_float(u.x); o.y += __uint_as_float(u.y); o.z += __uint_as_float(u.z); o.w += __uint_as_float(u.w);
          }
          sts128(sa, make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&p.tm_out, stage, n0 + c * 32, row0, 0);
          tma_store_commit();
        }
      }
      return;
    }
    if (lean<BN>(p, n0, N)) {
      float* d0 = out + (long long)(row0 + sr) * p.ldo + n0 + sc * 4;
      const long long step = 4LL * p.ldo;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        // line-phase operands of this chunk: chunk 0's were fetched in the prologue (before the accumulator wait);
        // later chunks fetch theirs here, in flight behind the TMEM load and the row phase.  (With four epilogue
        // warps per scheduler the other warps cover this latency; a second register set for the next chunk's addend
        // would push the 576-thread flavours over their register budget.)
        if (c > 0) lean_fetch(p, out, row0, n0 + c * 32, M, lane, pre.add, pre.b4);
        // row phase: accumulator -> staging (thread t owns row t)
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float a0 = has_k ? __uint_as_float(r[4 * q]) * alpha : 0.f, a1 = has_k ? __uint_as_float(r[4 * q + 1]) * alpha : 0.f;
          const float a2 = has_k ? __uint_as_float(r[4 * q + 2]) * alpha : 0.f, a3 = has_k ? __uint_as_float(r[4 * q + 3]) * alpha : 0.f;
          sts128(stage_addr(stage, lane, q), make_uint4(__float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3)));
        }
        __syncwarp();
        // line phase: + bias + addend, whole 128-byte lines out
        float* d = d0 + c * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 u = lds128(stage_addr(stage, i * 4 + sr, sc));
          float4 o;
          o.x = __uint_as_float(u.x) + pre.b4.x + pre.add[i].x;
          o.y = __uint_as_float(u.y) + pre.b4.y + pre.add[i].y;
          o.z = __uint_as_float(u.z) + pre.b4.z + pre.add[i].z;
          o.w = __uint_as_float(u.w) + pre.b4.w + pre.add[i].w;
          if (row0 + i * 4 + sr < M) {
            if (p.mode == 2) atomicAdd(reinterpret_cast<float4*>(d + i * step), o);   // split-K / shared outputs: one 16-byte red
            else *reinterpret_cast<float4*>(d + i * step) = o;
          }
        }
        __syncwarp();
      }
      return;
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n = n0 + c * 32;
      if (n >= N) break;
      if (has_add) {
        // the addend of this chunk was fetched one chunk (or one tile prologue) ago: park it in the staging tile
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sts128(stage_addr(stage, i * 4 + sr, sc), make_uint4(__float_as_uint(pre.add[i].x), __float_as_uint(pre.add[i].y),
                                                                __float_as_uint(pre.add[i].z), __float_as_uint(pre.add[i].w)));
        __syncwarp();
        // ... and start fetching the next chunk's addend now, so its latency overlaps this chunk's row phase
        if (c + 1 < BN / 32 && n + 32 < N) load_addend(p, out, row0, n + 32, M, N, lane, pre.add);
      }
      // row phase
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = has_k ? __uint_as_float(r[i]) * alpha : 0.0f;
      if (p.bias) {
#pragma unroll
        for (int q = 0; q < 8; ++q) add_bias4(v + 4 * q, pre.bias, c * 8 + q);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t sa = stage_addr(stage, lane, q);
        if (has_add) {
          const uint4 a = lds128(sa);
          v[4 * q] += __uint_as_float(a.x); v[4 * q + 1] += __uint_as_float(a.y);
          v[4 * q + 2] += __uint_as_float(a.z); v[4 * q + 3] += __uint_as_float(a.w);
        }
        sts128(sa, make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                              __float_as_uint(v[4 * q + 3])));
      }
      __syncwarp();
      // line phase: write out
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + sr, m = row0 + row, col = n + sc * 4;
        if (m < M && col < N) {
          int mo = m;
          if (p.rows_per_img > 0) mo = m + m / p.rows_per_img + 1;
          const uint4 u = lds128(stage_addr(stage, row, sc));
          float* d = out + (long long)mo * p.ldo + col;
          const float t[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
          if (p.mode == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (col + j < N) atomicAdd(d + j, t[j]);
          } else if (vec_ok && col + 4 <= N) {
            *reinterpret_cast<float4*>(d) = make_float4(t[0], t[1], t[2], t[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (col + j < N) d[j] = t[j];
          }
        }
      }
      __syncwarp();
    }
  }
};

// Class-head tail (reference src/models.py:35-36): sims[m][c] = max_{j<3} acc[m][3c + j];
// also records which prompt variant won (for the backward scatter).
struct EpiPool3 {
  static constexpr bool kSplitColumns = false;  // groups of 3 columns do not split at BN / 2
  struct Params {
    float* sims;        // [M, C]
    uint8_t* argmax;    // [M, C]
    int C;
  };
  struct Pre {};
  template <int BN>
  static __device__ __forceinline__ void prologue(const Params&, Pre&, int, int, int, int, int, int, uint32_t, uint64_t*) {}
  template <int BN>
  static __device__ __forceinline__ void run(const Params& p, const Pre&, uint32_t taddr, uint32_t stage,
                                             uint32_t bias_smem, int g, int row0, int lane, int n0, int M, int N,
                                             bool has_k, uint64_t*, uint32_t&) {
    (void)g; (void)has_k; (void)stage; (void)bias_smem;
    const int m = row0 + lane;
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
