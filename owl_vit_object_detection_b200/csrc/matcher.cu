// Bipartite matcher + push-pull loss on the device (no host round trip on the step path).
//
//   owl_matcher_cost   reference src/matcher.py:103-131   cost matrix, one HBM pass, coalesced, written
//                                                          target-major [B, Tmax, P] for the solver
//   owl_lsap           reference src/matcher.py:134-137   SciPy rectangular LSAP (Crouse 2016) restated
//                                                          for one warp per image, fp64 duals, same tie rule
//   owl_match_loss     reference src/matcher.py:138-159 + src/losses.py:42-69,100-108,16-40
//                                                          target_classes, IoU>0.85 ordered label sweep,
//                                                          box + class losses AND their gradients
//   owl_loss_backward  autograd of the above               scales / scatters the saved gradients
//
// Every fp32 quantity that feeds a discrete decision (assignment, IoU > 0.85) is computed with the
// reference's op order using non-contracted IEEE ops (__fadd_rn / __fmul_rn / __fdiv_rn), so it is
// bit-identical to the fp32 torch ops of the reference.
#include "common.h"
#include <algorithm>
#include <stdlib.h>
#include <math_constants.h>

namespace owl {

__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// exp of the softmax numerator (argument <= 0).  The class term cannot be bit-identical to the reference's CPU softmax
// anyway (Sleef vs CUDA expf, different summation order: SURVEY 8 a.1) and is ~1/80 of the cost, so the 2-ulp
// ex2.approx form (2 instructions instead of ~10; the kernel is instruction-issue bound, ncu: issue 85 % at T = 10) is
// used; bench.py checks the resulting assignments of 10 000 images per T against the reference arithmetic every run.
__device__ __forceinline__ float soft_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

struct Box { float x0, y0, x1, y1; };

// status bits (host raises what the reference would have raised inline): 1 degenerate box (src/matcher.py:34-35),
// 2 infeasible cost matrix, 4 num_targets outside [0, Tmax], 8 label outside [0, C) (an IndexError in the reference's
// gather, src/matcher.py:118).  Out-of-range values are clamped so that no kernel indexes out of bounds.
__device__ __forceinline__ int clamp_targets(int t, int Tmax) { return min(max(t, 0), Tmax); }

// reference src/matcher.py:8-21 (torchvision box_area + box_iou): returns iou, writes union
__device__ __forceinline__ float iou_union(const Box& a, const Box& b, float* uni) {
  const float area1 = fmul(fsub(a.x1, a.x0), fsub(a.y1, a.y0));
  const float area2 = fmul(fsub(b.x1, b.x0), fsub(b.y1, b.y0));
  const float w = fmaxf(fsub(fminf(a.x1, b.x1), fmaxf(a.x0, b.x0)), 0.0f);
  const float h = fmaxf(fsub(fminf(a.y1, b.y1), fmaxf(a.y0, b.y0)), 0.0f);
  const float inter = fmul(w, h);
  const float u = fsub(fadd(area1, area2), inter);
  *uni = u;
  return fdiv(inter, u);
}
// reference src/matcher.py:25-44
__device__ __forceinline__ float giou_pair(const Box& a, const Box& b) {
  float u;
  const float iou = iou_union(a, b, &u);
  const float cw = fmaxf(fsub(fmaxf(a.x1, b.x1), fminf(a.x0, b.x0)), 0.0f);
  const float ch = fmaxf(fsub(fmaxf(a.y1, b.y1), fminf(a.y0, b.y0)), 0.0f);
  const float area = fmul(cw, ch);
  return fsub(iou, fdiv(fsub(area, u), area));
}

// =====================================================================================================
// K11: cost matrix.  CTA = 32 predictions x all targets of one image.
//   phase 1: warp per prediction row: softmax over C classes (row read coalesced, shuffle reductions)
//   phase 2: lane = prediction, warp strides over targets: C = (L1 + (-p[label])) + (-GIoU); 128-byte
//            coalesced stores into costT[b][t][p0 .. p0+31].
// =====================================================================================================
constexpr int COST_ROWS = 64;      // predictions per CTA (two passes of 32 rows x 8 lanes through the softmax)
constexpr int COST_THREADS = 256;
#ifndef OWL_COST_MINB
#define OWL_COST_MINB 5
#endif

// The kernel is instruction-issue bound (ncu: issue slots 73 % busy, DRAM 21 %: the index-exact arithmetic costs
// ~4x a fast-math version), so the row length is a template parameter (no dead iterations / predicates) and a CTA
// covers 64 predictions (the per-CTA target set-up is amortised, and 2 x T (row group, target) items fill the
// eight warps of the pair phase better than T items do).
template <int NVEC /* float4 per lane: ceil(C / 32) */, bool VEC /* C % 4 == 0 */>
__global__ void __launch_bounds__(COST_THREADS, OWL_COST_MINB)
matcher_cost_kernel(const float* __restrict__ sims, const float* __restrict__ boxes,
                    const long long* __restrict__ labels, const float* __restrict__ tboxes,
                    const int* __restrict__ num_targets, float* __restrict__ costT, int P, int C, int Tmax,
                    int* __restrict__ status, float w_class, float w_bbox, float w_giou) {
  pdl_grid_wait();
  extern __shared__ float sm[];
  float* prob = sm;                                   // [64][C + 1]
  float* pbox = prob + COST_ROWS * (C + 1);           // [64][4]
  float* tbox = pbox + COST_ROWS * 4;                 // [Tmax][4]
  int* tlab = reinterpret_cast<int*>(tbox + Tmax * 4);  // [Tmax]

  const int b = blockIdx.y;
  const int p0 = blockIdx.x * COST_ROWS;
  const int T = clamp_targets(num_targets[b], Tmax);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && blockIdx.x == 0 && T != num_targets[b]) atomicOr(status, 4);

  for (int i = threadIdx.x; i < T * 4; i += COST_THREADS) tbox[i] = tboxes[(1LL * b * Tmax) * 4 + i];
  for (int i = threadIdx.x; i < T; i += COST_THREADS) {
    const long long l = labels[1LL * b * Tmax + i];
    const bool ok = l >= 0 && l < C;
    if (!ok && blockIdx.x == 0) atomicOr(status, 8);
    tlab[i] = ok ? static_cast<int>(l) : 0;
  }
  {
    const int i = threadIdx.x;                       // 64 rows x 4 coordinates = 256 threads
    const int p = p0 + (i >> 2);
    pbox[i] = p < P ? boxes[(1LL * b * P + p) * 4 + (i & 3)] : 0.0f;
  }
  // phase 1: softmax(sims[b, p, :])  (reference src/matcher.py:106-108).  Eight lanes per row, four rows per warp,
  // 32 rows per pass: every lane issues all of its 128-bit loads (of BOTH passes) before the first reduction.
  {
    const int sub = lane & 7;
    float4 v[2][NVEC];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 32 + warp * 4 + (lane >> 3);
      const int p = p0 + r;
      const bool row_ok = p < P;
      const float* row = sims + (1LL * b * P + (row_ok ? p : 0)) * C;
#pragma unroll
      for (int k = 0; k < NVEC; ++k) {
        const int c = k * 32 + sub * 4;
        float4 t = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        if (row_ok && c < C) {
          if (VEC) {
            t = __ldg(reinterpret_cast<const float4*>(row + c));
          } else {
            t.x = row[c];
            if (c + 1 < C) t.y = row[c + 1];
            if (c + 2 < C) t.z = row[c + 2];
            if (c + 3 < C) t.w = row[c + 3];
          }
        }
        v[pass][k] = t;
      }
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 32 + warp * 4 + (lane >> 3);
      float m = -CUDART_INF_F;
#pragma unroll
      for (int k = 0; k < NVEC; ++k)
        m = fmaxf(m, fmaxf(fmaxf(v[pass][k].x, v[pass][k].y), fmaxf(v[pass][k].z, v[pass][k].w)));
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NVEC; ++k) {
        const int c = k * 32 + sub * 4;
        float4& x = v[pass][k];
        x.x = c < C ? soft_exp(fsub(x.x, m)) : 0.f;
        x.y = c + 1 < C ? soft_exp(fsub(x.y, m)) : 0.f;
        x.z = c + 2 < C ? soft_exp(fsub(x.z, m)) : 0.f;
        x.w = c + 3 < C ? soft_exp(fsub(x.w, m)) : 0.f;
        s += (x.x + x.y) + (x.z + x.w);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float inv = fdiv(1.0f, s);
#pragma unroll
      for (int k = 0; k < NVEC; ++k) {
        const int c = k * 32 + sub * 4;
        float* dst = prob + r * (C + 1) + c;
        const float4 x = v[pass][k];
        if (c < C) dst[0] = fmul(x.x, inv);
        if (c + 1 < C) dst[1] = fmul(x.y, inv);
        if (c + 2 < C) dst[2] = fmul(x.z, inv);
        if (c + 3 < C) dst[3] = fmul(x.w, inv);
      }
    }
  }
  __syncthreads();
  // degenerate boxes: the reference asserts (src/matcher.py:34-35); we flag and let the host raise
  if (threadIdx.x < COST_ROWS && p0 + threadIdx.x < P) {
    const float* q = pbox + threadIdx.x * 4;
    if (!(q[2] >= q[0]) || !(q[3] >= q[1])) atomicOr(status, 1);
  }
  for (int i = threadIdx.x; i < T; i += COST_THREADS)
    if (!(tbox[i * 4 + 2] >= tbox[i * 4 + 0]) || !(tbox[i * 4 + 3] >= tbox[i * 4 + 1])) atomicOr(status, 1);

  // phase 2: items = (row group of 32 predictions, target); lane = prediction inside the group.  A warp keeps ONE row
  // group (the item stride, 8 warps, is even), so its prediction box, that box's area and its probability row are
  // loop invariants held in registers; only the target changes per item.
  {
    const int grp = warp & 1;
    const int row = grp * 32 + lane, p = p0 + row;
    if (p < P) {
      const Box pb = {pbox[row * 4 + 0], pbox[row * 4 + 1], pbox[row * 4 + 2], pbox[row * 4 + 3]};
      const float area1 = fmul(fsub(pb.x1, pb.x0), fsub(pb.y1, pb.y0));          // torchvision box_area(boxes1)
      const float* prow = prob + row * (C + 1);
      float* dst = costT + (1LL * b * Tmax) * P + p;
      for (int t = warp >> 1; t < T; t += COST_THREADS / 64) {
        const Box tb = {tbox[t * 4 + 0], tbox[t * 4 + 1], tbox[t * 4 + 2], tbox[t * 4 + 3]};
        // torch.cdist(p=1): ((|dx0| + |dy0|) + |dx1|) + |dy1|   (reference src/matcher.py:121)
        const float l1 = fadd(fadd(fadd(fabsf(fsub(pb.x0, tb.x0)), fabsf(fsub(pb.y0, tb.y0))),
                                   fabsf(fsub(pb.x1, tb.x1))), fabsf(fsub(pb.y1, tb.y1)));
        const float pr = prow[tlab[t]];
        // reference src/matcher.py:8-21,25-44 (same op order as iou_union / giou_pair above, area1 hoisted)
        const float area2 = fmul(fsub(tb.x1, tb.x0), fsub(tb.y1, tb.y0));
        const float iw = fmaxf(fsub(fminf(pb.x1, tb.x1), fmaxf(pb.x0, tb.x0)), 0.0f);
        const float ih = fmaxf(fsub(fminf(pb.y1, tb.y1), fmaxf(pb.y0, tb.y0)), 0.0f);
        const float inter = fmul(iw, ih);
        const float uni = fsub(fadd(area1, area2), inter);
        const float iou = fdiv(inter, uni);
        const float cw = fmaxf(fsub(fmaxf(pb.x1, tb.x1), fminf(pb.x0, tb.x0)), 0.0f);
        const float ch = fmaxf(fsub(fmaxf(pb.y1, tb.y1), fminf(pb.y0, tb.y0)), 0.0f);
        const float hull = fmul(cw, ch);
        const float gi = fsub(iou, fdiv(fsub(hull, uni), hull));
        // (w_bbox * cost_bbox + w_class * cost_class) + w_giou * cost_giou with cost_class = -p, cost_giou = -giou
        // (src/matcher.py:127-131; every product and sum rounds separately, like the torch ops).  With the reference's
        // weights (all 1, src/matcher.py:58-60) the products are exact and this is (l1 - p) - giou bit for bit.
        dst[1LL * t * P] = fadd(fadd(fmul(w_bbox, l1), fmul(w_class, -pr)), fmul(w_giou, -gi));
      }
    }
  }
}

// =====================================================================================================
// K12: rectangular LSAP, one warp per image.  Solves the transposed problem (rows = targets,
// columns = predictions) exactly like SciPy does when rows > cols.  See oracle/lsap.c for the scalar
// restatement this mirrors step by step.
// =====================================================================================================
struct Cand { double v; int un; int it; };

__device__ __forceinline__ Cand cand_combine(const Cand& a, const Cand& b) {
  if (a.it < 0) return b;
  if (b.it < 0) return a;
  if (a.v < b.v) return a;
  if (b.v < a.v) return b;
  if (a.un != b.un) return a.un ? a : b;      // ties go to a still-unassigned column ...
  if (a.un) return a.it > b.it ? a : b;        // ... the last such in scan order,
  return a.it < b.it ? a : b;                  // else the first column that reached the minimum
}

constexpr int LSAP_WARPS = 4;

// bytes of the solver's own state in the CTA-per-image kernel (the staged cost rows follow, 16-byte aligned)
__host__ __device__ inline size_t lsap_block_base_bytes(int P, int Tmax) {
  const int Tpad = (Tmax + 4) & ~3;
  const size_t b = sizeof(double) * (2 * (size_t)P + Tpad) + sizeof(short) * (3 * (size_t)P + 3 * Tpad + 8);
  return (b + 15) & ~size_t(15);
}

__global__ void __launch_bounds__(LSAP_WARPS * 32)
lsap_kernel(const float* __restrict__ costT, const int* __restrict__ num_targets, int B, int P, int Tmax,
            int* __restrict__ match_pred /*[B,Tmax]*/, int* __restrict__ status) {
  pdl_grid_wait();
  extern __shared__ __align__(16) unsigned char lsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * LSAP_WARPS + warp;
  if (b >= B) return;
  const int Tpad = (Tmax + 4) & ~3;
  // per-warp carve-up
  const size_t per_warp = sizeof(double) * (2 * (size_t)P + Tpad) + sizeof(short) * (3 * (size_t)P + 3 * Tpad + 8);
  unsigned char* base = lsm + ((per_warp + 15) & ~size_t(15)) * warp;
  double* spc = reinterpret_cast<double*>(base);
  double* v = spc + P;
  double* u = v + P;
  short* path = reinterpret_cast<short*>(u + Tpad);
  short* row4col = path + P;
  short* remaining = row4col + P;
  short* col4row = remaining + P;
  short* sr_list = col4row + Tpad;
  short* sc_list = sr_list + Tpad + 4;

  const int nr = clamp_targets(num_targets[b], Tmax), nc = P;
  const float* cost = costT + 1LL * b * Tmax * P;

  for (int j = lane; j < nc; j += 32) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
  __syncwarp();

  for (int cur = 0; cur < nr; ++cur) {
    for (int j = lane; j < nc; j += 32) { spc[j] = CUDART_INF; remaining[j] = static_cast<short>(nc - 1 - j); }
    int num_remaining = nc, n_sr = 0, n_sc = 0;
    double min_val = 0.0;
    int i = cur, sink = -1;
    __syncwarp();
    while (sink < 0) {
      if (lane == 0) sr_list[n_sr] = static_cast<short>(i);
      ++n_sr;
      const double ui = u[i];
      const float* crow = cost + 1LL * i * nc;
      Cand best = {0.0, 0, -1};
      // The scan is bound by the latency of the cost loads (one dependent global load per column and lane, ~16 warps
      // per SM): a lane first collects the columns of EIGHT of its scan positions, issues their eight cost loads back
      // to back, and only then relaxes them - in the same increasing scan order as before.
      constexpr int KB = 9;     // scan positions per lane and pass (576 columns = two passes)
      for (int base = lane; base < num_remaining; base += 32 * KB) {
        int jj[KB];
        float cc[KB];
        double vv[KB], ss[KB];
        short rc[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const int it = base + 32 * k;
          jj[k] = it < num_remaining ? remaining[it] : -1;
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) cc[k] = jj[k] >= 0 ? __ldg(crow + jj[k]) : 0.0f;
#pragma unroll
        for (int k = 0; k < KB; ++k) {       // the columns of a pass are distinct: their state can be read up front
          const int j = max(jj[k], 0);
          vv[k] = v[j];
          ss[k] = spc[j];
          rc[k] = row4col[j];
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const int j = jj[k];
          if (j >= 0) {
            const double r = ((min_val + static_cast<double>(cc[k])) - ui) - vv[k];
            double sv = ss[k];
            if (r < sv) { path[j] = static_cast<short>(i); spc[j] = r; sv = r; }
            const Cand c = {sv, rc[k] == -1 ? 1 : 0, base + 32 * k};
            best = cand_combine(best, c);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Cand other;
        other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        other.un = __shfl_xor_sync(0xffffffffu, best.un, o);
        other.it = __shfl_xor_sync(0xffffffffu, best.it, o);
        best = cand_combine(best, other);
      }
      min_val = best.v;
      if (best.it < 0 || min_val == CUDART_INF) {  // infeasible (only with inf / nan costs)
        if (lane == 0) atomicOr(status, 2);
        sink = -2;
        break;
      }
      const int j = remaining[best.it];
      const int r4c = row4col[j];
      if (r4c == -1) sink = j; else i = r4c;
      __syncwarp();
      if (lane == 0) {
        sc_list[n_sc] = static_cast<short>(j);
        remaining[best.it] = remaining[num_remaining - 1];
      }
      ++n_sc;
      --num_remaining;
      __syncwarp();
    }
    if (sink < 0) break;
    // dual update (before augmenting, like the reference solver)
    if (lane == 0) u[cur] += min_val;
    for (int k = lane; k < n_sr; k += 32) {
      const int r = sr_list[k];
      if (r != cur) u[r] += min_val - spc[col4row[r]];
    }
    for (int k = lane; k < n_sc; k += 32) {
      const int j = sc_list[k];
      v[j] -= min_val - spc[j];
    }
    __syncwarp();
    if (lane == 0) {
      int j = sink;
      for (;;) {
        const int r = path[j];
        row4col[j] = static_cast<short>(r);
        const int t = col4row[r];
        col4row[r] = static_cast<short>(j);
        j = t;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }
  for (int t = lane; t < Tmax; t += 32) match_pred[1LL * b * Tmax + t] = t < nr ? col4row[t] : -1;
}

// Same solver, one CTA per image: used when the batch is too small to fill the GPU with one warp per image
// (a training step has 16 images; the slowest image sets the latency).  Every scan over the remaining columns is
// spread over NT threads; candidates are combined with the same associative tie rule.
template <int NT>
__global__ void __launch_bounds__(NT)
lsap_block_kernel(const float* __restrict__ costT, const int* __restrict__ num_targets, int P, int Tmax,
                  int* __restrict__ match_pred /*[B,Tmax]*/, int* __restrict__ status, int smem_rows) {
  pdl_grid_wait();
  extern __shared__ __align__(16) unsigned char lsm[];
  constexpr int NW = NT / 32;
  __shared__ Cand wbest[NW];
  __shared__ int s_sink, s_i;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int Tpad = (Tmax + 4) & ~3;
  double* spc = reinterpret_cast<double*>(lsm);
  double* v = spc + P;
  double* u = v + P;
  short* path = reinterpret_cast<short*>(u + Tpad);
  short* row4col = path + P;
  short* remaining = row4col + P;
  short* col4row = remaining + P;
  short* sr_list = col4row + Tpad;
  short* sc_list = sr_list + Tpad + 4;
  // the image's cost rows, staged once: every scan below would otherwise start with an exposed global-load
  // latency (~2 scans per target, 0.6 us each).  Rows that do not fit (smem_rows) are read from global memory.
  float* scost = reinterpret_cast<float*>(lsm + lsap_block_base_bytes(P, Tmax));

  const int nr = clamp_targets(num_targets[b], Tmax), nc = P;
  const float* cost = costT + 1LL * b * Tmax * P;
  const int rows_s = min(nr, smem_rows);
  if ((P & 3) == 0) {
    const float4* src = reinterpret_cast<const float4*>(cost);
    float4* dst = reinterpret_cast<float4*>(scost);
    for (int k = tid; k < rows_s * (P >> 2); k += NT) dst[k] = __ldg(src + k);
  } else {
    for (int k = tid; k < rows_s * P; k += NT) scost[k] = __ldg(cost + k);
  }
  for (int j = tid; j < nc; j += NT) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  for (int i = tid; i < nr; i += NT) { u[i] = 0.0; col4row[i] = -1; }
  __syncthreads();

  for (int cur = 0; cur < nr; ++cur) {
    for (int j = tid; j < nc; j += NT) { spc[j] = CUDART_INF; remaining[j] = static_cast<short>(nc - 1 - j); }
    int num_remaining = nc, n_sr = 0, n_sc = 0;
    double min_val = 0.0;
    int i = cur, sink = -1;
    __syncthreads();
    while (sink < 0) {
      if (tid == 0) sr_list[n_sr] = static_cast<short>(i);
      ++n_sr;
      const double ui = u[i];
      const bool in_smem = i < rows_s;   // block-uniform
      const float* crow = in_smem ? scost + i * nc : cost + 1LL * i * nc;
      Cand best = {0.0, 0, -1};
      for (int it = tid; it < num_remaining; it += NT) {
        const int j = remaining[it];
        const float cij = in_smem ? crow[j] : __ldg(crow + j);
        const double r = ((min_val + static_cast<double>(cij)) - ui) - v[j];
        double sj = spc[j];
        if (r < sj) { path[j] = static_cast<short>(i); spc[j] = r; sj = r; }
        const Cand c = {sj, row4col[j] == -1 ? 1 : 0, it};
        best = cand_combine(best, c);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Cand other;
        other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        other.un = __shfl_xor_sync(0xffffffffu, best.un, o);
        other.it = __shfl_xor_sync(0xffffffffu, best.it, o);
        best = cand_combine(best, other);
      }
      if (lane == 0) wbest[warp] = best;
      __syncthreads();
      best = wbest[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) best = cand_combine(best, wbest[w]);
      min_val = best.v;
      if (best.it < 0 || min_val == CUDART_INF) {  // infeasible (only with inf / nan costs); block-uniform
        if (tid == 0) atomicOr(status, 2);
        sink = -2;
        break;
      }
      const int j = remaining[best.it];
      const int r4c = row4col[j];
      if (r4c == -1) sink = j; else i = r4c;
      __syncthreads();   // everyone has read remaining[best.it] / wbest before they are overwritten
      if (tid == 0) {
        sc_list[n_sc] = static_cast<short>(j);
        remaining[best.it] = remaining[num_remaining - 1];
      }
      ++n_sc;
      --num_remaining;
      __syncthreads();
    }
    if (sink < 0) break;
    if (tid == 0) u[cur] += min_val;
    for (int k = tid; k < n_sr; k += NT) {
      const int r = sr_list[k];
      if (r != cur) u[r] += min_val - spc[col4row[r]];
    }
    for (int k = tid; k < n_sc; k += NT) {
      const int j = sc_list[k];
      v[j] -= min_val - spc[j];
    }
    __syncthreads();
    if (tid == 0) {
      int j = sink;
      for (;;) {
        const int r = path[j];
        row4col[j] = static_cast<short>(r);
        const int t = col4row[r];
        col4row[r] = static_cast<short>(j);
        j = t;
        if (r == cur) break;
      }
    }
    __syncthreads();
  }
  for (int t = tid; t < Tmax; t += NT) match_pred[1LL * b * Tmax + t] = t < nr ? col4row[t] : -1;
}

// =====================================================================================================
// K13 + K14: CTA per image.
// =====================================================================================================
constexpr int LOSS_THREADS = 256;
constexpr int LOSS_CHUNKS = 24;   // row slices per image in the class-loss kernel
constexpr int LOSS_WS = 64;       // floats of per-image workspace: [0..3] ce, bg, bbox, giou; [4] #positive rows;
                                  // [8 + 2 c + {0,1}] class-loss partial sums of row slice c
static_assert(8 + 2 * LOSS_CHUNKS <= LOSS_WS, "loss workspace");

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < LOSS_THREADS / 32; ++w) t += red[w];
  return t;
}

// sub-gradient weights torch uses for maximum / minimum (ties split evenly)
__device__ __forceinline__ float wmax(float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); }
__device__ __forceinline__ float wmin(float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); }

// d(1 - giou(s, t)) / ds, autograd of reference src/matcher.py:8-44 w.r.t. the first box
__device__ void giou_loss_grad(const Box& s, const Box& t, float* loss, float g[4]) {
  const float sw = s.x1 - s.x0, sh = s.y1 - s.y0;
  const float area1 = sw * sh, area2 = (t.x1 - t.x0) * (t.y1 - t.y0);
  const float ltx = fmaxf(s.x0, t.x0), lty = fmaxf(s.y0, t.y0), rbx = fminf(s.x1, t.x1), rby = fminf(s.y1, t.y1);
  const float dw = rbx - ltx, dh = rby - lty;
  const float w = fmaxf(dw, 0.f), h = fmaxf(dh, 0.f);
  const float inter = w * h;
  const float uni = area1 + area2 - inter;
  const float iou = inter / uni;
  const float cltx = fminf(s.x0, t.x0), clty = fminf(s.y0, t.y0), crbx = fmaxf(s.x1, t.x1), crby = fmaxf(s.y1, t.y1);
  const float dcw = crbx - cltx, dch = crby - clty;
  const float cw = fmaxf(dcw, 0.f), ch = fmaxf(dch, 0.f);
  const float area = cw * ch;
  const float gi = iou - (area - uni) / area;
  *loss = 1.0f - gi;
  // reverse mode, seed d(giou) = 1 (the caller negates)
  const float g_area = -uni / (area * area);
  const float g_uni = -inter / (uni * uni) + 1.0f / area;
  const float g_inter = 1.0f / uni - g_uni;
  const float g_area1 = g_uni;
  const float g_w = g_inter * h * (dw >= 0.f ? 1.f : 0.f);
  const float g_h = g_inter * w * (dh >= 0.f ? 1.f : 0.f);
  const float g_cw = g_area * ch * (dcw >= 0.f ? 1.f : 0.f);
  const float g_ch = g_area * cw * (dch >= 0.f ? 1.f : 0.f);
  float gx0 = -g_area1 * sh, gx1 = g_area1 * sh, gy0 = -g_area1 * sw, gy1 = g_area1 * sw;
  gx1 += g_w * wmin(s.x1, t.x1);   // rbx = min(x1, a1)
  gx0 -= g_w * wmax(s.x0, t.x0);   // ltx = max(x0, a0)
  gy1 += g_h * wmin(s.y1, t.y1);
  gy0 -= g_h * wmax(s.y0, t.y0);
  gx1 += g_cw * wmax(s.x1, t.x1);  // crbx = max(x1, a1)
  gx0 -= g_cw * wmin(s.x0, t.x0);  // cltx = min(x0, a0)
  gy1 += g_ch * wmax(s.y1, t.y1);
  gy0 -= g_ch * wmin(s.y0, t.y0);
  g[0] = -gx0; g[1] = -gy0; g[2] = -gx1; g[3] = -gy1;
}

__global__ void __launch_bounds__(LOSS_THREADS)
match_loss_kernel(const float* __restrict__ sims, const float* __restrict__ boxes,
                  const long long* __restrict__ labels, const float* __restrict__ tboxes,
                  const int* __restrict__ num_targets, const int* __restrict__ match_pred,
                  const float* __restrict__ scales, int P, int C, int Tmax, int bg,
                  long long* __restrict__ tc_matched /*[B,P]*/, long long* __restrict__ tc_final /*[B,P]*/,
                  long long* __restrict__ pred_sorted /*[B,Tmax]*/, long long* __restrict__ tgt_sorted /*[B,Tmax]*/,
                  float* __restrict__ losses /*[B,4]: ce,bg,bbox,giou*/, float* __restrict__ dsims /*[B,P,C]*/,
                  float* __restrict__ dl1 /*[B,Tmax,4]*/, float* __restrict__ dgiou /*[B,Tmax,4]*/, float inv_batch) {
  pdl_grid_wait();
  extern __shared__ __align__(16) unsigned char smraw[];
  float* pbox = reinterpret_cast<float*>(smraw);          // [P][4]
  int* tc = reinterpret_cast<int*>(pbox + 4 * P);         // [P]
  float* red = reinterpret_cast<float*>(tc + P);          // [8]
  __shared__ int s_npos;

  const int b = blockIdx.x;
  const int T = clamp_targets(num_targets[b], Tmax);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* mp = match_pred + 1LL * b * Tmax;

  for (int i = tid; i < 4 * P; i += LOSS_THREADS) pbox[i] = boxes[1LL * b * P * 4 + i];
  for (int p = tid; p < P; p += LOSS_THREADS) tc[p] = bg;
  if (tid == 0) s_npos = 0;
  __syncthreads();
  // reference src/matcher.py:148-157: target_classes[pred] = labels[tgt]
  for (int t = tid; t < T; t += LOSS_THREADS) {
    const int p = mp[t];
    if (p >= 0) tc[p] = static_cast<int>(labels[1LL * b * Tmax + t]);
  }
  // reference src/matcher.py:138-146: indices sorted by prediction index (rank by counting, T <= Tmax small)
  for (int t = tid; t < Tmax; t += LOSS_THREADS) {
    if (t < T && mp[t] >= 0) {
      int rank = 0;
      for (int k = 0; k < T; ++k) rank += (mp[k] >= 0 && mp[k] < mp[t]) ? 1 : 0;
      pred_sorted[1LL * b * Tmax + rank] = mp[t];
      tgt_sorted[1LL * b * Tmax + rank] = t;
    }
  }
  for (int t = T + tid; t < Tmax; t += LOSS_THREADS) {
    pred_sorted[1LL * b * Tmax + t] = -1;
    tgt_sorted[1LL * b * Tmax + t] = -1;
  }
  __syncthreads();
  for (int p = tid; p < P; p += LOSS_THREADS) tc_matched[1LL * b * P + p] = tc[p];

  // ---- box losses on matched pairs (reference src/losses.py:42-69) -----------------------------------
  float l1_acc = 0.f, gi_acc = 0.f;
  const float inv_t = 1.0f / static_cast<float>(T);
  for (int t = tid; t < T; t += LOSS_THREADS) {
    const int p = mp[t];
    float g1[4] = {0.f, 0.f, 0.f, 0.f}, g2[4] = {0.f, 0.f, 0.f, 0.f};
    if (p >= 0) {
      const Box s = {pbox[4 * p], pbox[4 * p + 1], pbox[4 * p + 2], pbox[4 * p + 3]};
      const float* tp = tboxes + (1LL * b * Tmax + t) * 4;
      const Box tb = {tp[0], tp[1], tp[2], tp[3]};
      const float d0 = s.x0 - tb.x0, d1 = s.y0 - tb.y0, d2 = s.x1 - tb.x1, d3 = s.y1 - tb.y1;
      l1_acc += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
      const float sc = inv_t * inv_batch;
      g1[0] = (d0 > 0.f ? 1.f : (d0 < 0.f ? -1.f : 0.f)) * sc;
      g1[1] = (d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f)) * sc;
      g1[2] = (d2 > 0.f ? 1.f : (d2 < 0.f ? -1.f : 0.f)) * sc;
      g1[3] = (d3 > 0.f ? 1.f : (d3 < 0.f ? -1.f : 0.f)) * sc;
      float lg;
      giou_loss_grad(s, tb, &lg, g2);
      gi_acc += lg;
#pragma unroll
      for (int k = 0; k < 4; ++k) g2[k] *= sc;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      dl1[(1LL * b * Tmax + t) * 4 + k] = g1[k];
      dgiou[(1LL * b * Tmax + t) * 4 + k] = g2[k];
    }
  }
  for (int t = T + tid; t < Tmax; t += LOSS_THREADS)
    for (int k = 0; k < 4; ++k) { dl1[(1LL * b * Tmax + t) * 4 + k] = 0.f; dgiou[(1LL * b * Tmax + t) * 4 + k] = 0.f; }
  const float l1_sum = block_sum(l1_acc, red);
  const float gi_sum = block_sum(gi_acc, red);

  // ---- IoU > 0.85 label sweep, in index order, labels written in iteration i visible to i' > i --------
  // (reference src/losses.py:100-106).  Two barriers per labelled row keep every thread's
  // view of tc[] identical to the sequential loop.
  for (int i = 0; i < P; ++i) {
    const int lab = tc[i];
    if (lab == bg) continue;       // block-uniform: all threads read the same, already-synchronised value
    __syncthreads();               // everyone has read tc[i] (and all earlier rows) before anyone writes
    const Box bi = {pbox[4 * i], pbox[4 * i + 1], pbox[4 * i + 2], pbox[4 * i + 3]};
    for (int j = tid; j < P; j += LOSS_THREADS) {
      const Box bj = {pbox[4 * j], pbox[4 * j + 1], pbox[4 * j + 2], pbox[4 * j + 3]};
      float u;
      if (iou_union(bi, bj, &u) > 0.85f) tc[j] = lab;
    }
    __syncthreads();
  }
  int npos_local = 0;
  for (int p = tid; p < P; p += LOSS_THREADS) {
    tc_final[1LL * b * P + p] = tc[p];
    npos_local += tc[p] != bg;
  }
  atomicAdd(&s_npos, npos_local);
  __syncthreads();
  if (tid == 0) {
    float* L = losses + 1LL * b * LOSS_WS;
    L[2] = l1_sum * inv_t;
    L[3] = gi_sum * inv_t;
    L[4] = static_cast<float>(s_npos);      // exact: P < 2^24
  }
}

// ---- class loss + gradient (reference src/losses.py:16-40; BCELoss on |sim| with class weights) ---------------
// Grid (LOSS_CHUNKS, B): a CTA owns a slice of an image's prediction rows, a warp one row at a time.  A row's
// similarities are all fetched before the first is used (the SM issues in order; one CTA per image with a load per
// loop iteration spent 130 us on exposed latencies).  Partial sums go to the per-image workspace; loss_reduce_kernel
// adds them in a fixed order, so the result is deterministic.
constexpr int LOSS_MAXC_PER_LANE = 8;   // C <= 256

__global__ void __launch_bounds__(LOSS_THREADS)
class_loss_kernel(const float* __restrict__ sims, const long long* __restrict__ tc_final,
                  const float* __restrict__ scales, int P, int C, int bg, float* __restrict__ losses,
                  float* __restrict__ dsims, float inv_batch) {
  pdl_grid_wait();
  __shared__ float red[8];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* L = losses + 1LL * b * LOSS_WS;
  const int npos = static_cast<int>(L[4]), nbg = P - npos;
  const float inv_pos = 1.0f / static_cast<float>(npos), inv_bg = 1.0f / static_cast<float>(nbg);
  const int rows = (P + LOSS_CHUNKS - 1) / LOSS_CHUNKS;
  const int p0 = chunk * rows, p1 = min(P, p0 + rows);
  float w[LOSS_MAXC_PER_LANE];
#pragma unroll
  for (int k = 0; k < LOSS_MAXC_PER_LANE; ++k) {
    const int c = lane + 32 * k;
    w[k] = (scales && c < C) ? __ldg(scales + c) : 1.0f;
  }
  float ce_acc = 0.f, bg_acc = 0.f;
  for (int p = p0 + warp; p < p1; p += LOSS_THREADS / 32) {
    const float* srow = sims + (1LL * b * P + p) * C;
    float* drow = dsims + (1LL * b * P + p) * C;
    float sv[LOSS_MAXC_PER_LANE];
#pragma unroll
    for (int k = 0; k < LOSS_MAXC_PER_LANE; ++k) {
      const int c = lane + 32 * k;
      sv[k] = c < C ? srow[c] : 0.f;
    }
    const int lab = static_cast<int>(tc_final[1LL * b * P + p]);
    const bool pos = lab != bg;
    const float rscale = (pos ? inv_pos : inv_bg) * inv_batch;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < LOSS_MAXC_PER_LANE; ++k) {
      const int c = lane + 32 * k;
      if (c < C) {
        const float s = sv[k];
        const float q = fabsf(s);
        const float y = (pos && c == lab) ? 1.0f : 0.0f;
        // torch BCE: (y - 1) * max(log1p(-q), -100) - y * max(log(q), -100), times weight
        const float l = ((y - 1.0f) * fmaxf(log1pf(-q), -100.0f) - y * fmaxf(logf(q), -100.0f)) * w[k];
        const float e = expf(-l);
        const float om = 1.0f - e;
        acc += om * om * l;
        // d/dl [(1-e^-l)^2 l] = 2 (1-e^-l) e^-l l + (1-e^-l)^2 ; dl/dq = w (q - y) / max((1-q) q, 1e-12)
        const float dfdl = 2.0f * om * e * l + om * om;
        const float dldq = w[k] * (q - y) / fmaxf((1.0f - q) * q, 1e-12f);
        const float sgn = s > 0.f ? 1.0f : (s < 0.f ? -1.0f : 0.0f);
        drow[c] = dfdl * dldq * sgn * rscale;
      }
    }
    if (pos) ce_acc += acc; else bg_acc += acc;
  }
  const float ce_sum = block_sum(ce_acc, red);
  const float bg_sum = block_sum(bg_acc, red);
  if (threadIdx.x == 0) {
    L[8 + 2 * chunk] = ce_sum;
    L[8 + 2 * chunk + 1] = bg_sum;
  }
}

// per image: class losses from the chunk partials (fixed order); then the mean over images (fixed order)
__global__ void loss_reduce_kernel(float* __restrict__ per_image, int B, int P, float* __restrict__ out4) {
  pdl_grid_wait();
  const int k = threadIdx.x;
  if (k < 2) {
    for (int b = 0; b < B; ++b) {
      float* L = per_image + 1LL * b * LOSS_WS;
      float s = 0.f;
      for (int c = 0; c < LOSS_CHUNKS; ++c) s += L[8 + 2 * c + k];
      const float npos = L[4];
      // mean over positive / background rows (NaN if there are none, like the reference)
      L[k] = s * (1.0f / (k == 0 ? npos : static_cast<float>(P) - npos));
    }
  }
  __syncwarp();
  if (k < 4) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += per_image[1LL * b * LOSS_WS + k];
    out4[k] = s / static_cast<float>(B);
  }
}

// d(sims) = saved * (row positive ? g_ce : g_bg);  d(boxes)[b, pred] = g_bbox * dl1 + g_giou * dgiou
__global__ void loss_backward_kernel(const float* __restrict__ dsims_unit, const long long* __restrict__ tc_final,
                                     const int* __restrict__ match_pred, const float* __restrict__ dl1,
                                     const float* __restrict__ dgiou, const float* __restrict__ gup /*[4]*/,
                                     int B, int P, int C, int Tmax, int bg, float* __restrict__ dsims,
                                     float* __restrict__ dboxes) {
  pdl_grid_wait();
  const float g_ce = gup[0], g_bg = gup[1], g_l1 = gup[2], g_gi = gup[3];
  const long long n = 1LL * B * P * C;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) {
    const long long row = i / C;
    dsims[i] = dsims_unit[i] * (tc_final[row] != bg ? g_ce : g_bg);
  }
  // dboxes is zero-filled by the caller (cudaMemsetAsync) before this kernel; matched rows are unique per image
  const long long m = 1LL * B * Tmax;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < m; i += 1LL * gridDim.x * blockDim.x) {
    const int p = match_pred[i];
    if (p >= 0) {
      const long long b = i / Tmax;
#pragma unroll
      for (int k = 0; k < 4; ++k) dboxes[(b * P + p) * 4 + k] = g_l1 * dl1[i * 4 + k] + g_gi * dgiou[i * 4 + k];
    }
  }
}

}  // namespace owl

using namespace owl;

extern "C" int owl_matcher_cost(const float* sims, const float* boxes, const long long* labels, const float* tboxes,
                                const int* num_targets, float* costT, int B, int P, int C, int Tmax, int* status,
                                float cost_class, float cost_bbox, float cost_giou, void* stream) {
  OWL_CHECK_ARG(sims && boxes && labels && tboxes && num_targets && costT && status, "matcher_cost: null argument");
  OWL_CHECK_ARG(B > 0 && P > 0 && C > 0 && Tmax > 0, "matcher_cost: empty dimension");
  OWL_CHECK_ARG(C <= 256, "matcher_cost: C = %d classes is more than the 256 this kernel keeps in registers", C);
  const size_t smem = sizeof(float) * (COST_ROWS * (C + 1) + COST_ROWS * 4 + Tmax * 4) + sizeof(int) * Tmax;
  OWL_CHECK_ARG(smem <= 100 * 1024, "matcher_cost: C = %d / Tmax = %d need %zu bytes of shared memory", C, Tmax, smem);
  dim3 grid((P + COST_ROWS - 1) / COST_ROWS, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nvec = (C + 31) / 32;
  const bool vec = (C & 3) == 0;
#define OWL_COST_CASE(NV)                                                                                           \
  case NV:                                                                                                          \
    if (vec) {                                                                                                      \
      static SmemOptIn optin;                                                                                       \
      OWL_CUDA(ensure_smem(optin, matcher_cost_kernel<NV, true>, 100 * 1024));                                      \
      OWL_LAUNCH((matcher_cost_kernel<NV, true>), grid, COST_THREADS, smem, s, sims, boxes, labels, tboxes, num_targets, \
                 costT, P, C, Tmax, status, cost_class, cost_bbox, cost_giou);                                                                        \
    } else {                                                                                                        \
      static SmemOptIn optin;                                                                                       \
      OWL_CUDA(ensure_smem(optin, matcher_cost_kernel<NV, false>, 100 * 1024));                                     \
      OWL_LAUNCH((matcher_cost_kernel<NV, false>), grid, COST_THREADS, smem, s, sims, boxes, labels, tboxes, num_targets, \
                 costT, P, C, Tmax, status, cost_class, cost_bbox, cost_giou);                                                                        \
    }                                                                                                               \
    break;
  switch (nvec) {
    OWL_COST_CASE(1) OWL_COST_CASE(2) OWL_COST_CASE(3) OWL_COST_CASE(4) OWL_COST_CASE(5) OWL_COST_CASE(6) OWL_COST_CASE(7)
    OWL_COST_CASE(8)
  }
#undef OWL_COST_CASE
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

static size_t lsap_smem_per_warp(int P, int Tmax) {
  const int Tpad = (Tmax + 4) & ~3;
  const size_t per_warp = sizeof(double) * (2 * (size_t)P + Tpad) + sizeof(short) * (3 * (size_t)P + 3 * Tpad + 8);
  return (per_warp + 15) & ~size_t(15);
}

extern "C" int owl_lsap(const float* costT, const int* num_targets, int B, int P, int Tmax, int* match_pred,
                        int* status, void* stream) {
  OWL_CHECK_ARG(costT && num_targets && match_pred && status, "lsap: null argument");
  OWL_CHECK_ARG(B > 0 && P > 0 && Tmax > 0, "lsap: empty dimension");
  OWL_CHECK_ARG(Tmax <= P, "lsap: more targets (%d) than predictions (%d) is not supported", Tmax, P);
  OWL_CHECK_ARG(P < 32768, "lsap: P must fit int16");
  static const int forced = [] { const char* e = getenv("OWL_LSAP_MODE"); return e ? atoi(e) : 0; }();   // dev: 1 = CTA per image, 2 = warp per image
  if (forced == 1 || (forced == 0 && B <= 2 * num_sms())) {
    // few images: one CTA per image so that the slowest image finishes sooner
    const size_t base1 = lsap_block_base_bytes(P, Tmax);
    OWL_CHECK_ARG(base1 <= 200 * 1024, "lsap: P = %d needs %zu bytes of shared memory", P, base1);
    const int smem_rows = static_cast<int>(std::min<size_t>(Tmax, (200 * 1024 - base1) / (sizeof(float) * P)));
    const size_t smem1 = base1 + sizeof(float) * P * smem_rows;
    constexpr int NT = 256;      // 128 threads: same latency (57 vs 56 us at T = 29), 64: slower (72 us)
    static SmemOptIn optin1;
    OWL_CUDA(ensure_smem(optin1, lsap_block_kernel<NT>, smem1));
    OWL_LAUNCH(lsap_block_kernel<NT>, B, NT, smem1, static_cast<cudaStream_t>(stream), costT, num_targets, P, Tmax,
               match_pred, status, smem_rows);
    OWL_CUDA(cudaGetLastError());
    return OWL_OK;
  }
  const size_t smem = lsap_smem_per_warp(P, Tmax) * LSAP_WARPS;
  OWL_CHECK_ARG(smem <= 227 * 1024, "lsap: P = %d needs %zu bytes of shared memory", P, smem);
  static SmemOptIn optin;
  OWL_CUDA(ensure_smem(optin, lsap_kernel, smem));
  OWL_LAUNCH(lsap_kernel, (B + LSAP_WARPS - 1) / LSAP_WARPS, LSAP_WARPS * 32, smem, static_cast<cudaStream_t>(stream), 
      costT, num_targets, B, P, Tmax, match_pred, status);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_match_loss(const float* sims, const float* boxes, const long long* labels, const float* tboxes,
                              const int* num_targets, const int* match_pred, const float* scales, int B, int P,
                              int C, int Tmax, int bg_label, long long* tc_matched, long long* tc_final,
                              long long* pred_sorted, long long* tgt_sorted, float* losses_per_image,
                              float* losses_mean4, float* dsims_unit, float* dl1, float* dgiou, void* stream) {
  OWL_CHECK_ARG(sims && boxes && labels && tboxes && num_targets && match_pred && tc_matched && tc_final &&
                    pred_sorted && tgt_sorted && losses_per_image && losses_mean4 && dsims_unit && dl1 && dgiou,
                "match_loss: null argument");
  OWL_CHECK_ARG(B > 0 && P > 0 && C > 0 && Tmax > 0, "match_loss: empty dimension");
  OWL_CHECK_ARG(C <= 32 * LOSS_MAXC_PER_LANE, "match_loss: C = %d classes (max %d)", C, 32 * LOSS_MAXC_PER_LANE);
  const size_t smem = sizeof(float) * 4 * P + sizeof(int) * P + sizeof(float) * 8;
  OWL_CHECK_ARG(smem <= 200 * 1024, "match_loss: P = %d too large", P);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static SmemOptIn optin;
  OWL_CUDA(ensure_smem(optin, match_loss_kernel, smem));
  OWL_LAUNCH(match_loss_kernel, B, LOSS_THREADS, smem, s, sims, boxes, labels, tboxes, num_targets, match_pred, scales, P, C,
                                                  Tmax, bg_label, tc_matched, tc_final, pred_sorted, tgt_sorted,
                                                  losses_per_image, dsims_unit, dl1, dgiou, 1.0f / B);
  OWL_CUDA(cudaGetLastError());
  OWL_LAUNCH(class_loss_kernel, dim3(LOSS_CHUNKS, B), LOSS_THREADS, 0, s, sims, tc_final, scales, P, C, bg_label,
             losses_per_image, dsims_unit, 1.0f / B);
  OWL_CUDA(cudaGetLastError());
  OWL_LAUNCH(loss_reduce_kernel, 1, 32, 0, s, losses_per_image, B, P, losses_mean4);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

extern "C" int owl_loss_backward(const float* dsims_unit, const long long* tc_final, const int* match_pred,
                                 const float* dl1, const float* dgiou, const float* upstream4, int B, int P, int C,
                                 int Tmax, int bg_label, float* dsims, float* dboxes, void* stream) {
  OWL_CHECK_ARG(dsims_unit && tc_final && match_pred && dl1 && dgiou && upstream4 && dsims && dboxes,
                "loss_backward: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OWL_CUDA(cudaMemsetAsync(dboxes, 0, sizeof(float) * 4 * (size_t)B * P, s));
  const long long n = 1LL * B * P * C;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148LL * 16));
  OWL_LAUNCH(loss_backward_kernel, blocks, 256, 0, s, dsims_unit, tc_final, match_pred, dl1, dgiou, upstream4, B, P, C, Tmax,
                                              bg_label, dsims, dboxes);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
