// Detection post-processing on the device (reference src/models.py:122-146 `PostProcess`):
// best class per prediction, confidence threshold, class-aware NMS (torchvision batched_nms, coordinate trick),
// survivors in decreasing-score order.  One CTA per image, everything in shared memory, no host round trip; the
// reference is batch-1 with boolean-mask indexing (a sync each) and a CPU/torchvision NMS.
//
// Bit-exactness against the reference's CPU path (oracle/postprocess_oracle.py): every fp32 step of torchvision's
// nms is a separately rounded operation, mirrored here with __fadd_rn / __fsub_rn / __fmul_rn / __fdiv_rn; the
// score order is the stable descending sort (ties keep prediction order), realised as a bitonic sort of 64-bit
// keys (~ordered(score) << 32 | prediction index); the IoU threshold is compared in double like the C++ kernel.
#include "common.h"
#include <cuda_runtime.h>

namespace owl {

constexpr int PP_THREADS = 512;

// monotonic float -> uint map (larger float <=> larger uint)
__device__ __forceinline__ uint32_t pp_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(PP_THREADS)
postprocess_kernel(const float* __restrict__ boxes, const float* __restrict__ sims, int P, int C, int Ppow2,
                   float conf_thr, double iou_thr, float* __restrict__ out_boxes, long long* __restrict__ out_classes,
                   float* __restrict__ out_scores, int* __restrict__ out_count) {
  pdl_grid_wait();
  extern __shared__ __align__(16) unsigned char pp_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(pp_smem);          // [Ppow2]
  float4* obox = reinterpret_cast<float4*>(keys + Ppow2);                              // [P] offset boxes, sorted order
  float* area = reinterpret_cast<float*>(obox + P);                                    // [P]
  float* score = area + P;                                                             // [P] by prediction
  int* cls = reinterpret_cast<int*>(score + P);                                        // [P] by prediction
  unsigned char* supp = reinterpret_cast<unsigned char*>(cls + P);                     // [P] sorted order
  __shared__ unsigned int s_maxc;   // ordered-uint of the largest coordinate among the candidates
  __shared__ int s_n, s_kept;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* bx = boxes + 4LL * b * P;
  const float* sm = sims + 1LL * b * P * C;
  if (tid == 0) { s_maxc = 0u; s_n = 0; s_kept = 0; }

  // 1. best class per prediction: first maximum wins (torch.max(dim=1) on the CPU)
  for (int p = warp; p < P; p += PP_THREADS / 32) {
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = sm[1LL * p * C + c];
      if (arg == 0x7fffffff || v > best) { best = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
    }
    if (lane == 0) { score[p] = best; cls[p] = arg; }
  }
  __syncthreads();

  // 2. confidence threshold -> sort keys; largest coordinate of the candidates
  int n_local = 0;
  unsigned int maxc_local = 0u;
  for (int p = tid; p < Ppow2; p += PP_THREADS) {
    unsigned long long key = ~0ull;
    if (p < P && score[p] > conf_thr) {
      key = (static_cast<unsigned long long>(~pp_ordered(score[p])) << 32) | static_cast<unsigned int>(p);
      const float4 q = *reinterpret_cast<const float4*>(bx + 4 * p);
      maxc_local = max(maxc_local, max(max(pp_ordered(q.x), pp_ordered(q.y)), max(pp_ordered(q.z), pp_ordered(q.w))));
      ++n_local;
    }
    keys[p] = key;
  }
  if (n_local) { atomicAdd(&s_n, n_local); atomicMax(&s_maxc, maxc_local); }
  __syncthreads();
  const int n = s_n;

  // 3. bitonic sort, ascending keys = descending score, ties by ascending prediction index (stable order)
  for (int k = 2; k <= Ppow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < Ppow2; i += PP_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], c2 = keys[l];
          const bool up = (i & k) == 0;
          if ((a > c2) == up) { keys[i] = c2; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }

  // 4. torchvision's coordinate trick: box + class * (max coordinate + 1); areas
  {
    const unsigned int mu = s_maxc;
    const float maxc = __uint_as_float((mu & 0x80000000u) ? (mu & 0x7fffffffu) : ~mu);
    const float step = __fadd_rn(maxc, 1.0f);
    for (int k = tid; k < n; k += PP_THREADS) {
      const int p = static_cast<int>(keys[k] & 0xffffffffu);
      const float4 q = *reinterpret_cast<const float4*>(bx + 4 * p);
      const float off = __fmul_rn(static_cast<float>(cls[p]), step);
      float4 o;
      o.x = __fadd_rn(q.x, off); o.y = __fadd_rn(q.y, off); o.z = __fadd_rn(q.z, off); o.w = __fadd_rn(q.w, off);
      obox[k] = o;
      area[k] = __fmul_rn(__fsub_rn(o.z, o.x), __fsub_rn(o.w, o.y));
      supp[k] = 0;
    }
  }
  __syncthreads();

  // 5. greedy sweep in score order (torchvision nms_kernel_impl): a kept box suppresses every later box whose IoU
  //    with it exceeds the threshold.  One barrier per candidate keeps every thread's view of supp[] sequential.
  for (int i = 0; i < n; ++i) {
    if (!supp[i]) {                         // block-uniform (read after the barrier below / above)
      const float4 bi = obox[i];
      const float ai = area[i];
      for (int j = i + 1 + tid; j < n; j += PP_THREADS) {
        if (supp[j]) continue;
        const float4 bj = obox[j];
        const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
        const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
        const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
        const float inter = __fmul_rn(w, h);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, area[j]), inter));
        if (static_cast<double>(ovr) > iou_thr) supp[j] = 1;
      }
    }
    __syncthreads();
  }

  // 6. survivors, in order: one warp compacts with ballots
  if (warp == 0) {
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
      const int k = k0 + lane;
      const bool keep = k < n && !supp[k];
      const unsigned int m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int dst = base + __popc(m & ((1u << lane) - 1u));
        const int p = static_cast<int>(keys[k] & 0xffffffffu);
        *reinterpret_cast<float4*>(out_boxes + 4LL * (1LL * b * P + dst)) = *reinterpret_cast<const float4*>(bx + 4 * p);
        out_classes[1LL * b * P + dst] = cls[p];
        out_scores[1LL * b * P + dst] = score[p];
      }
      base += __popc(m);
    }
    if (lane == 0) out_count[b] = base;
  }
}

}  // namespace owl

using namespace owl;

extern "C" int owl_postprocess(const float* boxes, const float* sims, int B, int P, int C, float confidence_threshold,
                               double iou_threshold, float* out_boxes, long long* out_classes, float* out_scores,
                               int* out_count, void* stream) {
  OWL_CHECK_ARG(boxes && sims && out_boxes && out_classes && out_scores && out_count, "postprocess: null argument");
  OWL_CHECK_ARG(B > 0 && P > 0 && C > 0, "postprocess: empty dimension");
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0,
                "postprocess: boxes must be 16-byte aligned");
  int pp2 = 32;
  while (pp2 < P) pp2 <<= 1;
  const size_t smem = sizeof(unsigned long long) * pp2 + (sizeof(float4) + 3 * sizeof(float)) * P + P + 16;
  OWL_CHECK_ARG(smem <= 200 * 1024, "postprocess: P = %d needs %zu bytes of shared memory", P, smem);
  static SmemOptIn optin;
  OWL_CUDA(ensure_smem(optin, postprocess_kernel, smem));
  OWL_LAUNCH(postprocess_kernel, B, PP_THREADS, smem, static_cast<cudaStream_t>(stream), boxes, sims, P, C, pp2,
             confidence_threshold, iou_threshold, out_boxes, out_classes, out_scores, out_count);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
