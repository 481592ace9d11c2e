// Gradient all-reduce through the NVSwitch (NVLS): the one exchange step of the data-parallel path (SURVEY §8e;
// reference main.py:90-91 has no distributed code at all - the reduction is what "N GPUs == one big batch" needs).
//
// The flat fp32 gradient buffer lives in SYMMETRIC memory: the same allocation on every rank, mapped by every GPU, plus
// one MULTICAST address that stands for "this offset on all ranks".  Two-shot all-reduce in one kernel per rank:
//   rank r owns slice r of the buffer:  v = multimem.ld_reduce.add [mc + i]   (the switch reads the 16 bytes from every
//                                       GPU and returns their fp32 sum)
//                                       multimem.st [mc + i] = v              (the switch writes the sum into every
//                                       GPU's copy, the issuing one included)
// Per GPU 1/N of the buffer crosses NVLink in each direction per phase instead of NCCL's ring traffic, and no SM-side
// reduction code runs at all.  The cross-rank barriers before (every rank's backward has finished writing its
// gradients) and after (every slice has landed everywhere) are the symmetric-memory handle's own signal-pad barriers,
// issued by the host code on the same stream (owl_vit_object_detection_b200/collective.py); all three launches are
// plain stream work and are captured into the step's CUDA graph.
#include "common.h"
#include <algorithm>

namespace owl {

__device__ __forceinline__ float4 mm_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// A round trip through the switch is microseconds: every thread keeps U 16-byte reductions in flight before it stores
// the first result, and the grid is sized so that a thread makes one or two such passes.
template <int U>
__global__ void __launch_bounds__(256)
allreduce_multimem_kernel(float* __restrict__ mc, long long n4_lo, long long n4_hi) {
  pdl_grid_wait();
  const long long stride = 1LL * gridDim.x * blockDim.x;
  long long i = n4_lo + blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n4_hi; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = mm_ld_reduce(mc + 4 * (i + k * stride));
#pragma unroll
    for (int k = 0; k < U; ++k) mm_st(mc + 4 * (i + k * stride), v[k]);
  }
  for (; i < n4_hi; i += stride) mm_st(mc + 4 * i, mm_ld_reduce(mc + 4 * i));
}

}  // namespace owl

extern "C" int owl_allreduce_multimem(float* multicast_ptr, long long n, int rank, int world, void* stream) {
  using namespace owl;
  OWL_CHECK_ARG(multicast_ptr && n > 0 && n % 4 == 0 && world > 0 && rank >= 0 && rank < world,
                "allreduce_multimem: bad arguments");
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0, "allreduce_multimem: unaligned multicast pointer");
  const long long n4 = n / 4;
  const long long per = (n4 + world - 1) / world;
  const long long lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
  if (lo >= hi) return OWL_OK;
  const long long threads = hi - lo;
  // The kernel runs alone between the backward pass and AdamW.  Four reductions in flight per thread; two or eight
  // measure the same (120-124 us for 35.2 MB on 8 GPUs): the switch path, not the issue side, bounds it.
  constexpr int U = 4;
  const unsigned blocks = static_cast<unsigned>(std::min<long long>((threads / U + 255) / 256 + 1, 8LL * num_sms()));
  OWL_LAUNCH(allreduce_multimem_kernel<U>, blocks, 256, 0, static_cast<cudaStream_t>(stream), multicast_ptr, lo, hi);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
