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

// A round trip through the switch is microseconds: every thread keeps FOUR 16-byte reductions in flight before it
// stores the first result, and the grid is sized so that a thread makes one or two such passes.
__global__ void __launch_bounds__(256)
allreduce_multimem_kernel(float* __restrict__ mc, long long n4_lo, long long n4_hi) {
  pdl_grid_wait();
  const long long stride = 1LL * gridDim.x * blockDim.x;
  long long i = n4_lo + blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4_hi; i += 4 * stride) {
    const float4 a = mm_ld_reduce(mc + 4 * i), b = mm_ld_reduce(mc + 4 * (i + stride));
    const float4 c = mm_ld_reduce(mc + 4 * (i + 2 * stride)), d = mm_ld_reduce(mc + 4 * (i + 3 * stride));
    mm_st(mc + 4 * i, a);
    mm_st(mc + 4 * (i + stride), b);
    mm_st(mc + 4 * (i + 2 * stride), c);
    mm_st(mc + 4 * (i + 3 * stride), d);
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
  // the kernel runs alone between the backward pass and AdamW: fill the machine (8 CTAs of 256 threads per SM)
  const unsigned blocks = static_cast<unsigned>(std::min<long long>((threads / 4 + 255) / 256 + 1, 8LL * num_sms()));
  OWL_LAUNCH(allreduce_multimem_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), multicast_ptr, lo, hi);
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}
