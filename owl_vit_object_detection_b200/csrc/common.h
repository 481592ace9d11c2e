// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>

#include "../../include/owl_b200.h"

namespace owl {

enum : int {
  OWL_OK = 0,
  OWL_ERR_ARG = -1,       // invalid argument
  OWL_ERR_DRIVER = -2,    // driver entry point / tensor-map encode failure
  OWL_ERR_UNSUPPORTED = -3,
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);   // records + returns the (positive) cudaError_t

#define OWL_CUDA(expr)                                     \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return ::owl::cuda_fail(_e, #expr); \
  } while (0)

#define OWL_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::owl::set_error(__VA_ARGS__);    \
      return ::owl::OWL_ERR_ARG;        \
    }                                   \
  } while (0)

// 3-D fp16 tensor map, 128-byte swizzle, zero OOB fill.
// dims: inner (contiguous), rows, batches; strides in elements.
int make_tensor_map_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                        uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows);

// same for fp32 elements (box_inner * 4 bytes <= 128: one swizzle row)
int make_tensor_map_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride,
                        uint32_t box_inner, uint32_t box_rows);

int num_sms();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is PER DEVICE: a process that drives several GPUs has to opt in on
// each of them.  One SmemOptIn per call site (kernel instantiation) remembers the largest size configured per device.
struct SmemOptIn { size_t bytes[64] = {}; };
template <class Kernel>
cudaError_t ensure_smem(SmemOptIn& st, Kernel kernel, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  dev &= 63;
  if (bytes > st.bytes[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return e;
    st.bytes[dev] = bytes;
  }
  return cudaSuccess;
}

// Programmatic dependent launch (PDL): every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_grid_wait() (griddepcontrol.wait) before it
// touches global memory, then releases its own dependents (griddepcontrol.launch_dependents).  The next kernel's
// CTAs can therefore be scheduled onto SMs as the current kernel's CTAs retire and run their set-up (barrier init,
// TMEM allocation, descriptor prefetch) under the current kernel's tail; a step is ~170 short kernels, so the
// fill / drain bubbles are a measurable part of it.  OWL_PDL=0 in the environment turns the attribute off.
bool pdl_enabled();

// L2 persistence: one address window (the fp32 residual stream, read three times and written twice per encoder layer)
// is tagged cudaAccessPropertyPersisting on every launch of this library (owl_l2_persist, common.cu), so it stays in
// the L2 set-aside instead of being evicted by the streaming GEMM operands between its uses.
struct L2Window { void* base; size_t bytes; float hit_ratio; };
const L2Window& l2_window();

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int n = 0;
  const L2Window& w = l2_window();
  if (w.bytes > 0) {
    attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[n].val.accessPolicyWindow.base_ptr = w.base;
    attr[n].val.accessPolicyWindow.num_bytes = w.bytes;
    attr[n].val.accessPolicyWindow.hitRatio = w.hit_ratio;
    attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define OWL_LAUNCH(kernel, grid, block, smem, stream, ...) \
  OWL_CUDA(::owl::launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, 1, __VA_ARGS__))

#ifdef __CUDACC__
// first statement of every kernel (after purely local set-up): wait for the producer grids, then let our own
// dependents be scheduled as SMs free up
__device__ __forceinline__ void pdl_grid_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

}  // namespace owl
