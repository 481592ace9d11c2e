// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

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

int num_sms();

}  // namespace owl
