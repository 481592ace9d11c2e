#include "common.h"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace owl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return static_cast<int>(e);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // libcuda is reached through the runtime, so the library links without -lcuda and loads on a
    // machine without a driver (the CPU-only build box).
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                        uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return OWL_ERR_DRIVER;
  }
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p not 16-byte aligned", base);
  OWL_CHECK_ARG((row_stride * 2) % 16 == 0, "tensor map: row stride %llu elements is not a multiple of 16 bytes",
                (unsigned long long)row_stride);
  OWL_CHECK_ARG(batches <= 1 || (batch_stride * 2) % 16 == 0, "tensor map: batch stride not a multiple of 16 bytes");
  OWL_CHECK_ARG(inner > 0 && rows > 0 && batches > 0, "tensor map: empty dimension");
  cuuint64_t dims[3] = {inner, rows, batches};
  cuuint64_t strides[2] = {row_stride * 2, (batches > 1 ? batch_stride : rows * row_stride) * 2};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner %llu rows %llu batches %llu ld %llu)",
              static_cast<int>(r), (unsigned long long)inner, (unsigned long long)rows,
              (unsigned long long)batches, (unsigned long long)row_stride);
    return OWL_ERR_DRIVER;
  }
  return OWL_OK;
}

int make_tensor_map_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride,
                        uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return OWL_ERR_DRIVER;
  }
  OWL_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base %p not 16-byte aligned", base);
  OWL_CHECK_ARG((row_stride * 4) % 16 == 0 && box_inner * 4 <= 128, "tensor map (f32): bad row stride / box");
  cuuint64_t dims[3] = {inner, rows, 1};
  cuuint64_t strides[2] = {row_stride * 4, rows * row_stride * 4};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f32) failed with CUresult %d", static_cast<int>(r));
    return OWL_ERR_DRIVER;
  }
  return OWL_OK;
}

static L2Window g_l2_window = {nullptr, 0, 0.f};
const L2Window& l2_window() { return g_l2_window; }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OWL_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n = 148;
  }
  return n;
}

}  // namespace owl

// Zero-fill on the stream (a memset node when captured into a CUDA graph): the atomically accumulated buffers of a step
// (flat gradient buffer, dqn / dcl scratch, matcher status word) are cleared through this, not through framework kernels.
extern "C" int owl_zero(void* ptr, long long bytes, void* stream) {
  if (!ptr || bytes < 0) { owl::set_error("owl_zero: bad arguments"); return owl::OWL_ERR_ARG; }
  if (bytes == 0) return owl::OWL_OK;
  OWL_CUDA(cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream)));
  return owl::OWL_OK;
}

extern "C" const char* owl_last_error(void) { return owl::g_err; }
extern "C" int owl_abi_version(void) { return 6; }

extern "C" int owl_l2_persist(const void* base, long long bytes, float hit_ratio) {
  using namespace owl;
  // Best effort: the window is a hint.  Any failure (no set-aside on the device, a limit the driver refuses, MIG)
  // leaves the launches untagged instead of failing the caller.
  static long long limit_set = -1;
  g_l2_window = {nullptr, 0, 0.f};
  if (base == nullptr || bytes <= 0) return OWL_OK;
  int dev = 0, max_persist = 0, max_window = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev) != cudaSuccess ||
      max_persist <= 0 || max_window <= 0) {
    (void)cudaGetLastError();
    return OWL_OK;
  }
  const long long win = bytes < max_window ? bytes : max_window;
  const long long want = win < max_persist ? win : max_persist;
  if (want != limit_set) {   // not capturable: callers set the window before they capture a graph (Engine.forward)
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, static_cast<size_t>(want)) != cudaSuccess) {
      (void)cudaGetLastError();
      return OWL_OK;
    }
    limit_set = want;
  }
  // a window larger than the set-aside is sampled with hit_ratio so the persisting lines do not thrash each other
  float ratio = hit_ratio > 0.f ? hit_ratio : 1.0f;
  if (win > want) ratio *= static_cast<float>(want) / static_cast<float>(win);
  g_l2_window = {const_cast<void*>(base), static_cast<size_t>(win), ratio};
  return OWL_OK;
}
