// K-major A, MN-major B (dgrad: dX = dY W with W as stored; attention P V with V as stored).
#include "gemm_plan.h"
namespace owl {
int gemm_launch_kmn(const GemmPlan& p, cudaStream_t s) {
  if (p.epilogue == 0) { OWL_GEMM_DISPATCH_BN(false, true, EpiF16, p.p16) }
  if (p.epilogue == 1) { OWL_GEMM_DISPATCH_BN(false, true, EpiF32, p.p32) }
  set_error("gemm(k,mn): unsupported epilogue %d", p.epilogue);
  return OWL_ERR_UNSUPPORTED;
}
}  // namespace owl
