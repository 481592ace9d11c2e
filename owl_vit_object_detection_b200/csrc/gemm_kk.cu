// K-major A, K-major B (forward linears: y = x W^T).
#include "gemm_plan.h"
namespace owl {
int gemm_launch_kk(const GemmPlan& p, cudaStream_t s) {
  if (p.epilogue == 0) { OWL_GEMM_DISPATCH_BN(false, false, EpiF16, p.p16) }
  if (p.epilogue == 1) { OWL_GEMM_DISPATCH_BN(false, false, EpiF32, p.p32) }
  if (p.epilogue == 2) {
    if (p.bn == 256) return gemm_launch_one<256, false, false, EpiPool3>(p, p.pp, s);
    if (p.bn == 192) return gemm_launch_one<192, false, false, EpiPool3>(p, p.pp, s);
  }
  set_error("gemm(kk): unsupported epilogue %d / tile %d", p.epilogue, p.bn);
  return OWL_ERR_UNSUPPORTED;
}
}  // namespace owl
