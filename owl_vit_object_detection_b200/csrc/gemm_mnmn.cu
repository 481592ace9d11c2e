// MN-major A, MN-major B (wgrad: dW = dY^T X with both as stored; attention dV = P^T dO, dK = dS^T Q).
#include "gemm_plan.h"
namespace owl {
int gemm_launch_mnmn(const GemmPlan& p, cudaStream_t s) {
  if (p.epilogue == 0 && p.act == ACT_NONE) { OWL_GEMM_DISPATCH_BN(true, true, EpiF16<ACT_NONE>, p.p16) }
  if (p.epilogue == 1) { OWL_GEMM_DISPATCH_BN(true, true, EpiF32, p.p32) }
  set_error("gemm(mn,mn): unsupported epilogue %d / act %d", p.epilogue, p.act);
  return OWL_ERR_UNSUPPORTED;
}
}  // namespace owl
