// K-major A, MN-major B (dgrad: dX = dY W with W as stored; attention P V with V as stored).
#include "gemm_plan.h"
namespace owl {
int gemm_launch_kmn(const GemmPlan& p, cudaStream_t s) {
  if (p.epilogue == 0) {
    if (p.act == ACT_NONE) { OWL_GEMM_DISPATCH_BN_CM(false, true, EpiF16<ACT_NONE>, p.p16) }
    if (p.act == ACT_QGELU_GRAD) { OWL_GEMM_DISPATCH_BN_CM(false, true, EpiF16<ACT_QGELU_GRAD>, p.p16) }
    if (p.act == ACT_GELU_GRAD) { OWL_GEMM_DISPATCH_BN(false, true, EpiF16<ACT_GELU_GRAD>, p.p16) }
  }
  if (p.epilogue == 1) { OWL_GEMM_DISPATCH_BN_CM(false, true, EpiF32, p.p32) }
  set_error("gemm(k,mn): unsupported epilogue %d / act %d", p.epilogue, p.act);
  return OWL_ERR_UNSUPPORTED;
}
}  // namespace owl
