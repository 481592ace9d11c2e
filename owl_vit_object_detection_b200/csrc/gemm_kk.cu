// K-major A, K-major B (forward linears: y = x W^T).
#include "gemm_plan.h"
namespace owl {
int gemm_launch_kk(const GemmPlan& p, cudaStream_t s) {
  if (p.epilogue == 0) {
    if (p.act == ACT_NONE) { OWL_GEMM_DISPATCH_BN_CM(false, false, EpiF16<ACT_NONE>, p.p16) }
    if (p.act == ACT_QGELU) { OWL_GEMM_DISPATCH_BN_CM(false, false, EpiF16<ACT_QGELU>, p.p16) }
    // attention backward: K = head_dim, one k-block per tile, the epilogue is the kernel -> two CTAs per SM
    if (p.act == ACT_EXP_ROW && p.bn == 128 && p.cm == 1 && p.gs.K <= 128)
      return gemm_launch_one<128, false, false, EpiF16<ACT_EXP_ROW>, 1, 2>(p, p.p16, s);
    if (p.act == ACT_SMAX_GRAD && p.bn == 128 && p.cm == 1 && p.gs.K <= 128)
      return gemm_launch_one<128, false, false, EpiF16<ACT_SMAX_GRAD>, 1, 2>(p, p.p16, s);
    if (p.act == ACT_EXP_ROW) { OWL_GEMM_DISPATCH_BN(false, false, EpiF16<ACT_EXP_ROW>, p.p16) }
    if (p.act == ACT_SMAX_GRAD) { OWL_GEMM_DISPATCH_BN(false, false, EpiF16<ACT_SMAX_GRAD>, p.p16) }
    if (p.act == ACT_GELU) { OWL_GEMM_DISPATCH_BN(false, false, EpiF16<ACT_GELU>, p.p16) }
  }
  if (p.epilogue == 1) { OWL_GEMM_DISPATCH_BN_CM(false, false, EpiF32, p.p32) }
  if (p.epilogue == 2 && p.bn == 256) return gemm_launch_one<256, false, false, EpiPool3>(p, p.pp, s);
  set_error("gemm(kk): unsupported epilogue %d / act %d / tile %d", p.epilogue, p.act, p.bn);
  return OWL_ERR_UNSUPPORTED;
}
}  // namespace owl
