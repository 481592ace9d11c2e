// Host-side GEMM plan: tensor maps + shape + epilogue parameters, built once and launched many times.
#pragma once
#include "common.h"
#include "gemm_tc.cuh"
#include <algorithm>

namespace owl {

struct GemmPlan {
  CUtensorMap tmA, tmB;
  GemmShape gs;
  int bn;
  int cm;        // CTAs per cluster along M (B tile multicast); 1 = no cluster
  int a_mn, b_mn;
  int epilogue;  // 0 f16, 1 f32, 2 pool3
  EpiF16Params p16;
  int act;
  EpiF32::Params p32;
  EpiPool3::Params pp;
  int grid;
  long long tiles;   // cluster tiles of the whole problem
};

// `bn` = 0 lets the planner choose the N tile.
int gemm_plan_build(const owl_gemm_args& a, int bn, GemmPlan* plan);
int gemm_plan_launch(const GemmPlan& plan, cudaStream_t stream);

// per-major-combination launchers (one translation unit each, so they compile in parallel)
int gemm_launch_kk(const GemmPlan& p, cudaStream_t s);
int gemm_launch_kmn(const GemmPlan& p, cudaStream_t s);
int gemm_launch_mnmn(const GemmPlan& p, cudaStream_t s);

template <int BN, bool A_MN, bool B_MN, class Epi, int CM = 1, int MINB = 1>
int gemm_launch_one(const GemmPlan& p, const typename Epi::Params& ep, cudaStream_t s) {
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, Epi, CM, MINB>;
  static SmemOptIn optin;  // per instantiation
  OWL_CUDA(ensure_smem(optin, kern, gemm_smem_bytes(BN, CM, MINB)));
  // persistent grid: one CTA per SM slot the flavour is built for
  const int grid = MINB == 1 ? p.grid : static_cast<int>(std::min<long long>(p.tiles, 1LL * MINB * num_sms()));
  OWL_CUDA(launch_pdl(kern, dim3(grid), dim3(gemm_threads(BN)), gemm_smem_bytes(BN, CM, MINB), s, CM, p.tmA, p.tmB, p.gs, ep));
  OWL_CUDA(cudaGetLastError());
  return OWL_OK;
}

#define OWL_GEMM_DISPATCH_BN(A_MN, B_MN, EPI, params)                                         \
  switch (p.bn * 8 + p.cm) {                                                                  \
    case 64 * 8 + 1:  return gemm_launch_one<64, A_MN, B_MN, EPI>(p, params, s);              \
    case 128 * 8 + 1: return gemm_launch_one<128, A_MN, B_MN, EPI>(p, params, s);             \
    case 192 * 8 + 1: return gemm_launch_one<192, A_MN, B_MN, EPI>(p, params, s);             \
    case 256 * 8 + 1: return gemm_launch_one<256, A_MN, B_MN, EPI>(p, params, s);             \
    default: set_error("gemm: unsupported N tile %d / cluster %d", p.bn, p.cm); return OWL_ERR_UNSUPPORTED; \
  }
// forward / dgrad GEMMs additionally come in a 2-CTA-cluster flavour (B tile multicast)
// the 192-wide pair flavour stages 96 B rows per CTA: K-major B only (MN-major B moves whole 64-wide chunks)
#define OWL_GEMM_CASE_192_PAIR(A_MN, B_MN, EPI, params) \
    case 192 * 8 + 2: if constexpr (!(B_MN)) return gemm_launch_one<192, A_MN, false, EPI, 2>(p, params, s); else break;
#define OWL_GEMM_DISPATCH_BN_CM(A_MN, B_MN, EPI, params)                                      \
  switch (p.bn * 8 + p.cm) {                                                                  \
    case 64 * 8 + 1:  return gemm_launch_one<64, A_MN, B_MN, EPI>(p, params, s);              \
    case 128 * 8 + 1: return gemm_launch_one<128, A_MN, B_MN, EPI>(p, params, s);             \
    case 192 * 8 + 1: return gemm_launch_one<192, A_MN, B_MN, EPI>(p, params, s);             \
    case 256 * 8 + 1: return gemm_launch_one<256, A_MN, B_MN, EPI>(p, params, s);             \
    case 128 * 8 + 2: return gemm_launch_one<128, A_MN, B_MN, EPI, 2>(p, params, s);          \
    OWL_GEMM_CASE_192_PAIR(A_MN, B_MN, EPI, params)                                           \
    case 256 * 8 + 2: return gemm_launch_one<256, A_MN, B_MN, EPI, 2>(p, params, s);          \
    default: set_error("gemm: unsupported N tile %d / cluster %d", p.bn, p.cm); return OWL_ERR_UNSUPPORTED; \
  }

}  // namespace owl
