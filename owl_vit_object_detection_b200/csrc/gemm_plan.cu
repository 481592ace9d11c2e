#include "gemm_plan.h"
#include <stdlib.h>

namespace owl {

static int pick_bn(int M, int N, int G, int split_k, int epilogue, bool b_mn) {
  if (epilogue == 2) return 256;
  const int sms = num_sms();
  const int mb = (M + GEMM_BM - 1) / GEMM_BM;
  int best = 0;
  double best_cost = 0;
  const int cands[4] = {256, 192, 128, 64};
  for (int c = 0; c < 4; ++c) {
    const int bn = cands[c];
    if (bn > 64 && bn >= 2 * N && N > 0) continue;  // mostly padding
    const long long tiles = 1LL * mb * ((N + bn - 1) / bn) * G * split_k;
    const long long waves = (tiles + sms - 1) / sms;
    // per-tile time ~ MMA time (prop. to bn) + a fixed pipeline fill / epilogue tail
    const double cost = static_cast<double>(waves) * (bn + 40.0);
    if (best == 0 || cost < best_cost * 0.999) { best = bn; best_cost = cost; }
  }
  (void)b_mn;
  return best ? best : 64;
}

int gemm_plan_build(const owl_gemm_args& a, int bn, GemmPlan* plan) {
  OWL_CHECK_ARG(a.a && a.b && a.out, "gemm: null operand");
  OWL_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem %d x %d x %d", a.M, a.N, a.K);
  OWL_CHECK_ARG(a.batches_outer >= 1 && a.heads >= 1, "gemm: batches_outer / heads must be >= 1");
  OWL_CHECK_ARG(a.split_k >= 1, "gemm: split_k must be >= 1");
  OWL_CHECK_ARG(a.epilogue >= 0 && a.epilogue <= 2, "gemm: unknown epilogue %d", a.epilogue);
  OWL_CHECK_ARG(a.split_k == 1 || (a.epilogue == 1 && a.out_mode == 2),
                "gemm: split_k > 1 needs the fp32 epilogue in atomic mode");
  OWL_CHECK_ARG(a.epilogue != 2 || (a.N % 3 == 0 && a.argmax && !a.a_mn && !a.b_mn && a.batches_outer * a.heads == 1),
                "gemm: pool3 epilogue needs N %% 3 == 0, K-major operands, no batching and an argmax buffer");
  OWL_CHECK_ARG(!(a.act == 3 || a.act == 4 || a.act == 6) || a.act_src, "gemm: act %d needs act_src", a.act);
  OWL_CHECK_ARG(!(a.act == 5 || a.act == 6) || (a.rowvec && a.rowvec_stride >= a.M), "gemm: act %d needs rowvec", a.act);

  GemmPlan& p = *plan;
  p.a_mn = a.a_mn ? 1 : 0;
  p.b_mn = a.b_mn ? 1 : 0;
  p.epilogue = a.epilogue;
  const int G = a.batches_outer * a.heads;
  p.bn = bn ? bn : pick_bn(a.M, a.N, G, a.split_k, a.epilogue, p.b_mn);
  OWL_CHECK_ARG(p.bn == 64 || p.bn == 128 || p.bn == 192 || p.bn == 256, "gemm: N tile %d not supported", p.bn);
  // 2-CTA clusters (B tile multicast) for the big forward / dgrad GEMMs; args.cluster_m: 0 = auto, 1 = off, 2 = on
  {
    const int mb_ = (a.M + GEMM_BM - 1) / GEMM_BM;
    const bool eligible = !p.a_mn && p.bn >= 128 && a.epilogue != 2 && a.act != 2 && a.act < 4 && mb_ >= 2;
    const long long tiles_ = 1LL * mb_ * ((a.N + p.bn - 1) / p.bn) * G * a.split_k;
    if (a.cluster_m == 2) {
      OWL_CHECK_ARG(eligible && !(p.bn == 192 && p.b_mn), "gemm: cluster_m = 2 is not built for this operand layout / epilogue / tile");
      p.cm = 2;
    } else if (a.cluster_m == 0) {
      // measured on B200 (tools/sweep_gemm.py): the pair flavour wins when the main loop is long (fc2, K = 3072:
      // 43.4 vs 47.4 us at bn = 192; 8192^3: 1.31 vs 1.22 PFLOP/s) and loses on the K = 768 layer shapes, whose
      // time goes to the epilogue rather than the MMA main loop
      p.cm = (eligible && tiles_ >= num_sms() && a.K >= 3072 && !(p.bn == 192 && p.b_mn)) ? 2 : 1;
    } else {
      OWL_CHECK_ARG(a.cluster_m == 1, "gemm: cluster_m must be 0, 1 or 2");
      p.cm = 1;
    }
  }
  OWL_CHECK_ARG(a.epilogue != 2 || a.N <= 256, "gemm: pool3 epilogue supports N <= 256 (got %d)", a.N);
  OWL_CHECK_ARG(!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  OWL_CHECK_ARG(!((a.act == 3 || a.act == 4 || a.act == 6) && a.pre_out), "gemm: act' epilogues cannot also save pre_out");
  OWL_CHECK_ARG(a.act >= 0 && a.act <= 6 && (a.act == 0 || a.epilogue == 0), "gemm: act %d needs the fp16 epilogue", a.act);

  // The third tensor-map dimension enumerates (outer, head) with a common stride when that is
  // expressible; per-head column offsets cover heads packed inside a row (QKV buffer).
  GemmShape& gs = p.gs;
  gs.M = a.M; gs.N = a.N; gs.K = a.K;
  gs.G = G; gs.H = a.heads;
  gs.split_k = a.split_k;
  static const int debug = [] { const char* e = getenv("OWL_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  gs.debug = debug;

  auto build_operand = [&](const void* base, bool mn, int rows_mn, long long ld, long long outer_stride,
                           long long head_stride, int head_col, int box_rows_k_major, CUtensorMap* tm,
                           int* col_off, int* sb, int* sh) -> int {
    // memory matrix: K-major -> [rows_mn, K] ; MN-major -> [K, rows_mn]
    const uint64_t inner = mn ? static_cast<uint64_t>(rows_mn) : static_cast<uint64_t>(a.K);
    const uint64_t rows = mn ? static_cast<uint64_t>(a.K) : static_cast<uint64_t>(rows_mn);
    uint64_t batches = 1, bstride = 0;
    *col_off = head_col;
    *sb = 0; *sh = 0;
    if (G > 1) {
      if (a.heads > 1 && head_stride != 0 && a.batches_outer > 1) {
        // need outer_stride == heads * head_stride to enumerate (outer, head) on one axis
        OWL_CHECK_ARG(outer_stride == head_stride * a.heads,
                      "gemm: outer stride must equal heads * head stride when both vary");
        batches = G; bstride = head_stride; *sb = a.heads; *sh = 1;
      } else if (a.heads > 1 && head_stride != 0) {
        batches = a.heads; bstride = head_stride; *sh = 1;
      } else if (a.batches_outer > 1 && outer_stride != 0) {
        batches = a.batches_outer; bstride = outer_stride; *sb = 1;
      }
    }
    const uint64_t inner_total = inner + static_cast<uint64_t>(head_col) * (a.heads - 1);
    const uint32_t box_rows = mn ? 64u : static_cast<uint32_t>(box_rows_k_major);
    return make_tensor_map_f16(tm, base, inner_total, rows, batches, static_cast<uint64_t>(ld), bstride, 64u,
                               box_rows);
  };
  int rc = build_operand(a.a, p.a_mn, a.M, a.a_ld, a.a_outer_stride, a.a_head_stride, a.a_head_col, GEMM_BM,
                         &p.tmA, &gs.a_col_off, &gs.a_sb, &gs.a_sh);
  if (rc) return rc;
  rc = build_operand(a.b, p.b_mn, a.N, a.b_ld, a.b_outer_stride, a.b_head_stride, a.b_head_col, p.bn / p.cm, &p.tmB,
                     &gs.b_col_off, &gs.b_sb, &gs.b_sh);
  if (rc) return rc;

  const float alpha = a.alpha == 0.0f ? 1.0f : a.alpha;
  if (a.epilogue == 0) {
    EpiF16Params& e = p.p16;
    p.act = a.act;
    e.out = static_cast<__half*>(a.out);
    e.pre_out = static_cast<__half*>(a.pre_out);
    e.bias = a.bias;
    e.dact_src = static_cast<const __half*>(a.act_src);
    e.rowvec = a.rowvec; e.rowvec_stride = a.rowvec_stride;
    e.d_sb = a.act_src_outer_stride; e.d_sh = a.act_src_head_stride;
    e.ldo = static_cast<int>(a.ldo); e.ld_pre = static_cast<int>(a.ld_pre); e.ld_dact = static_cast<int>(a.ld_act_src);
    e.o_sb = a.o_outer_stride; e.o_sh = a.o_head_stride; e.H = a.heads;
    e.alpha = alpha; e.alpha_dev = a.alpha_dev;
    auto al16 = [](const void* q, long long ld) { return !q || ((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (ld * 2) % 16 == 0); };
    e.vec_ok = al16(a.out, a.ldo) && al16(a.pre_out, a.ld_pre) && al16(a.act_src, a.ld_act_src) &&
               (a.o_outer_stride * 2) % 16 == 0 && (a.o_head_stride * 2) % 16 == 0 &&
               (a.act_src_outer_stride * 2) % 16 == 0 && (a.act_src_head_stride * 2) % 16 == 0;
    // bulk-tensor stores of the output tiles (one [M, N] matrix, 16-byte aligned rows; every N tile >= 128 drains
    // 64-column slices per epilogue warp)
    static const bool tma_off = [] { const char* v = getenv("OWL_GEMM_TMA_STORE"); return v && v[0] == '0'; }();
    e.use_tma = (!tma_off && G == 1 && e.vec_ok && p.bn >= 128) ? 1 : 0;
    if (e.use_tma) {
      rc = make_tensor_map_f16(&e.tm_out, a.out, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M), 1,
                               static_cast<uint64_t>(a.ldo), 0, 64, 32);
      if (rc) return rc;
      if (a.pre_out) {
        rc = make_tensor_map_f16(&e.tm_pre, a.pre_out, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M), 1,
                                 static_cast<uint64_t>(a.ld_pre), 0, 64, 32);
        if (rc) return rc;
      }
    }
  } else if (a.epilogue == 1) {
    EpiF32::Params& e = p.p32;
    e.out = static_cast<float*>(a.out);
    e.bias = a.bias; e.resid = a.resid; e.pos = a.pos;
    e.ldo = static_cast<int>(a.ldo); e.ldr = static_cast<int>(a.ldr);
    e.o_sb = a.o_outer_stride; e.o_sh = a.o_head_stride; e.H = a.heads;
    e.mode = a.out_mode; e.rows_per_img = a.rows_per_img; e.alpha = alpha; e.alpha_dev = a.alpha_dev;
    auto al32 = [](const void* q, long long ld) { return !q || ((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (ld * 4) % 16 == 0); };
    e.vec_ok = al32(a.out, a.ldo) && al32(a.resid, a.ldr) && al32(a.pos, a.ldo) &&
               (a.o_outer_stride * 4) % 16 == 0 && (a.o_head_stride * 4) % 16 == 0;
    // TMA residual loads + TMA stores: plain store of one [M, N] matrix (no accumulate / atomics / row remap / pos-emb)
    static const bool tma_off32 = [] { const char* v = getenv("OWL_GEMM_TMA_STORE"); return v && v[0] == '0'; }();
    e.use_tma = (!tma_off32 && G == 1 && e.vec_ok && a.out_mode == 0 && a.split_k == 1 && a.rows_per_img == 0 && !a.pos &&
                 a.N % 32 == 0 && (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0)) ? 1 : 0;
    if (e.use_tma) {
      rc = make_tensor_map_f32(&e.tm_out, a.out, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M),
                               static_cast<uint64_t>(a.ldo), 32, 32);
      if (rc) return rc;
      if (a.resid) {
        rc = make_tensor_map_f32(&e.tm_resid, a.resid, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M),
                                 static_cast<uint64_t>(a.ldr), 32, 32);
        if (rc) return rc;
      }
    }
  } else {
    EpiPool3::Params& e = p.pp;
    e.sims = static_cast<float*>(a.out);
    e.argmax = a.argmax;
    e.C = a.N / 3;
  }
  const long long mbc = ((a.M + GEMM_BM - 1) / GEMM_BM + p.cm - 1) / p.cm;
  const long long ctiles = mbc * ((a.N + p.bn - 1) / p.bn) * G * a.split_k;   // cluster tiles
  const long long max_clusters = num_sms() / p.cm;
  p.grid = static_cast<int>((ctiles < max_clusters ? ctiles : max_clusters) * p.cm);
  p.tiles = ctiles;
  return OWL_OK;
}

int gemm_plan_launch(const GemmPlan& p, cudaStream_t s) {
  if (!p.a_mn && !p.b_mn) return gemm_launch_kk(p, s);
  if (!p.a_mn && p.b_mn) return gemm_launch_kmn(p, s);
  if (p.a_mn && p.b_mn) return gemm_launch_mnmn(p, s);
  set_error("gemm: MN-major A with K-major B is not instantiated");
  return OWL_ERR_UNSUPPORTED;
}

}  // namespace owl

extern "C" int owl_gemm(const owl_gemm_args* args, void* stream) {
  if (!args) { owl::set_error("owl_gemm: null args"); return owl::OWL_ERR_ARG; }
  owl::GemmPlan plan;
  int rc = owl::gemm_plan_build(*args, args->bn, &plan);
  if (rc) return rc;
  return owl::gemm_plan_launch(plan, static_cast<cudaStream_t>(stream));
}
