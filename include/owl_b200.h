/* owl_b200.h — C ABI of libowl_b200.so (hand-written sm_100a kernels for the OWL-ViT fine-tuning hot path).
 *
 * The reference (stevebottos/owl-vit-object-detection) has no FFI of its own: its hot path is Python
 * that calls third-party libraries.  Each entry point below names the reference
 * call site it replaces (file:line under /root/reference, or "HF:" = the HuggingFace OWL-ViT modeling
 * file the reference calls into).  INTEGRATION.md shows the ctypes stubs a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no allocation, no sync);
 *   - return value 0 = success, otherwise a negative owl error or a positive cudaError_t; the message is
 *     available from owl_last_error() (thread-local);
 *   - fp16 = IEEE binary16 storage, all accumulation in fp32; "f32" buffers are IEEE binary32.
 */
#ifndef OWL_B200_H
#define OWL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* owl_last_error(void);
/* Zero-fill `bytes` bytes at `ptr` on the stream (cudaMemsetAsync: a memset node inside a captured graph). */
int owl_zero(void* ptr, long long bytes, void* stream);
/* ABI version of this header; bumped on any signature change. */
int owl_abi_version(void);

/* L2 persistence window (no reference counterpart: the reference's residual stream is whatever ATen allocates).
 * Every later launch of this library tags accesses to [base, base + bytes) as persisting in the L2 set-aside
 * (cudaLaunchAttributeAccessPolicyWindow); the set-aside is sized to the window (capped by the device limits).
 * The engine can point it at the fp32 residual stream of the encoder (HF:490-511: read by both LayerNorms and both
 * residual adds of a layer; opt-in, measured neutral on B200).  base = NULL or bytes <= 0 clears it.  Best effort: where the device refuses the
 * set-aside the launches simply stay untagged (always returns 0).  Call it outside stream capture. */
int owl_l2_persist(const void* base, long long bytes, float hit_ratio);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 + TMEM + TMA).  Replaces every nn.Linear / Conv2d / matmul on the path:
 * HF:336 (patch embed), HF:439-441 (q,k,v), HF:459 (out_proj), HF:474-476 (MLP), HF:1020-1024 (box
 * head), reference src/models.py:25,35 (class head) and their autograd backward (reference main.py:90).
 *
 *   D[g][m][n] = alpha * sum_k A[g][m][k] * B[g][n][k]   then a fused epilogue.
 *
 * A and B are fp16 matrices in memory with `*_ld` elements between consecutive memory rows.
 *   *_mn == 0 : K-major  — memory row index = m (or n), memory column = k.
 *   *_mn == 1 : MN-major — memory row index = k,        memory column = m (or n)   (transposed read).
 * Batching: g = outer * heads + head; operand address = base + outer * *_outer_stride
 *           + head * *_head_stride (elements) + head * *_head_col (columns inside the row).
 */
typedef struct owl_gemm_args {
  const void* a;
  const void* b;
  int a_mn, b_mn;
  int M, N, K;
  long long a_ld, b_ld;
  int batches_outer, heads;                 /* both >= 1 */
  long long a_outer_stride, a_head_stride;  /* elements */
  long long b_outer_stride, b_head_stride;
  int a_head_col, b_head_col;               /* column offset per head */
  int split_k;                              /* >= 1; > 1 requires out_mode == 2 */
  int bn;                                   /* N tile: 0 = auto, else 64 / 128 / 256 */
  float alpha;

  /* epilogue */
  int epilogue;          /* 0 = fp16 out, 1 = fp32 out, 2 = max-pool-3 (class head) */
  void* out;             /* fp16 / fp32 [.., ldo] ; epilogue 2: float sims [M, N/3] */
  long long ldo;
  long long o_outer_stride, o_head_stride; /* elements */
  const float* bias;     /* [N] or NULL */
  int act;               /* 0 none, 1 quick_gelu, 2 gelu(erf), 3 *= quick_gelu'(act_src), 4 *= gelu'(act_src),
                            5 exp(v - rowvec[g][m]) (probabilities from a saved log-sum-exp, HF:398),
                            6 act_src[g][m][n] * (v - rowvec[g][m]) (softmax backward) */
  void* pre_out;         /* fp16, optional: value before the activation (saved for backward) */
  long long ld_pre;
  const void* act_src;   /* fp16 pre-activations for act 3/4 */
  long long ld_act_src;
  const float* resid;    /* epilogue 1: fp32 residual added, same row mapping as out */
  long long ldr;
  const float* pos;      /* epilogue 1: fp32 position table, row (m % rows_per_img) + 1, leading dim ldo */
  int rows_per_img;      /* epilogue 1: > 0 maps output row m -> m + m / rows_per_img + 1 */
  int out_mode;          /* epilogue 1: 0 store, 1 accumulate, 2 atomic add */
  uint8_t* argmax;       /* epilogue 2: winning prompt variant [M, N/3] */
  const float* alpha_dev; /* optional DEVICE scalar multiplied into alpha (gradient un-scaling without a host sync) */
  int cluster_m;         /* thread-block cluster along M with TMA multicast of the B tile: 0 = auto, 1 = off, 2 = pairs */
  const float* rowvec;   /* act 5 / 6: fp32 [batches_outer * heads][rowvec_stride], one value per output row */
  long long rowvec_stride;
  long long act_src_outer_stride, act_src_head_stride; /* elements; batch offsets of act_src (act 6) */
} owl_gemm_args;

int owl_gemm(const owl_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound forward kernels around the GEMMs (one warp per row, 128-bit accesses, fp32 statistics).
 */
/* HF:336 patch-embedding conv (kernel = stride = patch, no bias) as a gather: img [B,3,IS,IS] f32 NCHW ->
 * patches [B*(IS/patch)^2, ld] fp16 with column = c*patch^2 + ky*patch + kx (the conv weight's own order). */
int owl_im2col_f16(const float* img, void* patches, int B, int image_size, int patch, long long ld, void* stream);
/* reference src/dataset.py:64-71 (HF OwlViTImageProcessor: /255, CLIP mean / std) + HF:336 patch gather in ONE pass for
 * images that already have the model's resolution (PIL's resize is then the identity): img [B,IS,IS,3] uint8 RGB (HWC)
 * -> patches [B*(IS/patch)^2, ld] fp16, lut [3][256] fp32 = normalised value of byte v in channel c.  Bit-identical to
 * owl_preprocess_image + owl_im2col_f16 on such images; the host ships 1 byte per channel instead of 4. */
int owl_u8_patches_f16(const unsigned char* img, const float* lut, void* patches, int B, int image_size, int patch,
                       long long ld, void* stream);
/* HF:498,507,768 + reference src/models.py:80,86 LayerNorm over rows of length D (fp32 in; fp16 or fp32 out).
 * Row r is read at x + r*x_stride and written at y + r*y_stride.  When cls_emb != NULL, rows with
 * r %% tokens == 0 are replaced by cls_emb + pos0 first (the CLS row of the embedding, HF:338-343). */
int owl_layernorm(const float* x, long long x_stride, const float* gamma, const float* beta, void* y,
                  long long y_stride, int out_f16, int rows, int D, float eps, const float* cls_emb,
                  const float* pos0, int tokens, void* stream);
/* reference src/models.py:80-86 fused: feats[b,p] = LN2(LN1(x[b,1+p]) * ecls[b]), ecls[b] = LN1(x[b,0]). */
int owl_post_fuse(const float* x, const float* ecls, const float* g1, const float* b1, const float* g2,
                  const float* b2, void* feats_f16, int B, int P, int D, float eps, void* stream);
/* reference src/models.py:28-33: query_mode 0: e/(||e||+1e-6)   query_mode 1: q/||q|| + 1e-6.  fp16 out. */
int owl_rownorm_f16(const float* e, void* out_f16, int rows, int E, int query_mode, void* stream);
/* HF:1024 dense2 + reference src/models.py:71-73: boxes = corners(sigmoid(h W^T + b + box_bias)); also
 * stores the sigmoid output (cx,cy,w,h) for the backward pass. */
int owl_box_tail(const void* h_f16, const float* w, const float* bias, const float* box_bias, float* boxes,
                 float* sig, int M, int P, int D, void* stream);
int owl_cast_f16(const float* src, void* dst_f16, long long n, float scale, void* stream);
/* HF:379-404 fused attention forward: ctx[b, s, h*64 + d] = softmax(scale * q k^T) v for every (image, head), reading
 * the packed qkv buffer [B*S, 3*H*64] (q | k | v column blocks) and never materialising the scores (tcgen05: S and O
 * accumulate in TMEM, P is fed back to the tensor core from TMEM).  head_dim must be 64.  When lse != NULL it also
 * stores the natural-log log-sum-exp of the scaled scores, lse[b][h][s] (fp32), from which the backward pass
 * recomputes the probabilities (owl_gemm act 5). */
int owl_flash_attn_fwd(const void* qkv_f16, void* ctx_f16, float* lse, int B, int S, int H, int head_dim, float scale,
                       void* stream);
/* Softmax backward row term (autograd of HF:398): delta[b][h][s] = alpha * sum_d dctx[b,s,h*64+d] * ctx[b,s,h*64+d]
 * (= alpha * sum_j P dP), consumed by owl_gemm act 6. */
int owl_attn_delta(const void* ctx_f16, const void* dctx_f16, float* delta, int B, int S, int H, int head_dim,
                   float alpha, void* stream);
/* Fused attention backward (autograd of HF:393-404), head_dim 64: from the packed qkv [B*S, 3*H*64], dctx [B*S, H*64],
 * the forward's lse [B][H][S] and delta [B][H][S] (= scale * rowsum(dctx . ctx), owl_attn_delta) to dqkv [B*S, 3*H*64]
 * (fp16, same packing as qkv).  Scores, probabilities and their gradients stay in TMEM / shared memory; dq32
 * [B*S, H*64] fp32 is scratch (the per-key-block dQ contributions are accumulated there with 16-byte atomics). */
int owl_attn_bwd(const void* qkv_f16, const void* dctx_f16, const float* lse, const float* delta, void* dqkv_f16,
                 float* dq32, int B, int S, int H, int head_dim, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Image preprocessing (reference src/dataset.py:64-71 -> HF OwlViTImageProcessor of the pinned transformers 4.30.2:
 * PIL bicubic resize to out_size x out_size, rescale 1/255, CLIP mean / std, channels first), on raw uint8 pixels.
 * img_hwc [H][W][3] u8 RGB on the DEVICE (row_stride_bytes between rows), lut [3][256] f32 = value of channel c for
 * byte v (host-computed with the reference's op sequence), out_chw [3][out_size][out_size] f32.  The resample is
 * Pillow's ImagingResample bit for bit (integer coefficients with 22 fraction bits, horizontal pass then vertical
 * pass, each rounded to uint8).  workspace: owl_preprocess_workspace_bytes(H, W, out_size) bytes, 16-byte aligned. */
long long owl_preprocess_workspace_bytes(int H, int W, int out_size);
int owl_preprocess_image(const uint8_t* img_hwc, int H, int W, long long row_stride_bytes, const float* lut,
                         float* out_chw, int out_size, void* workspace, long long workspace_bytes, void* stream);
/* The same for a batch of n images of DIFFERENT sizes in three launches per 32 images (coefficients, horizontal pass,
 * vertical pass) instead of per image: images_host is a HOST array of descriptors (device pixel pointer, height,
 * width, bytes between rows), out_nchw [n][3][out_size][out_size] f32.  What a collate function for batch > 1 needs
 * (reference src/dataset.py:101-106 is batch-1).  workspace: owl_preprocess_batch_workspace_bytes(...) bytes. */
#define OWL_PRE_MAX_BATCH 32
typedef struct owl_pre_image {
  const uint8_t* pixels;
  int H, W;
  long long row_stride_bytes;
} owl_pre_image;
long long owl_preprocess_batch_workspace_bytes(const owl_pre_image* images_host, int n, int out_size);
int owl_preprocess_batch(const owl_pre_image* images_host, int n, const float* lut, float* out_nchw, int out_size,
                         void* workspace, long long workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Detection post-processing (reference src/models.py:122-146 `PostProcess`, eval path main.py:110-118): per image,
 * best class per prediction (first maximum), score > confidence_threshold, class-aware NMS exactly as
 * torchvision.ops.batched_nms's coordinate-trick path on the CPU (stable descending score order, fp32 IoU compared
 * with a double threshold).  boxes [B,P,4] f32 xyxy, sims [B,P,C] f32 -> the survivors of image b, in decreasing
 * score order, in out_boxes[b][0..count) / out_classes[b][..] (i64) / out_scores[b][..]; out_count [B] i32.
 * One CTA per image; the reference is batch-1 only. */
int owl_postprocess(const float* boxes, const float* sims, int B, int P, int C, float confidence_threshold,
                    double iou_threshold, float* out_boxes, long long* out_classes, float* out_scores, int* out_count,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Matcher + loss (reference src/matcher.py:85-159, src/losses.py:16-116), device-resident.
 * Targets are padded: labels [B,Tmax] i64, tboxes [B,Tmax,4] f32 xyxy, num_targets [B] i32.
 * `status` is a device int, OR-ed with 1 for a degenerate box (the reference asserts, src/matcher.py:34-35)
 * and 2 for an infeasible assignment; the caller zeroes it and checks it when it syncs anyway.
 */
/* src/matcher.py:103-131.  costT [B,Tmax,P] f32 (target-major so that the solver's scans are coalesced):
 * costT[b][t][p] = (cost_bbox * L1(box_p, tbox_t) + cost_class * -softmax(sims[b,p])[label_t]) + cost_giou * -GIoU(box_p,
 * tbox_t), the weights of HungarianMatcher.__init__ (src/matcher.py:55-60; the reference uses 1, 1, 1).  status bits:
 * 1 degenerate box, 2 infeasible, 4 num_targets outside [0, Tmax], 8 label outside [0, C). */
int owl_matcher_cost(const float* sims, const float* boxes, const long long* labels, const float* tboxes,
                     const int* num_targets, float* costT, int B, int P, int C, int Tmax, int* status,
                     float cost_class, float cost_bbox, float cost_giou, void* stream);
/* src/matcher.py:135-137 (scipy.optimize.linear_sum_assignment).  match_pred [B,Tmax] i32: prediction assigned
 * to each target, -1 for padding.  Index-exact with SciPy including its tie rule whenever num_targets < P
 * (SciPy transposes the problem only when rows > cols; the reference always has P = 576 > T). */
int owl_lsap(const float* costT, const int* num_targets, int B, int P, int Tmax, int* match_pred, int* status,
             void* stream);
/* src/matcher.py:138-159 + src/losses.py:42-69,100-108,16-40.
 *   tc_matched [B,P] i64  target_classes as the matcher returns them (background = bg_label)
 *   tc_final   [B,P] i64  after the IoU>0.85 ordered label sweep (src/losses.py:100-106)
 *   pred_sorted / tgt_sorted [B,Tmax] i64  the matcher's `indices` (sorted by prediction), -1 padded
 *   losses_per_image [B,64] (per-image workspace; [b][0..3] = loss_ce, loss_bg, loss_bbox, loss_giou on return),
 *   losses_mean4 [4]  (the same four, mean over images)
 *   dsims_unit [B,P,C], dl1 / dgiou [B,Tmax,4]  gradients for unit upstream grads, incl. the 1/B of the mean */
int owl_match_loss(const float* sims, const float* boxes, const long long* labels, const float* tboxes,
                   const int* num_targets, const int* match_pred, const float* scales, int B, int P, int C,
                   int Tmax, int bg_label, long long* tc_matched, long long* tc_final, long long* pred_sorted,
                   long long* tgt_sorted, float* losses_per_image, float* losses_mean4, float* dsims_unit,
                   float* dl1, float* dgiou, void* stream);
/* autograd of the four losses: upstream4 = d(total)/d(loss_ce, loss_bg, loss_bbox, loss_giou) on the device. */
int owl_loss_backward(const float* dsims_unit, const long long* tc_final, const int* match_pred, const float* dl1,
                      const float* dgiou, const float* upstream4, int B, int P, int C, int Tmax, int bg_label,
                      float* dsims, float* dboxes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward of the trainable part (autograd of reference main.py:90 under the freeze rule of reference
 * src/models.py:173-184) and the optimizer step (reference main.py:56-60,91).
 * Activation gradients are fp16 GEMM operands, pre-multiplied by a power of two S picked on the device:
 * gscale = {S, 1/S, scratch, scratch} (4 floats).  Kernels that write PARAMETER gradients multiply by 1/S
 * and ACCUMULATE (atomicAdd) into the flat fp32 gradient buffer, which the caller zeroes once per step.
 */
/* S = 2^floor(log2(target / max(|a|, |b|))) clamped to 2^+-24 (1 when the maximum is 0 / not finite). */
int owl_grad_scale(const float* a, long long na, const float* b, long long nb, float target, float* gscale,
                   void* stream);
/* reference src/models.py:36 MaxPool1d(3) backward: dfull [n,3] fp16 = S * dsims routed to the winning variant. */
int owl_pool3_bwd(const float* dsims, const uint8_t* argmax, const float* gscale, void* dfull_f16, long long n,
                  void* stream);
/* reference src/models.py:28-33 backward.  query_mode 0: fp16 out (still scaled); 1: fp32 out += value / S. */
int owl_rownorm_bwd(const float* e, const float* dy, void* out, int rows, int E, int query_mode,
                    const float* gscale, void* stream);
/* reference src/models.py:71-73 + HF:1024 backward: dz [M,4] = S * d(boxes)/d(logits); dpre1 [M,D] fp16 =
 * (dz W2) * gelu'(pre1); dw2 [4,D] += dz^T h1 / S; db2 [4] += sum dz / S. */
int owl_box_tail_bwd(const float* dboxes, const float* sig, const float* w2, const void* pre1_f16,
                     const void* h1_f16, const float* gscale, float* dz, void* dpre1_f16, float* dw2, float* db2,
                     int M, int D, void* stream);
/* bias gradients: out[n] += (gscale ? gscale[1] : 1) * sum_m x[m, n]   (x fp16 or fp32, N and ld even).
 * cast_out_f16 (optional, fp32 input, N and ld multiples of 4): also writes x as fp16 [M, N] in the same pass - the
 * operand of the next backward GEMMs (one read of the gradient instead of a cast pass plus a column-sum pass). */
int owl_colsum(const void* x, int is_f16, long long ld, int M, int N, const float* gscale, float* out,
               void* cast_out_f16, void* stream);
/* LayerNorm backward (HF:498,507; reference src/models.py:80): dx = dx_add + LN'(dy) (dx may be NULL when only the
 * parameter gradients are needed); dgamma / dbeta += 1/S * sums. */
int owl_layernorm_bwd(const float* x, long long x_stride, const float* dy, long long dy_stride, const float* gamma,
                      const float* dx_add, float* dx, long long dx_stride, float* dgamma, float* dbeta, int rows,
                      int D, float eps, const float* gscale, void* stream);
/* backward of owl_post_fuse: dx rows of the patch tokens, dcl [B,D] += (scaled), LayerNorm parameter grads += . */
int owl_post_fuse_bwd(const float* x, const float* ecls, const float* g1, const float* b1, const float* g2,
                      const float* dfeats, float* dx, float* dcl, float* dg1, float* db1, float* dg2, float* db2,
                      int B, int P, int D, float eps, const float* gscale, void* stream);
/* torch.optim.AdamW step (reference main.py:56-60,91) over a flat range; also refreshes the fp16 GEMM shadow.
 * state = 4 device floats {step count, 1 - beta1^step, sqrt(1 - beta2^step), unused}, zero-initialised by the
 * caller and advanced on the device (so a captured CUDA graph replays correctly).  grad_mul folds the
 * 1/world_size of the gradient all-reduce. */
int owl_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_f16, long long n,
              float lr, float beta1, float beta2, float eps, float weight_decay, float* state, float grad_mul,
              void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange (the reference has no distributed code).  Two-shot all-reduce (sum) of a
 * fp32 buffer that lives in symmetric memory, through the NVSwitch: rank r reduces slice r with multimem.ld_reduce on
 * the buffer's MULTICAST address and broadcasts the result with multimem.st.  The caller brackets the call with
 * cross-rank barriers on the same stream (every rank's buffer complete before, every slice landed after). */
int owl_allreduce_multimem(float* multicast_ptr, long long n, int rank, int world, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Query-bank initialisation: reference src/models.py:155-169 runs the HF TEXT tower once over three
 * prompts per class and keeps `text_embeds` as the learned queries.  The encoder layers reuse owl_layernorm / owl_gemm
 * (HF:490-511); these four entry points are the text-only pieces.  ids = int64 token ids [N prompts, S tokens]. */
/* HF:370-373: x[n*S + s, :] = token_embedding[ids[n, s]] + position_embedding[s] (fp32).  An id outside [0, vocab)
 * (IndexError in nn.Embedding) sets bit 8 of *status (may be NULL) and reads row 0. */
int owl_text_embed(const long long* ids, const float* tok_emb, const float* pos_emb, float* x, int rows, int S, int D,
                   int vocab, int* status, void* stream);
/* HF:379-404 under the mask of HF:661-666: packed q|k|v fp16 [N*S, 3*H*64] -> ctx fp16 [N*S, H*64]; query s sees key j
 * iff j <= s (causal) and mask[n, j] != 0 (attention_mask as int32, NULL = no padding).  S <= 32, head_dim 64. */
int owl_text_attn(const void* qkv_f16, const int* mask, void* ctx_f16, int N, int S, int H, int head_dim, float scale,
                  void* stream);
/* HF:677-684: final LayerNorm of the end-of-text row of each prompt (first argmax of its ids) -> fp16 [N, D]. */
int owl_text_pool_ln(const float* x, const long long* ids, const float* gamma, const float* beta, void* out_f16, int N,
                     int S, int D, float eps, void* stream);
/* HF:984: out[r, :] = in[r, :] / ||in[r, :]||_2 (fp32; in == out allowed). */
int owl_l2norm_rows(const float* in, float* out, int rows, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OWL_B200_H */
