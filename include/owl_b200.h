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
/* ABI version of this header; bumped on any signature change. */
int owl_abi_version(void);

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
  int bn;                                   /* N tile: 0 = auto, else 64 / 128 / 192 / 256 */
  float alpha;

  /* epilogue */
  int epilogue;          /* 0 = fp16 out, 1 = fp32 out, 2 = max-pool-3 (class head) */
  void* out;             /* fp16 / fp32 [.., ldo] ; epilogue 2: float sims [M, N/3] */
  long long ldo;
  long long o_outer_stride, o_head_stride; /* elements */
  const float* bias;     /* [N] or NULL */
  int act;               /* 0 none, 1 quick_gelu, 2 gelu(erf), 3 *= quick_gelu'(act_src), 4 *= gelu'(act_src) */
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
} owl_gemm_args;

int owl_gemm(const owl_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OWL_B200_H */
