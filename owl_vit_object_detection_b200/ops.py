"""Thin Python wrappers over the C ABI (one function per `owl_*` entry point).

These take torch CUDA tensors only for their device pointers / strides; all arithmetic happens in
libowl_b200.so.  Used by the parity tests and by the reference-facing host classes in `src/`.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs, check, lib, ptr, stream_ptr

ACT = {"none": 0, "quick_gelu": 1, "gelu": 2, "quick_gelu_grad": 3, "gelu_grad": 4, "exp_row": 5, "softmax_grad": 6}


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         a_mn: bool = False, b_mn: bool = False, a_ld: Optional[int] = None, b_ld: Optional[int] = None,
         ldo: Optional[int] = None, batches_outer: int = 1, heads: int = 1,
         a_outer_stride: int = 0, a_head_stride: int = 0, b_outer_stride: int = 0, b_head_stride: int = 0,
         a_head_col: int = 0, b_head_col: int = 0, o_outer_stride: int = 0, o_head_stride: int = 0,
         split_k: int = 1, bn: int = 0, alpha: float = 1.0, bias: Optional[torch.Tensor] = None,
         act: str = "none", pre_out: Optional[torch.Tensor] = None, act_src: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None, rows_per_img: int = 0,
         out_mode: int = 0, argmax: Optional[torch.Tensor] = None, pool3: bool = False,
         alpha_dev: Optional[torch.Tensor] = None, cluster_m: int = 0, rowvec: Optional[torch.Tensor] = None,
         rowvec_stride: int = 0, act_src_outer_stride: int = 0, act_src_head_stride: int = 0) -> torch.Tensor:
    """D = alpha * A @ B^T with a fused epilogue; see `struct owl_gemm_args` in include/owl_b200.h."""
    assert a.dtype == torch.float16 and b.dtype == torch.float16 and a.is_cuda and b.is_cuda
    g = GemmArgs()
    g.a, g.b = a.data_ptr(), b.data_ptr()
    g.a_mn, g.b_mn = int(a_mn), int(b_mn)
    g.M, g.N, g.K = M, N, K
    g.a_ld = a_ld if a_ld is not None else a.stride(-2)
    g.b_ld = b_ld if b_ld is not None else b.stride(-2)
    g.batches_outer, g.heads = batches_outer, heads
    g.a_outer_stride, g.a_head_stride = a_outer_stride, a_head_stride
    g.b_outer_stride, g.b_head_stride = b_outer_stride, b_head_stride
    g.a_head_col, g.b_head_col = a_head_col, b_head_col
    g.split_k, g.bn, g.alpha = split_k, bn, alpha
    if pool3:
        g.epilogue = 2
    elif out.dtype == torch.float16:
        g.epilogue = 0
    elif out.dtype == torch.float32:
        g.epilogue = 1
    else:
        raise TypeError(out.dtype)
    g.out = out.data_ptr()
    g.ldo = ldo if ldo is not None else out.stride(-2)
    g.o_outer_stride, g.o_head_stride = o_outer_stride, o_head_stride
    g.bias = ptr(bias)
    g.act = ACT[act]
    g.pre_out = ptr(pre_out)
    g.ld_pre = pre_out.stride(-2) if pre_out is not None else 0
    g.act_src = ptr(act_src)
    g.ld_act_src = act_src.stride(-2) if act_src is not None else 0
    g.resid = ptr(resid)
    g.ldr = resid.stride(-2) if resid is not None else 0
    g.pos = ptr(pos)
    g.rows_per_img = rows_per_img
    g.out_mode = out_mode
    g.argmax = ptr(argmax)
    g.alpha_dev = ptr(alpha_dev)
    g.cluster_m = cluster_m
    g.rowvec = ptr(rowvec)
    g.rowvec_stride = rowvec_stride
    g.act_src_outer_stride, g.act_src_head_stride = act_src_outer_stride, act_src_head_stride
    check(lib().owl_gemm(ctypes.byref(g), ctypes.c_void_p(stream_ptr())), "owl_gemm")
    return out


def _vp(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(ptr(t))


def _sp() -> ctypes.c_void_p:
    return ctypes.c_void_p(stream_ptr())


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t


def im2col_f16(img: torch.Tensor, patches: torch.Tensor, patch: int) -> torch.Tensor:
    """HF:336 patch gather: img [B,3,IS,IS] f32 -> patches [B*g*g, ld] f16 (owl_im2col_f16)."""
    B, _, IS, _ = img.shape
    _f32(img)
    assert patches.dtype == torch.float16 and patches.is_cuda
    check(lib().owl_im2col_f16(_vp(img), _vp(patches), B, IS, patch, ctypes.c_longlong(patches.stride(0)), _sp()),
          "owl_im2col_f16")
    return patches


def u8_patches_f16(img_u8: torch.Tensor, lut: torch.Tensor, patches: torch.Tensor, patch: int) -> torch.Tensor:
    """Raw RGB uint8 [B,IS,IS,3] at the model's resolution -> normalised fp16 patch rows (owl_u8_patches_f16)."""
    assert img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[3] == 3
    assert img_u8.is_contiguous() and img_u8.shape[1] == img_u8.shape[2]
    _f32(lut)
    assert lut.numel() == 768 and patches.dtype == torch.float16 and patches.is_cuda
    B, IS = int(img_u8.shape[0]), int(img_u8.shape[1])
    check(lib().owl_u8_patches_f16(_vp(img_u8), _vp(lut), _vp(patches), B, IS, patch,
                                   ctypes.c_longlong(patches.stride(0)), _sp()), "owl_u8_patches_f16")
    return patches


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, y: torch.Tensor, *, rows: int, D: int,
              eps: float, x_stride: Optional[int] = None, y_stride: Optional[int] = None,
              cls_emb: Optional[torch.Tensor] = None, pos0: Optional[torch.Tensor] = None,
              tokens: int = 0) -> torch.Tensor:
    assert x.dtype == torch.float32 and y.dtype in (torch.float16, torch.float32)
    xs = D if x_stride is None else x_stride
    ys = D if y_stride is None else y_stride
    check(lib().owl_layernorm(_vp(x), ctypes.c_longlong(xs), _vp(gamma), _vp(beta), _vp(y), ctypes.c_longlong(ys),
                              int(y.dtype == torch.float16), rows, D, ctypes.c_float(eps), _vp(cls_emb), _vp(pos0),
                              tokens, _sp()), "owl_layernorm")
    return y


def post_fuse(x, ecls, g1, b1, g2, b2, feats, *, B: int, P: int, D: int, eps: float):
    check(lib().owl_post_fuse(_vp(x), _vp(ecls), _vp(g1), _vp(b1), _vp(g2), _vp(b2), _vp(feats), B, P, D,
                              ctypes.c_float(eps), _sp()), "owl_post_fuse")
    return feats


def rownorm_f16(e: torch.Tensor, out: torch.Tensor, *, rows: int, E: int, query_mode: bool):
    _f32(e)
    assert out.dtype == torch.float16
    check(lib().owl_rownorm_f16(_vp(e), _vp(out), rows, E, int(query_mode), _sp()), "owl_rownorm_f16")
    return out


def box_tail(h16, w, bias, box_bias, boxes, sig, *, M: int, P: int, D: int):
    check(lib().owl_box_tail(_vp(h16), _vp(w), _vp(bias), _vp(box_bias), _vp(boxes), _vp(sig), M, P, D, _sp()),
          "owl_box_tail")
    return boxes


_L2_WINDOW = (0, 0)


def l2_persist(t: Optional[torch.Tensor], hit_ratio: float = 1.0) -> None:
    """Tag `t`'s storage as L2-persisting for every later launch (owl_l2_persist); None clears the window.
    Opt-in with OWL_L2_PERSIST=1: on B200 the step time with and without the window is the same within the
    run-to-run noise (3.594 / 3.597 ms with, 3.602 / 3.575 ms without, same box, 30 steps each)."""
    global _L2_WINDOW
    import os
    if os.environ.get("OWL_L2_PERSIST", "0") != "1":
        return
    key = (0, 0) if t is None else (t.data_ptr(), t.numel() * t.element_size())
    if key == _L2_WINDOW:
        return
    check(lib().owl_l2_persist(ctypes.c_void_p(key[0]), ctypes.c_longlong(key[1]), ctypes.c_float(hit_ratio)),
          "owl_l2_persist")
    _L2_WINDOW = key


def flash_attn_fwd(qkv16: torch.Tensor, ctx16: torch.Tensor, *, B: int, S: int, H: int, head_dim: int, scale: float,
                   lse: Optional[torch.Tensor] = None):
    """ctx = softmax(scale q k^T) v per (image, head); optionally lse [B, H, S] fp32 (natural log) for the backward."""
    assert qkv16.dtype == torch.float16 and ctx16.dtype == torch.float16 and qkv16.is_contiguous()
    if lse is not None:
        _f32(lse)
        assert lse.numel() == B * H * S
    check(lib().owl_flash_attn_fwd(_vp(qkv16), _vp(ctx16), _vp(lse), B, S, H, head_dim, ctypes.c_float(scale), _sp()),
          "owl_flash_attn_fwd")
    return ctx16


def attn_delta(ctx16: torch.Tensor, dctx16: torch.Tensor, delta: torch.Tensor, *, B: int, S: int, H: int,
               head_dim: int, alpha: float):
    """delta[b, h, s] = alpha * sum_d dctx[b, s, h, d] * ctx[b, s, h, d] (softmax backward row term)."""
    assert ctx16.dtype == torch.float16 and dctx16.dtype == torch.float16
    assert ctx16.is_contiguous() and dctx16.is_contiguous()
    _f32(delta)
    assert delta.numel() == B * H * S
    check(lib().owl_attn_delta(_vp(ctx16), _vp(dctx16), _vp(delta), B, S, H, head_dim, ctypes.c_float(alpha), _sp()),
          "owl_attn_delta")
    return delta


def attn_bwd(qkv16: torch.Tensor, dctx16: torch.Tensor, lse: torch.Tensor, delta: torch.Tensor, dqkv16: torch.Tensor,
             dq32: torch.Tensor, *, B: int, S: int, H: int, head_dim: int, scale: float):
    """Fused attention backward: dqkv (fp16, packed like qkv) from qkv, dctx, lse and delta; dq32 is fp32 scratch."""
    assert qkv16.dtype == torch.float16 and dctx16.dtype == torch.float16 and dqkv16.dtype == torch.float16
    assert qkv16.is_contiguous() and dctx16.is_contiguous() and dqkv16.is_contiguous()
    _f32(lse), _f32(delta), _f32(dq32)
    D = H * head_dim
    assert qkv16.numel() == B * S * 3 * D and dqkv16.numel() == B * S * 3 * D and dq32.numel() == B * S * D
    check(lib().owl_attn_bwd(_vp(qkv16), _vp(dctx16), _vp(lse), _vp(delta), _vp(dqkv16), _vp(dq32), B, S, H, head_dim,
                             ctypes.c_float(scale), _sp()), "owl_attn_bwd", kernels=2)
    return dqkv16


def zero(t: torch.Tensor) -> torch.Tensor:
    """t[...] = 0 through owl_zero (cudaMemsetAsync on the current stream; no framework kernel)."""
    assert t.is_cuda and t.is_contiguous()
    with torch.cuda.device(t.device):
        check(lib().owl_zero(_vp(t), _ll(t.numel() * t.element_size()), _sp()), "owl_zero", kernels=0)
    return t


def cast_f16(src: torch.Tensor, dst: torch.Tensor, scale: float = 1.0):
    _f32(src)
    assert dst.dtype == torch.float16 and dst.numel() == src.numel()
    check(lib().owl_cast_f16(_vp(src), _vp(dst), ctypes.c_longlong(src.numel()), ctypes.c_float(scale), _sp()),
          "owl_cast_f16")
    return dst


# ------------------------------------------------------------------------------------------ matcher + loss
def matcher_cost(sims, boxes, labels, tboxes, num_targets, costT, status, cost_class: float = 1.0,
                 cost_bbox: float = 1.0, cost_giou: float = 1.0):
    """reference src/matcher.py:103-131 -> costT [B,Tmax,P] (owl_matcher_cost)."""
    B, P, C = sims.shape
    Tmax = labels.shape[1]
    _f32(sims), _f32(boxes), _f32(tboxes), _f32(costT)
    assert labels.dtype == torch.int64 and num_targets.dtype == torch.int32 and status.dtype == torch.int32
    check(lib().owl_matcher_cost(_vp(sims), _vp(boxes), _vp(labels), _vp(tboxes), _vp(num_targets), _vp(costT), B, P,
                                 C, Tmax, _vp(status), ctypes.c_float(cost_class), ctypes.c_float(cost_bbox),
                                 ctypes.c_float(cost_giou), _sp()), "owl_matcher_cost")
    return costT


def lsap(costT, num_targets, match_pred, status):
    """reference src/matcher.py:135-137 (SciPy LSAP) -> match_pred [B,Tmax] i32 (owl_lsap)."""
    B, Tmax, P = costT.shape
    assert match_pred.dtype == torch.int32
    check(lib().owl_lsap(_vp(costT), _vp(num_targets), B, P, Tmax, _vp(match_pred), _vp(status), _sp()), "owl_lsap")
    return match_pred


LOSS_WS = 64   # floats of per-image workspace owl_match_loss needs ([:, :4] = ce, bg, bbox, giou on return)


def match_loss(sims, boxes, labels, tboxes, num_targets, match_pred, scales, bg_label, *, tc_matched, tc_final,
               pred_sorted, tgt_sorted, losses_per_image, losses_mean4, dsims_unit, dl1, dgiou):
    B, P, C = sims.shape
    assert losses_per_image.is_contiguous() and losses_per_image.numel() >= B * LOSS_WS
    Tmax = labels.shape[1]
    with torch.cuda.device(sims.device):
        check(lib().owl_match_loss(_vp(sims), _vp(boxes), _vp(labels), _vp(tboxes), _vp(num_targets), _vp(match_pred),
                                   _vp(scales), B, P, C, Tmax, bg_label, _vp(tc_matched), _vp(tc_final),
                                   _vp(pred_sorted), _vp(tgt_sorted), _vp(losses_per_image), _vp(losses_mean4),
                                   _vp(dsims_unit), _vp(dl1), _vp(dgiou), _sp()), "owl_match_loss", kernels=3)


def loss_backward(dsims_unit, tc_final, match_pred, dl1, dgiou, upstream4, bg_label, dsims, dboxes):
    B, P, C = dsims_unit.shape
    Tmax = match_pred.shape[1]
    with torch.cuda.device(dsims_unit.device):
        check(lib().owl_loss_backward(_vp(dsims_unit), _vp(tc_final), _vp(match_pred), _vp(dl1), _vp(dgiou),
                                      _vp(upstream4), B, P, C, Tmax, bg_label, _vp(dsims), _vp(dboxes), _sp()),
              "owl_loss_backward")


def preprocess_workspace_bytes(H: int, W: int, out_size: int) -> int:
    fn = lib().owl_preprocess_workspace_bytes
    fn.restype = ctypes.c_longlong
    return int(fn(H, W, out_size))


def preprocess_image(img_hwc: torch.Tensor, lut: torch.Tensor, out_chw: torch.Tensor, workspace: torch.Tensor):
    """uint8 [H,W,3] CUDA (rows may be strided) -> out_chw [3,S,S] fp32: Pillow-exact bicubic resize + the lut."""
    assert img_hwc.is_cuda and img_hwc.dtype == torch.uint8 and img_hwc.dim() == 3 and img_hwc.shape[2] == 3
    assert img_hwc.stride(2) == 1 and img_hwc.stride(1) == 3, "pixels must be packed RGB"
    _f32(lut), _f32(out_chw)
    assert lut.numel() == 768 and out_chw.dim() == 3 and out_chw.shape[0] == 3 and out_chw.shape[1] == out_chw.shape[2]
    assert workspace.dtype == torch.uint8 and workspace.is_cuda
    H, W = int(img_hwc.shape[0]), int(img_hwc.shape[1])
    check(lib().owl_preprocess_image(_vp(img_hwc), H, W, ctypes.c_longlong(img_hwc.stride(0)), _vp(lut), _vp(out_chw),
                                     int(out_chw.shape[1]), _vp(workspace), ctypes.c_longlong(workspace.numel()), _sp()),
          "owl_preprocess_image", kernels=3)
    return out_chw


class PreImage(ctypes.Structure):
    """Mirror of `struct owl_pre_image` (include/owl_b200.h)."""
    _fields_ = [("pixels", ctypes.c_void_p), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("row_stride_bytes", ctypes.c_longlong)]


def _pre_descriptors(images):
    arr = (PreImage * len(images))()
    for i, im in enumerate(images):
        assert im.is_cuda and im.dtype == torch.uint8 and im.dim() == 3 and im.shape[2] == 3, (im.dtype, im.shape)
        assert im.stride(2) == 1 and im.stride(1) == 3, "pixels must be packed RGB"
        arr[i].pixels, arr[i].H, arr[i].W = im.data_ptr(), int(im.shape[0]), int(im.shape[1])
        arr[i].row_stride_bytes = int(im.stride(0))
    return arr


def preprocess_batch_workspace_bytes(images, out_size: int) -> int:
    fn = lib().owl_preprocess_batch_workspace_bytes
    fn.restype = ctypes.c_longlong
    return int(fn(_pre_descriptors(images), len(images), out_size))


def preprocess_batch(images, lut: torch.Tensor, out: torch.Tensor, workspace: torch.Tensor) -> torch.Tensor:
    """uint8 [H_i, W_i, 3] CUDA images of different sizes -> out [n, 3, S, S] fp32 in three launches per 32 images
    (owl_preprocess_batch: Pillow-exact bicubic resize + the rescale / normalise lut)."""
    n = len(images)
    _f32(lut), _f32(out)
    assert lut.numel() == 768 and out.dim() == 4 and out.shape[0] == n and out.shape[1] == 3
    assert out.shape[2] == out.shape[3] and workspace.dtype == torch.uint8 and workspace.is_cuda
    check(lib().owl_preprocess_batch(_pre_descriptors(images), n, _vp(lut), _vp(out), int(out.shape[2]), _vp(workspace),
                                     ctypes.c_longlong(workspace.numel()), _sp()),
          "owl_preprocess_batch", kernels=3 * ((n + 31) // 32))
    return out


def postprocess(boxes: torch.Tensor, sims: torch.Tensor, confidence_threshold: float, iou_threshold: float):
    """reference src/models.py:122-146 for a whole batch on the device: returns (out_boxes [B,P,4], out_classes
    [B,P] i64, out_scores [B,P], count [B] i32); the first count[b] rows of image b are its detections."""
    _f32(boxes), _f32(sims)
    B, P, C = sims.shape
    assert boxes.shape == (B, P, 4)
    dev = sims.device
    out_boxes = torch.zeros((B, P, 4), dtype=torch.float32, device=dev)
    out_classes = torch.zeros((B, P), dtype=torch.int64, device=dev)
    out_scores = torch.zeros((B, P), dtype=torch.float32, device=dev)
    count = torch.zeros((B,), dtype=torch.int32, device=dev)
    check(lib().owl_postprocess(_vp(boxes), _vp(sims), B, P, C, ctypes.c_float(confidence_threshold),
                                ctypes.c_double(iou_threshold), _vp(out_boxes), _vp(out_classes), _vp(out_scores),
                                _vp(count), _sp()), "owl_postprocess")
    return out_boxes, out_classes, out_scores, count


# ------------------------------------------------------------------------------------------ backward kernels
def _ll(v: int) -> ctypes.c_longlong:
    return ctypes.c_longlong(v)


def _fl(v: float) -> ctypes.c_float:
    return ctypes.c_float(v)


def grad_scale(a: torch.Tensor, b: Optional[torch.Tensor], gscale: torch.Tensor, target: float = 64.0):
    assert gscale.dtype == torch.float32 and gscale.numel() >= 4
    check(lib().owl_grad_scale(_vp(a), _ll(a.numel()), _vp(b), _ll(0 if b is None else b.numel()), _fl(target),
                               _vp(gscale), _sp()), "owl_grad_scale", kernels=2)
    return gscale


def pool3_bwd(dsims, argmax, gscale, dfull16):
    check(lib().owl_pool3_bwd(_vp(dsims), _vp(argmax), _vp(gscale), _vp(dfull16), _ll(dsims.numel()), _sp()),
          "owl_pool3_bwd")
    return dfull16


def rownorm_bwd(e, dy, out, *, rows: int, E: int, query_mode: bool, gscale):
    assert out.dtype == (torch.float32 if query_mode else torch.float16)
    check(lib().owl_rownorm_bwd(_vp(e), _vp(dy), _vp(out), rows, E, int(query_mode), _vp(gscale), _sp()),
          "owl_rownorm_bwd")
    return out


def box_tail_bwd(dboxes, sig, w2, pre1_16, h1_16, gscale, dz, dpre1_16, dw2, db2, *, M: int, D: int):
    check(lib().owl_box_tail_bwd(_vp(dboxes), _vp(sig), _vp(w2), _vp(pre1_16), _vp(h1_16), _vp(gscale), _vp(dz),
                                 _vp(dpre1_16), _vp(dw2), _vp(db2), M, D, _sp()), "owl_box_tail_bwd", kernels=2)


def colsum(x: torch.Tensor, out: torch.Tensor, *, M: int, N: int, gscale=None, ld: Optional[int] = None,
           cast_to: Optional[torch.Tensor] = None):
    """out[n] += unscale * sum_m x[m, n]; `cast_to` (fp32 x only) also receives x as fp16 [M, N] in the same pass."""
    assert x.dtype in (torch.float16, torch.float32) and out.dtype == torch.float32
    if cast_to is not None:
        assert x.dtype == torch.float32 and cast_to.dtype == torch.float16 and cast_to.is_contiguous()
        assert cast_to.numel() == M * N
    check(lib().owl_colsum(_vp(x), int(x.dtype == torch.float16), _ll(x.stride(-2) if ld is None else ld), M, N,
                           _vp(gscale), _vp(out), _vp(cast_to), _sp()), "owl_colsum")
    return out


def layernorm_bwd(x, dy, gamma, dgamma, dbeta, *, rows: int, D: int, eps: float, gscale=None, dx=None, dx_add=None,
                  x_stride: Optional[int] = None, dy_stride: Optional[int] = None, dx_stride: Optional[int] = None):
    check(lib().owl_layernorm_bwd(_vp(x), _ll(D if x_stride is None else x_stride), _vp(dy),
                                  _ll(D if dy_stride is None else dy_stride), _vp(gamma), _vp(dx_add), _vp(dx),
                                  _ll(D if dx_stride is None else dx_stride), _vp(dgamma), _vp(dbeta), rows, D,
                                  _fl(eps), _vp(gscale), _sp()), "owl_layernorm_bwd")


def post_fuse_bwd(x, ecls, g1, b1, g2, dfeats, dx, dcl, dg1, db1, dg2, db2, *, B: int, P: int, D: int, eps: float,
                  gscale):
    check(lib().owl_post_fuse_bwd(_vp(x), _vp(ecls), _vp(g1), _vp(b1), _vp(g2), _vp(dfeats), _vp(dx), _vp(dcl),
                                  _vp(dg1), _vp(db1), _vp(dg2), _vp(db2), B, P, D, _fl(eps), _vp(gscale), _sp()),
          "owl_post_fuse_bwd")


def adamw(params, grads, exp_avg, exp_avg_sq, params16, *, lr: float, beta1: float, beta2: float, eps: float,
          weight_decay: float, state: torch.Tensor, grad_mul: float = 1.0):
    n = params.numel()
    assert grads.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    check(lib().owl_adamw(_vp(params), _vp(grads), _vp(exp_avg), _vp(exp_avg_sq), _vp(params16), _ll(n), _fl(lr),
                          _fl(beta1), _fl(beta2), _fl(eps), _fl(weight_decay), _vp(state), _fl(grad_mul), _sp()),
          "owl_adamw", kernels=2)


def allreduce_multimem(multicast_ptr: int, n: int, rank: int, world: int):
    """In-switch all-reduce (sum) of a symmetric fp32 buffer addressed through its multicast pointer
    (owl_allreduce_multimem); the caller issues the cross-rank barriers around it."""
    check(lib().owl_allreduce_multimem(ctypes.c_void_p(multicast_ptr), _ll(n), rank, world, _sp()), "owl_allreduce_multimem")


# ------------------------------------------------------------------------------------------ text tower (row N4)
def text_embed(ids: torch.Tensor, tok_emb: torch.Tensor, pos_emb: torch.Tensor, x: torch.Tensor, *, S: int,
               status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """HF:370-373: x[n*S + s] = tok_emb[ids[n, s]] + pos_emb[s] (owl_text_embed)."""
    assert ids.is_cuda and ids.dtype == torch.int64 and ids.is_contiguous()
    _f32(tok_emb), _f32(pos_emb), _f32(x)
    rows, D = ids.numel(), tok_emb.shape[1]
    assert x.numel() == rows * D and pos_emb.shape[0] >= S and pos_emb.shape[1] == D
    check(lib().owl_text_embed(_vp(ids), _vp(tok_emb), _vp(pos_emb), _vp(x), rows, S, D, int(tok_emb.shape[0]),
                               _vp(status), _sp()), "owl_text_embed")
    return x


def text_attn(qkv16: torch.Tensor, mask: Optional[torch.Tensor], ctx16: torch.Tensor, *, N: int, S: int, H: int,
              head_dim: int, scale: float) -> torch.Tensor:
    """HF:379-404 with the causal + padding mask of the text model (owl_text_attn)."""
    assert qkv16.dtype == torch.float16 and ctx16.dtype == torch.float16 and qkv16.is_contiguous()
    assert qkv16.numel() == N * S * 3 * H * head_dim and ctx16.numel() == N * S * H * head_dim
    if mask is not None:
        assert mask.is_cuda and mask.dtype == torch.int32 and mask.is_contiguous() and mask.numel() == N * S
    check(lib().owl_text_attn(_vp(qkv16), _vp(mask), _vp(ctx16), N, S, H, head_dim, ctypes.c_float(scale), _sp()),
          "owl_text_attn")
    return ctx16


def text_pool_ln(x: torch.Tensor, ids: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out16: torch.Tensor, *,
                 N: int, S: int, D: int, eps: float) -> torch.Tensor:
    """HF:677-684: final LayerNorm of each prompt's end-of-text row -> fp16 [N, D] (owl_text_pool_ln)."""
    _f32(x), _f32(gamma), _f32(beta)
    assert ids.dtype == torch.int64 and ids.is_contiguous() and out16.dtype == torch.float16
    check(lib().owl_text_pool_ln(_vp(x), _vp(ids), _vp(gamma), _vp(beta), _vp(out16), N, S, D, ctypes.c_float(eps),
                                 _sp()), "owl_text_pool_ln")
    return out16


def l2norm_rows(src: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """HF:984: rows divided by their L2 norm (owl_l2norm_rows)."""
    _f32(src), _f32(out)
    rows, D = src.shape
    check(lib().owl_l2norm_rows(_vp(src), _vp(out), rows, D, _sp()), "owl_l2norm_rows")
    return out
