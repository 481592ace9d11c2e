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

ACT = {"none": 0, "quick_gelu": 1, "gelu": 2, "quick_gelu_grad": 3, "gelu_grad": 4}


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         a_mn: bool = False, b_mn: bool = False, a_ld: Optional[int] = None, b_ld: Optional[int] = None,
         ldo: Optional[int] = None, batches_outer: int = 1, heads: int = 1,
         a_outer_stride: int = 0, a_head_stride: int = 0, b_outer_stride: int = 0, b_head_stride: int = 0,
         a_head_col: int = 0, b_head_col: int = 0, o_outer_stride: int = 0, o_head_stride: int = 0,
         split_k: int = 1, bn: int = 0, alpha: float = 1.0, bias: Optional[torch.Tensor] = None,
         act: str = "none", pre_out: Optional[torch.Tensor] = None, act_src: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None, rows_per_img: int = 0,
         out_mode: int = 0, argmax: Optional[torch.Tensor] = None, pool3: bool = False) -> torch.Tensor:
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
    check(lib().owl_gemm(ctypes.byref(g), ctypes.c_void_p(stream_ptr())), "owl_gemm")
    return out
