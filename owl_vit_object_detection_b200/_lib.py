"""ctypes binding of libowl_b200.so (the C ABI declared in include/owl_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libowl_b200.so")

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


class OwlError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    """Mirror of `struct owl_gemm_args` (include/owl_b200.h)."""
    _fields_ = [
        ("a", c_void_p), ("b", c_void_p),
        ("a_mn", c_int), ("b_mn", c_int),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("a_ld", c_ll), ("b_ld", c_ll),
        ("batches_outer", c_int), ("heads", c_int),
        ("a_outer_stride", c_ll), ("a_head_stride", c_ll),
        ("b_outer_stride", c_ll), ("b_head_stride", c_ll),
        ("a_head_col", c_int), ("b_head_col", c_int),
        ("split_k", c_int), ("bn", c_int), ("alpha", c_float),
        ("epilogue", c_int), ("out", c_void_p), ("ldo", c_ll),
        ("o_outer_stride", c_ll), ("o_head_stride", c_ll),
        ("bias", c_void_p), ("act", c_int),
        ("pre_out", c_void_p), ("ld_pre", c_ll),
        ("act_src", c_void_p), ("ld_act_src", c_ll),
        ("resid", c_void_p), ("ldr", c_ll),
        ("pos", c_void_p), ("rows_per_img", c_int), ("out_mode", c_int),
        ("argmax", c_void_p),
        ("alpha_dev", c_void_p),
        ("cluster_m", c_int),
        ("rowvec", c_void_p), ("rowvec_stride", c_ll),
        ("act_src_outer_stride", c_ll), ("act_src_head_stride", c_ll),
    ]


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OwlError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        L = ctypes.CDLL(LIB_PATH)
        L.owl_last_error.restype = ctypes.c_char_p
        L.owl_abi_version.restype = c_int
        _lib = L
    return _lib


KERNEL_LAUNCHES = 0   # kernels enqueued through the C ABI by this process (bench.py reports it)


def check(rc: int, what: str, kernels: int = 1) -> None:
    global KERNEL_LAUNCHES
    KERNEL_LAUNCHES += kernels
    if rc != 0:
        msg = lib().owl_last_error().decode(errors="replace")
        raise OwlError(f"{what} failed (code {rc}): {msg}")


def stream_ptr() -> int:
    """The current torch stream of the CURRENT device.  Kernels are launched on the current device: callers that hold
    tensors of another device enter `torch.cuda.device(t.device)` first (Engine / the loss classes do)."""
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
