"""GEMM micro-benchmark on the B/32 layer shapes at batch 16 (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops

M = 16 * 577
which = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bn = int(sys.argv[3]) if len(sys.argv) > 3 else 0


def t(fn, name, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s")


def r16(*s):
    return (torch.randn(*s, device="cuda") * 0.05).half()


x768, x3072 = r16(M, 768), r16(M, 3072)
w_qkv, w_o, w_1, w_2 = r16(2304, 768), r16(768, 768), r16(3072, 768), r16(768, 3072)
b768, b2304, b3072 = torch.randn(768, device="cuda"), torch.randn(2304, device="cuda"), torch.randn(3072, device="cuda")
o2304, o3072 = torch.empty(M, 2304, device="cuda", dtype=torch.half), torch.empty(M, 3072, device="cuda", dtype=torch.half)
res, o32 = torch.randn(M, 768, device="cuda"), torch.empty(M, 768, device="cuda")
if which in ("all", "qkv"):
    t(lambda: ops.gemm(x768, w_qkv, o2304, M=M, N=2304, K=768, bias=b2304, bn=bn), "qkv  f16 bias", 2.0 * M * 2304 * 768)
if which in ("all", "fc1"):
    t(lambda: ops.gemm(x768, w_1, o3072, M=M, N=3072, K=768, bias=b3072, act="quick_gelu", bn=bn), "fc1  f16 bias qgelu", 2.0 * M * 3072 * 768)
    t(lambda: ops.gemm(x768, w_1, o3072, M=M, N=3072, K=768, bn=bn), "fc1  f16 plain", 2.0 * M * 3072 * 768)
if which in ("all", "fc2"):
    t(lambda: ops.gemm(x3072, w_2, o32, M=M, N=768, K=3072, bias=b768, resid=res, bn=bn), "fc2  f32 bias resid", 2.0 * M * 768 * 3072)
if which in ("all", "out"):
    t(lambda: ops.gemm(x768, w_o, o32, M=M, N=768, K=768, bias=b768, resid=res, bn=bn), "out  f32 bias resid", 2.0 * M * 768 * 768)
    t(lambda: ops.gemm(x768, w_o, o32, M=M, N=768, K=768, bn=bn), "out  f32 plain", 2.0 * M * 768 * 768)
