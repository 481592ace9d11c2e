"""GEMM micro-benchmark on the B/32 layer shapes at batch 16 (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops

M = 16 * 577
which = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cm = int(sys.argv[4]) if len(sys.argv) > 4 else 0


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
    t(lambda: ops.gemm(x768, w_qkv, o2304, M=M, N=2304, K=768, bias=b2304, bn=bn, cluster_m=cm), "qkv  f16 bias", 2.0 * M * 2304 * 768)
if which in ("all", "fc1"):
    t(lambda: ops.gemm(x768, w_1, o3072, M=M, N=3072, K=768, bias=b3072, act="quick_gelu", bn=bn, cluster_m=cm), "fc1  f16 bias qgelu", 2.0 * M * 3072 * 768)
    t(lambda: ops.gemm(x768, w_1, o3072, M=M, N=3072, K=768, bias=b3072, bn=bn, cluster_m=cm), "fc1  f16 bias", 2.0 * M * 3072 * 768)
    t(lambda: ops.gemm(x768, w_1, o3072, M=M, N=3072, K=768, bn=bn, cluster_m=cm), "fc1  f16 plain", 2.0 * M * 3072 * 768)
if which in ("all", "fc2"):
    t(lambda: ops.gemm(x3072, w_2, o32, M=M, N=768, K=3072, bias=b768, resid=res, bn=bn, cluster_m=cm), "fc2  f32 bias resid", 2.0 * M * 768 * 3072)
if which in ("all", "out"):
    t(lambda: ops.gemm(x768, w_o, o32, M=M, N=768, K=768, bias=b768, resid=res, bn=bn, cluster_m=cm), "out  f32 bias resid", 2.0 * M * 768 * 768)
    t(lambda: ops.gemm(x768, w_o, o32, M=M, N=768, K=768, bn=bn, cluster_m=cm), "out  f32 plain", 2.0 * M * 768 * 768)
if which in ("all", "cublas"):
    # library reference on the same shapes (plain GEMM, no epilogue): what cuBLAS gets out of the chip here
    t(lambda: torch.matmul(x768, w_qkv.t()), "cublas qkv", 2.0 * M * 2304 * 768)
    t(lambda: torch.matmul(x768, w_1.t()), "cublas fc1", 2.0 * M * 3072 * 768)
    t(lambda: torch.matmul(x3072, w_2.t()), "cublas fc2", 2.0 * M * 768 * 3072)
    t(lambda: torch.matmul(x768, w_o.t()), "cublas out", 2.0 * M * 768 * 768)
    big = torch.randn(8192, 8192, device="cuda").half()
    t(lambda: torch.matmul(big, big), "cublas 8192^3", 2.0 * 8192 ** 3)
    o8 = torch.empty(8192, 8192, device="cuda", dtype=torch.half)
    t(lambda: ops.gemm(big, big, o8, M=8192, N=8192, K=8192, cluster_m=1), "ours 8192^3 cm=1", 2.0 * 8192 ** 3)
    t(lambda: ops.gemm(big, big, o8, M=8192, N=8192, K=8192, cluster_m=2), "ours 8192^3 cm=2", 2.0 * 8192 ** 3)
