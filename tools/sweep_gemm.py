"""Sweep tile / cluster flavours on one GEMM shape (dev tool): sweep_gemm.py M N K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
a = (torch.randn(M, K, device="cuda") * 0.05).half()
b = (torch.randn(N, K, device="cuda") * 0.05).half()
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda")
o16 = torch.empty(M, N, device="cuda", dtype=torch.half)
o32 = torch.empty(M, N, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
print(f"M={M} N={N} K={K}   (us)")
for bn in (128, 192, 256):
    for cm in (1, 2):
        r = [t(lambda: ops.gemm(a, b, o16, M=M, N=N, K=K, bn=bn, cluster_m=cm)),
             t(lambda: ops.gemm(a, b, o16, M=M, N=N, K=K, bn=bn, cluster_m=cm, bias=bias, act="quick_gelu")),
             t(lambda: ops.gemm(a, b, o32, M=M, N=N, K=K, bn=bn, cluster_m=cm)),
             t(lambda: ops.gemm(a, b, o32, M=M, N=N, K=K, bn=bn, cluster_m=cm, bias=bias, resid=res))]
        print(f"bn={bn} cm={cm}: f16 plain {r[0]:6.1f}  f16 bias+qgelu {r[1]:6.1f}  f32 plain {r[2]:6.1f}  f32 bias+resid {r[3]:6.1f}")
print(f"cublas f16: {t(lambda: torch.matmul(a, b.t())):6.1f}")
