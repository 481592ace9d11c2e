"""Launches one GEMM flavour a few times (for ncu captures; dev tool): run_gemm.py out|fc2|qkv|fc1 [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops
M = 16 * 577
which = sys.argv[1] if len(sys.argv) > 1 else "out"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
r16 = lambda *s: (torch.randn(*s, device="cuda") * 0.05).half()
if which == "out":
    a, w, bias, res, o = r16(M, 768), r16(768, 768), torch.randn(768, device="cuda"), torch.randn(M, 768, device="cuda"), torch.empty(M, 768, device="cuda")
    f = lambda: ops.gemm(a, w, o, M=M, N=768, K=768, bias=bias, resid=res)
elif which == "fc2":
    a, w, bias, res, o = r16(M, 3072), r16(768, 3072), torch.randn(768, device="cuda"), torch.randn(M, 768, device="cuda"), torch.empty(M, 768, device="cuda")
    f = lambda: ops.gemm(a, w, o, M=M, N=768, K=3072, bias=bias, resid=res)
elif which == "qkv":
    a, w, bias, o = r16(M, 768), r16(2304, 768), torch.randn(2304, device="cuda"), torch.empty(M, 2304, device="cuda", dtype=torch.half)
    f = lambda: ops.gemm(a, w, o, M=M, N=2304, K=768, bias=bias)
else:
    a, w, bias, o = r16(M, 768), r16(3072, 768), torch.randn(3072, device="cuda"), torch.empty(M, 3072, device="cuda", dtype=torch.half)
    f = lambda: ops.gemm(a, w, o, M=M, N=3072, K=768, bias=bias, act="quick_gelu")
for _ in range(n):
    f()
torch.cuda.synchronize()
