"""Epilogue cost experiments on the out-proj / fc2 shapes (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops
M = 16 * 577
iters = 20
def t(fn, name, flops):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:40s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
def r16(*s): return (torch.randn(*s, device="cuda") * 0.05).half()
x768, x3072 = r16(M, 768), r16(M, 3072)
w_o, w_2 = r16(768, 768), r16(768, 3072)
b768 = torch.randn(768, device="cuda")
res, o32 = torch.randn(M, 768, device="cuda"), torch.empty(M, 768, device="cuda")
o16 = torch.empty(M, 768, device="cuda", dtype=torch.half)
for bn in (128, 256):
  for cm in (1, 2):
    tag = f"bn={bn} cm={cm}"
    f = 2.0 * M * 768 * 768
    t(lambda: ops.gemm(x768, w_o, o32, M=M, N=768, K=768, bn=bn, cluster_m=cm), "out f32 plain " + tag, f)
    t(lambda: ops.gemm(x768, w_o, o32, M=M, N=768, K=768, bias=b768, bn=bn, cluster_m=cm), "out f32 bias " + tag, f)
    t(lambda: ops.gemm(x768, w_o, o32, M=M, N=768, K=768, resid=res, bn=bn, cluster_m=cm), "out f32 resid " + tag, f)
    t(lambda: ops.gemm(x768, w_o, res, M=M, N=768, K=768, resid=res, bn=bn, cluster_m=cm), "out f32 resid inplace " + tag, f)
    t(lambda: ops.gemm(x768, w_o, o16, M=M, N=768, K=768, bias=b768, bn=bn, cluster_m=cm), "out f16 bias " + tag, f)
    f = 2.0 * M * 768 * 3072
    t(lambda: ops.gemm(x3072, w_2, o32, M=M, N=768, K=3072, bn=bn, cluster_m=cm), "fc2 f32 plain " + tag, f)
    t(lambda: ops.gemm(x3072, w_2, o32, M=M, N=768, K=3072, bias=b768, resid=res, bn=bn, cluster_m=cm), "fc2 f32 bias resid " + tag, f)
    t(lambda: ops.gemm(x3072, w_2, o16, M=M, N=768, K=3072, bias=b768, bn=bn, cluster_m=cm), "fc2 f16 bias " + tag, f)
