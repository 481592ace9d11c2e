"""Where does a GEMM's time go?  Runs one shape with OWL_GEMM_DEBUG = 0 (normal), 1 (epilogue does nothing),
2 (no TMA loads), 3 (neither: MMA issue only).  dev tool: gemm_debug.py M N K"""
import os, subprocess, sys
if len(sys.argv) > 4:
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
    out = []
    for bn in (128, 192, 256):
        for cm in (1, 2):
            out.append(f"bn={bn} cm={cm}: f16+bias+qgelu {t(lambda: ops.gemm(a, b, o16, M=M, N=N, K=K, bn=bn, cluster_m=cm, bias=bias, act='quick_gelu')):6.1f}"
                       f"  f32+bias+resid {t(lambda: ops.gemm(a, b, o32, M=M, N=N, K=K, bn=bn, cluster_m=cm, bias=bias, resid=res)):6.1f}")
    print(f"OWL_GEMM_DEBUG={os.environ.get('OWL_GEMM_DEBUG', '0')}  M={M} N={N} K={K} (us)\n  " + "\n  ".join(out))
else:
    for dbg in ("0", "1", "3"):
        env = dict(os.environ, OWL_GEMM_DEBUG=dbg)
        subprocess.run([sys.executable, __file__] + sys.argv[1:4] + ["child"], env=env)
