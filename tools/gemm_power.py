"""SM clock and board power while ONE GEMM shape runs back to back for ~1 s, per OWL_GEMM_DEBUG mode (dev tool).
usage: gemm_power.py M N K bn cm"""
import os, subprocess, sys, time
if len(sys.argv) > 6:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from owl_vit_object_detection_b200 import ops
    M, N, K, bn, cm = (int(x) for x in sys.argv[1:6])
    a = (torch.randn(M, K, device="cuda") * 0.05).half()
    b = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.half)
    mode = sys.argv[6]
    if mode == "cublas":
        fn = lambda: torch.matmul(a, b.t(), out=o16)
    else:
        fn = lambda: ops.gemm(a, b, o16, M=M, N=N, K=K, bn=bn, cluster_m=cm, bias=bias, act="quick_gelu")
    for _ in range(10): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(200): fn()
    torch.cuda.synchronize()
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"],
                           stdout=subprocess.PIPE, text=True)
    time.sleep(0.2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 150
    for _ in range(reps): g.replay()
    e1.record()
    torch.cuda.synchronize()
    smi.terminate()
    out = smi.communicate()[0].strip().splitlines()
    rows = [tuple(float(x) for x in l.split(",")) for l in out if "," in l]
    busy = [r for r in rows if r[1] > 400]
    us = e0.elapsed_time(e1) / (reps * 200) * 1e3
    if busy:
        clk = sorted(r[0] for r in busy)[len(busy) // 2]
        pw = sorted(r[1] for r in busy)[len(busy) // 2]
    else:
        clk = pw = float("nan")
    print(f"{mode:>8s} debug={os.environ.get('OWL_GEMM_DEBUG', '0')}: {us:6.2f} us/launch  median SM clock {clk:.0f} MHz  power {pw:.0f} W  ({len(busy)} samples)  "
          f"-> {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s, {us * clk / 1e0:.0f} clk*1e-3... per launch {us * clk:.0f} Mclk*1e-6")
else:
    for mode, dbg in (("ours", "0"), ("ours", "1"), ("ours", "3"), ("ours", "4"), ("cublas", "0")):
        env = dict(os.environ, OWL_GEMM_DEBUG=dbg)
        subprocess.run([sys.executable, __file__] + sys.argv[1:6] + [mode], env=env)
