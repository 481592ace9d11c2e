"""Times Engine.forward at a given batch and prints per-kernel-group CUDA-event timings (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import synth
from owl_vit_object_detection_b200.engine import Engine
from owl_vit_object_detection_b200.params import ParamLayout

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = synth.B32
layout = ParamLayout(cfg)
flat = torch.randn(layout.total, device="cuda") * 0.02
eng = Engine(cfg, layout, flat)
img = torch.randn(B, 3, cfg.image_size, cfg.image_size, device="cuda")
for _ in range(3):
    eng.forward(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    eng.forward(img)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"eager forward B={B}: {ms:.3f} ms/iter, {B / ms * 1e3:.1f} img/s")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.forward(img)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        out = eng.forward(img)
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 1.1496e11 * B
print(f"graph forward B={B}: {ms:.3f} ms/iter, {B / ms * 1e3:.1f} img/s, {fl / ms / 1e9:.1f} TFLOP/s")
