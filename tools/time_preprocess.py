"""Times DevicePreprocessor on a batch of COCO-sized raw images (dev tool).  usage: python tools/time_preprocess.py [B H W]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import synth
from owl_vit_object_detection_b200.preprocess import DevicePreprocessor

B, H, W = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 480, 640)
imgs = [torch.from_numpy(synth.make_raw_image(H - 8 * (i % 3), W - 16 * (i % 4), seed=i)).cuda() for i in range(B)]
pre = DevicePreprocessor(768)
out = torch.empty((B, 3, 768, 768), dtype=torch.float32, device="cuda")
for _ in range(3):
    pre(imgs, out)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pre(imgs, out)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
mb = sum(i.numel() for i in imgs) / 1e6 + out.numel() * 4 / 1e6
print(f"preprocess {B} images ~{H}x{W} -> 768: {best * 1e3:.0f} us per batch ({mb / best:.0f} GB/s in + out)")
