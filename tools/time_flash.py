"""Times the fused attention forward with CUDA events (dev tool).
usage: [OWL_FA_GEN=24|26|34|36] python tools/time_flash.py [B S H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops

B, S, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 577, 12)
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn((B * S, 3 * D), generator=g, device="cuda") * 0.5).half()
ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")


def run():
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=64, scale=0.125)


for _ in range(5):
    run()
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
fl = 4.0 * B * H * S * S * 64
print(f"flash gen={os.environ.get('OWL_FA_GEN', 'default')} B={B} S={S} H={H}: {best * 1e3:.1f} us  {fl / best / 1e9:.0f} TFLOP/s")
