"""Times the fused attention forward with CUDA events (dev tool).  usage: [OWL_FA_GEN=1] python tools/time_flash.py [B S H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops
B, S, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 577, 12)
dh = 64
D = H * dh
qkv = torch.randn((B * S, 3 * D), device="cuda").half()
ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
lse = torch.zeros((B * H * S,), dtype=torch.float32, device="cuda")
for _ in range(5):
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    e0.record()
    for _ in range(20):
        ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
fl = 4.0 * B * H * S * S * dh
print(f"flash gen={os.environ.get('OWL_FA_GEN', '2')} B={B} S={S} H={H}: {best * 1e3:.1f} us  {fl / best / 1e9:.0f} TFLOP/s")
