"""Launches the fused attention backward a few times on the B/32 batch-16 shape (for ncu captures; dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops
B, S, H, dh = 16, 577, 12, 64
D = H * dh
scale = dh ** -0.5
qkv = torch.randn((B * S, 3 * D), device="cuda").half()
dctx = torch.randn((B * S, D), device="cuda").half()
ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
lse = torch.zeros((B * H, S), device="cuda"); delta = torch.zeros((B * H, S), device="cuda")
dqkv = torch.zeros((B * S, 3 * D), dtype=torch.float16, device="cuda"); dq32 = torch.empty((B * S, D), device="cuda")
ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=scale, lse=lse)
ops.attn_delta(ctx, dctx, delta, B=B, S=S, H=H, head_dim=dh, alpha=scale)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    ops.attn_bwd(qkv, dctx, lse, delta, dqkv, dq32, B=B, S=S, H=H, head_dim=dh, scale=scale)
torch.cuda.synchronize()
