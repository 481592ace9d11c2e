import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops, _lib
B, S, H, dh = 16, 577, 12, 64
D = H * dh
qkv = torch.randn((B * S, 3 * D), device="cuda").half()
ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
torch.cuda.synchronize()
_lib.lib().owl_flash_attn_debug(ctypes.c_void_p(dbg.data_ptr()))
ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
torch.cuda.synchronize()
_lib.lib().owl_flash_attn_debug(ctypes.c_void_p(0))
t = dbg.cpu().tolist()
t0 = t[0]
names = {0: "setup done", 1: "after pdl wait", 30: "PV last done", 31: "epilogue stored"}
for j in range(3):
    names[2 + 4 * j] = f"S{j} ready"; names[3 + 4 * j] = f"max{j} known"; names[4 + 4 * j] = f"P{j} written"; names[5 + 4 * j] = f"p_full{j} arrived"
for k in sorted(names):
    if t[k]: print(f"{names[k]:18s} +{(t[k] - t0) / 1000:.2f} us")
