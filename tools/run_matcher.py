"""Launches the matcher cost + LSAP kernels on one 16384-image chunk per T (for ncu captures; dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops
P, C, CH = 576, 80, 16384
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = torch.Generator(device="cuda").manual_seed(T)
sims = torch.rand((CH, P, C), generator=g, device="cuda") * 0.4 - 0.1
def boxes(k):
    cxy = 0.1 + 0.8 * torch.rand((CH, k, 2), generator=g, device="cuda")
    wh = 0.02 + 0.48 * torch.rand((CH, k, 2), generator=g, device="cuda")
    lo = (cxy - wh / 2).clamp(0.0, 1.0)
    return torch.cat([lo, torch.maximum((cxy + wh / 2).clamp(0.0, 1.0), lo + 1e-3)], dim=-1).contiguous()
pred, tgt = boxes(P), boxes(T)
lab = torch.randint(0, C, (CH, T), generator=g, device="cuda")
costT = torch.empty((CH, T, P), device="cuda")
match = torch.empty((CH, T), dtype=torch.int32, device="cuda")
status = torch.zeros(1, dtype=torch.int32, device="cuda")
nt = torch.full((CH,), T, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.matcher_cost(sims, pred, lab, tgt, nt, costT, status)
    ops.lsap(costT, nt, match, status)
torch.cuda.synchronize()
