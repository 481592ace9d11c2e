"""In-step LSAP latency (16 COCO-shaped images): CTA-per-image vs warp-per-image kernel (dev tool).
usage: OWL_LSAP_MODE=1|2 python tools/lsap_step_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops, synth
cfg = synth.B32
B, P, C = 16, 576, 80
res = []
for seed in range(200, 206):
    lab, box, nt = synth.make_targets(cfg, B, seed=seed)
    g = torch.Generator().manual_seed(seed)
    sims = (torch.rand((B, P, C), generator=g) * 0.4 - 0.1).cuda()
    cxy = 0.1 + 0.8 * torch.rand((B, P, 2), generator=g); wh = 0.02 + 0.48 * torch.rand((B, P, 2), generator=g)
    lo = (cxy - wh / 2).clamp(0, 1); pred = torch.cat([lo, torch.maximum((cxy + wh / 2).clamp(0, 1), lo + 1e-3)], -1).cuda()
    lab, box, nt = lab.cuda(), box.cuda(), nt.cuda()
    costT = torch.zeros((B, lab.shape[1], P), device="cuda"); match = torch.zeros((B, lab.shape[1]), dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.matcher_cost(sims, pred, lab, box, nt, costT, status)
    for _ in range(3): ops.lsap(costT, nt, match, status)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.lsap(costT, nt, match, status)
    e1.record(); torch.cuda.synchronize()
    res.append((int(nt.max()), e0.elapsed_time(e1) / 20 * 1e3, match.clone()))
print("mode", os.environ.get("OWL_LSAP_MODE", "auto"), [(t, round(us, 1)) for t, us, _ in res])
torch.save([m.cpu() for _, _, m in res], f"/tmp/lsap_mode_{os.environ.get('OWL_LSAP_MODE', 'auto')}.pt")
