"""Dev tool: repeats tests/test_module_gpu.py::test_trainstep_host_buffers_graph_matches_eager and reports, per
trainable tensor, how far the graph-driven and the eager-driven parameters are apart after 5 steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from owl_vit_object_detection_b200 import synth
from owl_vit_object_detection_b200.train import TrainStep
from src.losses import PushPullLoss
from src.models import FusedAdamW, OwlViT

cfg = synth.TINY
B, n_slots, n_steps = 2, 2, 5
scales = synth.make_class_scales(cfg).cuda()
host = []
for s in range(3):
    img = synth.make_images(cfg, B, seed=40 + s).pin_memory()
    lab, box, nt = synth.make_targets(cfg, B, seed=50 + s, max_t=8)
    host.append((img, lab.pin_memory(), box.pin_memory(), nt.pin_memory()))


def drive(use_graph):
    sd = synth.make_weights(cfg, seed=0)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device="cuda")
    crit = PushPullLoss(cfg.n_classes, scales)
    opt = FusedAdamW(model, lr=1e-3, weight_decay=0.1)
    step = TrainStep(model, crit, opt, batch=B, max_targets=host[0][1].shape[1], use_graph=use_graph, n_input_slots=n_slots)
    for s in range(n_slots):
        step.load(*host[s], slot=s)
    torch.cuda.synchronize()
    step.warmup()
    pending, nxt = None, step.load(*host[0])
    for i in range(n_steps):
        cur = nxt
        if i + 1 < n_steps:
            nxt = step.load(*host[(i + 1) % len(host)])
        step.run(slot=cur, readback=True)
        if pending is not None:
            step.result(pending)
        pending = cur
    step.result(pending)
    torch.cuda.synchronize()
    return model, model.flat_params.detach().float().cpu().numpy().copy()


ref = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for mode in (True, False):
        model, p = drive(mode)
        if ref is None:
            ref = p
            continue
        L = model.layout
        rows = []
        for n in L.trainable:
            o, k = L.offsets[n], L._numel(n)
            d = np.abs(p[o:o + k] - ref[o:o + k])
            if d.max() > 2e-5:
                rows.append(f"{n}: {int((d > 2e-5).sum())}/{k} off, max {d.max():.2e}")
        print(f"rep {rep} graph={mode}: " + ("identical within 2e-5" if not rows else "; ".join(rows)), flush=True)
