"""Runs a few eager (non-graph) train steps at batch 16 so that ncu sees one kernel per launch (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import _lib, synth
from owl_vit_object_detection_b200.loss import PushPullLoss
from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT
from owl_vit_object_detection_b200.train import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = synth.B32
sd = synth.make_weights(cfg, seed=0)
model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device="cuda")
crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).cuda())
opt = FusedAdamW(model, lr=3e-6, weight_decay=0.1)
step = TrainStep(model, crit, opt, batch=B, use_graph=False, n_input_slots=1, raw_u8=True)
lab, box, nt = synth.make_targets(cfg, B, seed=200)
step.load(synth.make_images_u8(cfg, B, seed=100), lab, box, nt, slot=0)
torch.cuda.synchronize()
for i in range(steps):
    l0 = _lib.KERNEL_LAUNCHES
    out = step.run(slot=0)
    torch.cuda.synchronize()
    print("step", i, "kernels", _lib.KERNEL_LAUNCHES - l0, out.tolist())
