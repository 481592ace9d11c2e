"""OWL-ViT-L/14 840x840 (BASELINE.json configs[3], an extension over the reference: SURVEY D5) train step, batch 4
per GPU: runs a few steps, prints losses and the step time (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import synth
from owl_vit_object_detection_b200.loss import PushPullLoss
from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT
from owl_vit_object_detection_b200.train import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = synth.L14
sd = synth.make_weights(cfg, seed=0)
model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device="cuda")
crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).cuda())
opt = FusedAdamW(model, lr=3e-6, weight_decay=0.1)
step = TrainStep(model, crit, opt, batch=B, n_input_slots=1)
lab, box, nt = synth.make_targets(cfg, B, seed=200)
step.load(synth.make_images(cfg, B, seed=100), lab, box, nt, slot=0)
step.warmup()
torch.cuda.synchronize()
for i in range(3):
    print("step", i, step.run(slot=0).tolist())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    out = step.run(slot=0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"L/14 840px batch {B}: {ms:.2f} ms/step = {B / ms * 1e3:.1f} images/s; losses {out.tolist()}; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
