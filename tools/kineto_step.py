"""Per-kernel device times of one eager train step from the torch profiler (CUPTI activity records, no replay): dev tool.
usage: python tools/kineto_step.py [batch] > table"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from torch.profiler import profile, ProfilerActivity
from owl_vit_object_detection_b200 import synth
from owl_vit_object_detection_b200.loss import PushPullLoss
from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT
from owl_vit_object_detection_b200.train import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = synth.B32
sd = synth.make_weights(cfg, seed=0)
model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device="cuda")
crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).cuda())
opt = FusedAdamW(model, lr=3e-6, weight_decay=0.1)
step = TrainStep(model, crit, opt, batch=B, use_graph=False, n_input_slots=1)
lab, box, nt = synth.make_targets(cfg, B, seed=200)
step.load(synth.make_images(cfg, B, seed=100), lab, box, nt, slot=0)
for _ in range(3):
    step.run(slot=0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step.run(slot=0)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in ev:
    a = agg.setdefault(e.name[:110], [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.elapsed_us()
tot = sum(a[1] for a in agg.values())
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"# {len(ev)} kernels, sum of kernel times {tot:.1f} us, first-start to last-end {span:.1f} us (eager, PDL overlap)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {n:4d} x {t / n:8.1f}  {100 * t / tot:5.1f}%  {k}")
if len(sys.argv) > 2:
    print("# in order")
    for e in ev:
        print(f"{e.time_range.start - ev[0].time_range.start:10.1f} {e.time_range.elapsed_us():8.1f}  {e.name[:100]}")
