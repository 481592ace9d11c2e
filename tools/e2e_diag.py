"""Where does the e2e loop lose time against the device-resident loop?  (dev tool)"""
import sys, time
sys.path.insert(0, ".")
import torch
from owl_vit_object_detection_b200 import synth
from owl_vit_object_detection_b200.loss import PushPullLoss
from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT
from owl_vit_object_detection_b200.train import TrainStep

cfg, B, n_slots, steps = synth.B32, 16, 6, 100
dev = torch.device("cuda", 0)
sd = synth.make_weights(cfg, seed=0)
model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device=dev)
crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).to(dev))
opt = FusedAdamW(model, lr=3e-6, weight_decay=0.1)
step = TrainStep(model, crit, opt, batch=B, n_input_slots=n_slots, raw_u8=True)
host = []
for s in range(n_slots):
    lab, box, nt = synth.make_targets(cfg, B, seed=200 + s)
    host.append((synth.make_images_u8(cfg, B, seed=100 + s).pin_memory(), lab.pin_memory(), box.pin_memory(), nt.pin_memory()))
for s in range(n_slots):
    step.load(*host[s], slot=s)
torch.cuda.synchronize()
step.warmup()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(name, body):
    for _ in range(5):
        step.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    body()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1) / steps:.4f} ms/step (host wall {1e3 * (time.perf_counter() - t0) / steps:.4f})", flush=True)


def plain():
    for _ in range(steps):
        step.run()


def with_readback():
    pending = None
    for i in range(steps):
        cur = i % n_slots
        step.run(slot=cur, readback=True)
        if pending is not None:
            step.result(pending)
        pending = cur
    step.result(pending)


def with_load():
    nxt = step.load(*host[0])
    for i in range(steps):
        cur = nxt
        if i + 1 < steps:
            nxt = step.load(*host[(i + 1) % n_slots])
        step.run(slot=cur)


def with_image_only_load():
    for i in range(steps):
        s = i % n_slots
        step.copy_stream.wait_event(step.consumed[s])
        with torch.cuda.stream(step.copy_stream):
            step.slots[s]["image"].copy_(host[s][0], non_blocking=True)
            step.loaded[s].record(step.copy_stream)
        step.run(slot=s)


def full():
    nxt = step.load(*host[0])
    pending = None
    for i in range(steps):
        cur = nxt
        if i + 1 < steps:
            nxt = step.load(*host[(i + 1) % n_slots])
        step.run(slot=cur, readback=True)
        if pending is not None:
            step.result(pending)
        pending = cur
    step.result(pending)


for rep in range(2):
    timed("device-resident (value loop)", plain)
    timed("+ readback / result only", with_readback)
    timed("+ load only (4 H2D copies)", with_load)
    timed("+ image-only load", with_image_only_load)
    timed("full e2e loop", full)
