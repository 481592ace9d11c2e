"""Generates tests/golden/train3_b32.npz: THREE optimizer steps of the REAL reference (read-only at /root/reference),
replaying reference main.py:56-91 literally - `torch.optim.AdamW(model.parameters(), lr, weight_decay)`,
`optimizer.zero_grad()`, forward, `PushPullLoss`, `loss.backward()`, `optimizer.step()` - on CPU fp32, batch 1.

Run here (the build container), never on the GPU box: `python tests/golden/make_golden_train3.py`.
Stored: the four losses and the (sub-sampled) forward outputs of every step and, per trainable tensor, the
(sub-sampled) parameter displacement after the three steps.  Inputs and initial weights are the seeded tensors of owl_vit_object_detection_b200.synth.

lr = 1e-4 (the reference's config.yaml uses 3e-6: three such steps move a weight by ~1e-5, below what an fp16-operand
forward can resolve; 1e-4 keeps the same code path with a displacement that is measurable), weight_decay = 0.1
(config.yaml).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from owl_vit_object_detection_b200 import synth  # noqa: E402
from make_golden import build_reference_model, load_reference, sub  # noqa: E402

LR, WD, STEPS = 1e-4, 0.1, 3
# Images of synth.make_images(cfg, 10, seed=2) / make_targets(cfg, 10, seed=3) used at the three steps.  Chosen so that
# no discrete decision of the loss (assignment, IoU > 0.85 label sweep) sits on its threshold: with Gaussian noise of
# twice the fp16 path's deviation on the forward outputs the reference loss moves by <= 1.2 % at every step for (8, 9, 8),
# while e.g. image 0 moves by 10 % (a flipped assignment), which would make any comparison after its step a coin toss.
SEQ, N_IMG = (8, 9, 8), 10


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    rmodels, rlosses, _ = load_reference()
    cfg = synth.B32
    sd = synth.make_weights(cfg, seed=0)
    model = build_reference_model(rmodels, cfg, sd)
    image = synth.make_images(cfg, N_IMG, seed=2)
    labels, tboxes, nt = synth.make_targets(cfg, N_IMG, seed=3)
    crit = rlosses.PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg))
    opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD)          # reference main.py:56-60
    start = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    out = {"lr": np.float32(LR), "weight_decay": np.float32(WD), "steps": np.int32(STEPS),
           "seq": np.array(SEQ, dtype=np.int32), "n_images": np.int32(N_IMG)}
    model.train()
    for step in range(STEPS):
        b = SEQ[step]
        t = int(nt[b])
        opt.zero_grad()                                                          # reference main.py:74
        boxes, _, sims, _ = model(image[b:b + 1])                                # :82
        losses = crit(sims, labels[b:b + 1, :t], boxes, tboxes[b:b + 1, :t])     # :83
        loss = losses["loss_ce"] + losses["loss_bg"] + losses["loss_bbox"] + losses["loss_giou"]
        loss.backward()                                                          # :90
        opt.step()                                                               # :91
        for k, v in losses.items():
            out[f"{k}{step}"] = np.float32(v.item())
        out[f"sims{step}"] = sub(sims[0])          # forward outputs of every step: the continuous quantities a
        out[f"boxes{step}"] = sub(boxes[0])        # consumer can compare without going through discrete decisions
        print(step, {k: round(v.item(), 5) for k, v in losses.items()})
    for n, p in model.named_parameters():
        if p.requires_grad:
            d = p.detach() - start[n]
            out["disp." + n] = sub(d)
            out["dispnorm." + n] = np.float32(d.norm().item())
    np.savez_compressed(os.path.join(HERE, "train3_b32.npz"), **out)
    print("train3_b32.npz written")


if __name__ == "__main__":
    main()
