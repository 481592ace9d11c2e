"""Per-tensor report for SURVEY R14: FusedAdamW vs torch.optim.AdamW over model.parameters() (reference
main.py:56-60,91), three steps on the TINY configuration.  Prints, per parameter tensor, how many elements are more
than `tol` apart and the largest difference; also run-to-run differences of each path with itself, and whether the
fp16 GEMM operands follow the fp32 parameters after every optimizer step."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from owl_vit_object_detection_b200 import synth  # noqa: E402


def run(kind, steps=3, lr=1e-3, check_shadow=True, cfg=synth.TINY, B=2):
    from src.losses import PushPullLoss
    from src.models import FusedAdamW, OwlViT
    sd = synth.make_weights(cfg, seed=1)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg).to("cuda")
    img = synth.make_images(cfg, B, seed=15).cuda()
    labels, tboxes, nt = [x.cuda() for x in synth.make_targets(cfg, B, seed=13, max_t=8)]
    crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).cuda())
    fused = kind == "fused"
    opt = (FusedAdamW(model, lr=lr, weight_decay=0.1) if fused
           else torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=0.1))
    grads = []
    stale = 0
    for _ in range(steps):
        opt.zero_grad(set_to_none=False) if fused else opt.zero_grad()
        boxes, _, sims, _ = model(img)
        if check_shadow:
            eng = model.engine
            lo = model.layout.train_begin
            stale += int((eng.flat16[lo:] != model.flat_params[lo:].half()).sum().item())
        l = crit(sims, labels, boxes, tboxes, num_targets=nt)
        (l["loss_ce"] + l["loss_bg"] + l["loss_bbox"] + l["loss_giou"]).backward()
        grads.append(model.flat_grad.detach().cpu().numpy().copy())
        opt.step()
    torch.cuda.synchronize()
    return model, model.flat_params.detach().cpu().numpy().copy(), grads, stale


def report(model, a, b, tol, title):
    L = model.layout
    print(f"--- {title}: elements with |diff| > {tol:g}")
    tot = 0
    for n in L.trainable:
        o = L.offsets[n]
        k = L._numel(n)
        d = np.abs(a[o:o + k] - b[o:o + k])
        bad = int((d > tol).sum())
        tot += bad
        if bad or d.max() > tol / 10:
            print(f"  {n:55s} n={k:7d} bad={bad:7d} max={d.max():.3e}")
    print(f"  total bad {tot}")


def main():
    m, pf, gf, sf = run("fused")
    _, pf2, gf2, _ = run("fused")
    _, pt, gt, st = run("torch")
    _, pt2, gt2, _ = run("torch")
    print("stale fp16 operand elements seen at forward time: fused", sf, "torch", st)
    report(m, pf, pf2, 2e-5, "fused vs fused (run to run)")
    report(m, pt, pt2, 2e-5, "torch vs torch (run to run)")
    report(m, pf, pt, 2e-5, "fused vs torch, params after 3 steps")
    lo = m.layout.train_begin
    for s in range(3):
        a = np.zeros_like(pf)
        b = np.zeros_like(pf)
        a[lo:], b[lo:] = gf[s], gt[s]
        scale = np.abs(gf[s]).max()
        report(m, a, b, 1e-6 * scale, f"gradients of step {s} fused-run vs torch-run (max |g| {scale:.3e})")


if __name__ == "__main__":
    main()
