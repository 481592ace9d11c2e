"""One small invocation of the whole hot path on cuda:0, checked against the oracle (driver smoke test)."""
from __future__ import annotations

import torch


def run(verbose: bool = True) -> None:
    from oracle import matcher_oracle as mo       # checker only
    from oracle import owlvit_oracle as oo        # checker only
    from . import synth
    from .loss import PushPullLoss
    from .model import FusedAdamW, OwlViT

    cfg = synth.TINY
    B = 2
    sd = synth.make_weights(cfg, seed=1)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device="cuda:0")
    img = synth.make_images(cfg, B, seed=5)
    labels, tboxes, nt = synth.make_targets(cfg, B, seed=3, max_t=8)
    scales = synth.make_class_scales(cfg)
    crit = PushPullLoss(cfg.n_classes, scales.cuda())
    opt = FusedAdamW(model, lr=3e-6, weight_decay=0.1)

    opt.zero_grad()
    boxes, _, sims, _ = model(img.cuda())
    losses = crit(sims, labels.cuda(), boxes, tboxes.cuda(), num_targets=nt.cuda())
    total = losses["loss_ce"] + losses["loss_bg"] + losses["loss_bbox"] + losses["loss_giou"]
    total.backward()
    before = model.flat_trainable.clone()
    opt.step()
    torch.cuda.synchronize()
    crit.check_status()

    rb, rs = oo.forward(sd, cfg, img)
    eb = (boxes.detach().cpu() - rb).abs().max().item()
    es = (sims.detach().cpu() - rs).abs().max().item()
    assert eb < 2e-3 and es < 1.5e-3, f"forward mismatch vs oracle: boxes {eb:.2e} sims {es:.2e}"
    lab_l = [labels[b, :nt[b]] for b in range(B)]
    box_l = [tboxes[b, :nt[b]] for b in range(B)]
    ref_losses, _, _ = mo.push_pull_loss(sims.detach().cpu(), boxes.detach().cpu(), lab_l, box_l, cfg.n_classes, scales)
    for k in ref_losses:
        a, r = losses[k].item(), ref_losses[k].item()
        assert abs(a - r) <= 1e-4 * max(1.0, abs(r)), f"{k}: {a} vs oracle {r}"
    g = model.flat_grad
    assert torch.isfinite(g).all() and g.abs().max().item() > 0, "gradients missing / not finite"
    assert not torch.equal(before, model.flat_trainable), "optimizer step did not change the parameters"
    # eval-side rows: PostProcess (src/models.py:122-146) and device preprocessing (src/dataset.py:64-71), bit-exact
    import numpy as np
    from oracle import postprocess_oracle as po   # checker only
    from oracle import preprocess_oracle as pre   # checker only
    from . import ops
    from .preprocess import DevicePreprocessor
    pb, ps = synth.make_postprocess_inputs("dense", n_images=1)
    ob, oc, osc, cnt = ops.postprocess(pb.cuda(), ps.cuda(), 0.01, 0.6)
    rb2, rc2, rs2 = po.postprocess_image(pb[0].numpy(), ps[0].numpy(), 0.01, 0.6)
    k = int(cnt[0])
    assert k == rc2.shape[0] and np.array_equal(oc[0, :k].cpu().numpy(), rc2) and \
        np.array_equal(ob[0, :k].cpu().numpy(), rb2), "postprocess mismatch vs oracle"
    raw = synth.make_raw_image(60, 80, seed=1)
    px = DevicePreprocessor(cfg.image_size, "cuda:0")([torch.from_numpy(raw)])[0].cpu().numpy()
    assert np.array_equal(px, pre.preprocess(raw, cfg.image_size)), "preprocess mismatch vs oracle"
    if verbose:
        print(f"smoke ok: forward err boxes {eb:.1e} sims {es:.1e}; losses "
              + ", ".join(f"{k}={v.item():.4f}" for k, v in losses.items())
              + f"; |grad|max {g.abs().max().item():.3e}")
