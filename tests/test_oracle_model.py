"""Pins oracle/owlvit_oracle.py (the fp32 restatement of the reference forward) against outputs of the
REAL reference OwlViT wrapper + HF tower on the seeded weights (tests/golden/model_b32.npz)."""
import os

import numpy as np
import torch

from oracle import matcher_oracle as mo
from oracle import owlvit_oracle as oo
from owl_vit_object_detection_b200 import synth


def test_forward_and_train_grads_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_b32.npz"))
    cfg = synth.B32
    sd = synth.make_weights(cfg, seed=0)
    train = synth.trainable_names(cfg)
    assert len(train) == 29 and sum(sd[n].numel() for n in train) == 8_791_812
    for n in train:
        sd[n].requires_grad_(True)
    image = synth.make_images(cfg, 2, seed=2)
    labels, tboxes, nt = synth.make_targets(cfg, 2, seed=3)
    scales = synth.make_class_scales(cfg)
    torch.set_num_threads(os.cpu_count() or 1)
    boxes, sims = oo.forward(sd, cfg, image[:1])
    np.testing.assert_allclose(boxes[0].detach().numpy(), g["boxes0"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(sims[0].detach().numpy(), g["sims0"], atol=2e-5, rtol=0)
    t = int(nt[0])
    l, _, _ = mo.push_pull_loss(sims, boxes, [labels[0, :t]], [tboxes[0, :t]], cfg.n_classes, scales)
    for k in ("loss_ce", "loss_bg", "loss_bbox", "loss_giou"):
        np.testing.assert_allclose(l[k].item(), g[k + "0"], rtol=2e-4)
    sum(l.values()).backward()
    for n in train:
        got = synth.subsample(sd[n].grad).numpy()
        ref = g["grad0." + n]
        scale = max(float(np.abs(ref).max()), 1e-12)
        # k_proj.bias has a mathematically zero gradient (softmax shift invariance): pure rounding noise
        assert np.abs(got - ref).max() <= 2e-3 * scale + 1e-7, n
        np.testing.assert_allclose(sd[n].grad.norm().item(), g["gnorm0." + n], rtol=2e-3, atol=1e-6)
