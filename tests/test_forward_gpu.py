"""GPU parity of the forward path (Engine.forward -> libowl_b200.so kernels) against
  * the golden outputs of the REAL reference on the seeded B/32 weights (tests/golden/model_b32.npz), and
  * the fp32 oracle (oracle/owlvit_oracle.py) on a tiny configuration.

Tolerance (stated, fp16 storage + fp32 accumulation through 12 layers): |pred_sims diff| <= 1.5e-3 (sims
live in [-1,1]), |pred_boxes diff| <= 2e-3 (boxes live in ~[0,1]); measured on B200: 1.7e-4 / 3.8e-4.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import owlvit_oracle as oo  # noqa: E402  (checker only)
from owl_vit_object_detection_b200 import synth  # noqa: E402

SIMS_ATOL = 1.5e-3
BOX_ATOL = 2e-3


def _engine(cfg, seed=0):
    from owl_vit_object_detection_b200.engine import Engine
    from owl_vit_object_detection_b200.params import ParamLayout
    sd = synth.make_weights(cfg, seed=seed)
    layout = ParamLayout(cfg)
    flat = layout.pack(sd, "cuda")
    return Engine(cfg, layout, flat), sd


def test_tiny_vs_oracle():
    cfg = synth.TINY
    eng, sd = _engine(cfg, seed=1)
    img = synth.make_images(cfg, 3, seed=5)
    boxes, sims = eng.forward(img.cuda())
    torch.cuda.synchronize()
    rb, rs = oo.forward(sd, cfg, img)
    print("tiny max err boxes %.2e sims %.2e" % ((boxes.cpu() - rb).abs().max(), (sims.cpu() - rs).abs().max()))
    np.testing.assert_allclose(boxes.cpu().numpy(), rb.numpy(), rtol=0, atol=BOX_ATOL)
    np.testing.assert_allclose(sims.cpu().numpy(), rs.numpy(), rtol=0, atol=SIMS_ATOL)


def test_b32_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_b32.npz"))
    cfg = synth.B32
    eng, _ = _engine(cfg, seed=0)
    img = synth.make_images(cfg, 2, seed=2).cuda()
    boxes, sims = eng.forward(img)
    # batch independence: image 0 alone gives the same numbers
    boxes1, sims1 = eng.forward(img[:1])
    torch.cuda.synchronize()
    for b in range(2):
        eb = np.abs(boxes[b].cpu().numpy() - g[f"boxes{b}"]).max()
        es = np.abs(sims[b].cpu().numpy() - g[f"sims{b}"]).max()
        print(f"B/32 image {b}: max err boxes {eb:.2e} sims {es:.2e}")
        assert eb <= BOX_ATOL and es <= SIMS_ATOL
    assert torch.equal(boxes1[0], boxes[0]) and torch.equal(sims1[0], sims[0])


def test_l14_840_forward_vs_oracle():
    """BASELINE.json configs[3] shape (OWL-ViT-L/14 @ 840: 3601 tokens, 24 layers, 16 heads, patch 14 -> K = 588
    padded to 592): forward only (not a reference capability; dims come from the config).  Two layers
    are enough to exercise every L/14-specific shape; the full 24-layer forward runs in bench tooling."""
    import dataclasses
    cfg = dataclasses.replace(synth.L14, layers=2)
    eng, sd = _engine(cfg, seed=3)
    img = synth.make_images(cfg, 1, seed=7)
    boxes, sims = eng.forward(img.cuda(), save_for_backward=False)
    torch.cuda.synchronize()
    rb, rs = oo.forward(sd, cfg, img)
    eb, es = (boxes.cpu() - rb).abs().max().item(), (sims.cpu() - rs).abs().max().item()
    print(f"L/14@840 (2 layers): max err boxes {eb:.2e} sims {es:.2e}")
    assert eb <= BOX_ATOL and es <= SIMS_ATOL
