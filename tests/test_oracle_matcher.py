"""Pins oracle/matcher_oracle.py against fixtures produced by the REAL reference HungarianMatcher /
PushPullLoss (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import matcher_oracle as mo
from owl_vit_object_detection_b200 import synth


@pytest.mark.parametrize("T", [10, 50, 100])
def test_matcher_indices_exact(golden_dir, T):
    g = np.load(os.path.join(golden_dir, f"matcher_T{T}.npz"))
    sims, pred, lab, tgt = synth.make_matcher_inputs(6, T, seed=4)
    tc, ind = mo.hungarian(sims, pred, list(lab), list(tgt), 80)
    for b in range(6):
        assert ind[b][0].tolist() == g[f"pred_idx{b}"].tolist()
        assert ind[b][1].tolist() == g[f"tgt_idx{b}"].tolist()
        assert tc[b].tolist() == g[f"tc{b}"].tolist()


@pytest.mark.parametrize("T", [10, 50, 100])
def test_loss_and_grads(golden_dir, T):
    g = np.load(os.path.join(golden_dir, f"matcher_T{T}.npz"))
    sims, pred, lab, tgt = synth.make_matcher_inputs(6, T, seed=4)
    scales = synth.make_class_scales(synth.B32)
    for b in range(2):
        s = sims[b].clone().requires_grad_(True)
        p = pred[b].clone().requires_grad_(True)
        l, _, _ = mo.push_pull_loss_image(s, p, lab[b], tgt[b], 80, scales)
        for k in ("loss_ce", "loss_bg", "loss_bbox", "loss_giou"):
            np.testing.assert_allclose(l[k].item(), g[f"{k}{b}"], rtol=1e-5, atol=1e-6)
        if b == 0:
            sum(l.values()).backward()
            np.testing.assert_allclose(s.grad.numpy(), g["dsims0"], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(p.grad.numpy(), g["dboxes0"], rtol=1e-4, atol=1e-7)


def test_propagation_chain_q7():
    # SURVEY Q7 known answer: links 0-1? no: [bg,1,bg,bg] with IoU links 1~2, 2~3 => [bg,1,1,1];
    # reversed order => no back-propagation.
    bg = 80
    def box(x):  # unit-height boxes shifted along x; IoU(x, x+0.05) = 0.95/1.05 > 0.85
        return [x, 0.0, x + 1.0, 1.0]
    boxes = torch.tensor([box(5.0), box(0.0), box(0.05), box(0.10)])
    tc = torch.tensor([bg, 1, bg, bg])
    assert mo.propagate_labels(tc, boxes, bg).tolist() == [bg, 1, 1, 1]
    boxes_r = torch.tensor([box(0.10), box(0.05), box(0.0), box(5.0)])
    tc_r = torch.tensor([bg, bg, 1, bg])
    assert mo.propagate_labels(tc_r, boxes_r, bg).tolist() == [bg, 1, 1, bg]


def test_giou_known_answers():
    a = torch.tensor([[0.0, 0.0, 1.0, 1.0]])
    assert mo.generalized_box_iou(a, a).item() == 1.0
    b = torch.tensor([[2.0, 2.0, 3.0, 3.0]])
    assert mo.generalized_box_iou(a, b).item() < 0
    with pytest.raises(AssertionError):
        mo.generalized_box_iou(torch.tensor([[1.0, 0.0, 0.0, 1.0]]), a)
