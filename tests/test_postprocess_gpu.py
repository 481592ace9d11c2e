"""GPU parity of owl_postprocess / src.models.PostProcess (reference src/models.py:122-146) against the oracle and
the golden outputs of the real reference class: bit-exact boxes / classes / scores and survivor order."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postprocess_oracle as po  # noqa: E402  (checker only)
from owl_vit_object_detection_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "postprocess.npz"))


@pytest.mark.parametrize("name", list(synth.POSTPROCESS_CASES))
def test_postprocess_matches_reference_golden(name):
    from owl_vit_object_detection_b200 import ops
    conf, iou = synth.POSTPROCESS_CASES[name]
    boxes, sims = synth.make_postprocess_inputs(name)
    ob, oc, os_, cnt = ops.postprocess(boxes.cuda(), sims.cuda(), conf, iou)
    torch.cuda.synchronize()
    for b in range(boxes.shape[0]):
        k = int(cnt[b])
        assert k == GOLD[f"{name}_classes{b}"].shape[0]
        np.testing.assert_array_equal(oc[b, :k].cpu().numpy(), GOLD[f"{name}_classes{b}"])
        np.testing.assert_array_equal(ob[b, :k].cpu().numpy(), GOLD[f"{name}_boxes{b}"])
        np.testing.assert_array_equal(os_[b, :k].cpu().numpy(), GOLD[f"{name}_scores{b}"])


@pytest.mark.parametrize("P,C,B", [(576, 80, 16), (3600, 80, 2), (33, 5, 3), (1, 1, 1)])
def test_postprocess_matches_oracle(P, C, B):
    from owl_vit_object_detection_b200 import ops
    g = torch.Generator().manual_seed(P + C)
    cxy = 0.2 + 0.6 * torch.rand((B, P, 2), generator=g)
    wh = 0.05 + 0.4 * torch.rand((B, P, 2), generator=g)
    boxes = torch.cat([(cxy - wh / 2).clamp(0, 1), (cxy + wh / 2).clamp(0, 1)], -1).contiguous()
    sims = (torch.rand((B, P, C), generator=g) * 0.6 - 0.1).contiguous()
    ob, oc, os_, cnt = ops.postprocess(boxes.cuda(), sims.cuda(), 0.3, 0.4)
    torch.cuda.synchronize()
    for b in range(B):
        rb, rc, rs = po.postprocess_image(boxes[b].numpy(), sims[b].numpy(), 0.3, 0.4)
        k = int(cnt[b])
        assert k == rc.shape[0]
        np.testing.assert_array_equal(oc[b, :k].cpu().numpy(), rc)
        np.testing.assert_array_equal(ob[b, :k].cpu().numpy(), rb)
        np.testing.assert_array_equal(os_[b, :k].cpu().numpy(), rs)


def test_postprocess_class_signature():
    """src.models.PostProcess keeps the reference's call signature and return shapes (batch 1)."""
    from src.models import PostProcess
    conf, iou = synth.POSTPROCESS_CASES["dense"]
    boxes, sims = synth.make_postprocess_inputs("dense")
    pb, pc, ps = PostProcess(confidence_threshold=conf, iou_threshold=iou)(boxes[:1].cuda(), sims[:1].cuda())
    k = GOLD["dense_classes0"].shape[0]
    assert pb.shape == (1, k, 4) and pc.shape == (1, k) and ps.shape == (1, k) and pc.dtype == torch.int64
    np.testing.assert_array_equal(pb[0].cpu().numpy(), GOLD["dense_boxes0"])
