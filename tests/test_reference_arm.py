"""The baseline arm of bench.py (`--impl reference`, cpu_baseline, torch_cuda_baseline, matcher reference timing) drives
the REAL reference classes through oracle/ref_arm.py: from /root/reference where it exists (this container), else from
the byte-compiled copy under oracle/_ref/ that `__graft_entry__.build()` makes (the GPU box).  Whichever is present
must reproduce the golden fixtures, which were generated from the real reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_arm  # checker only
from owl_vit_object_detection_b200 import synth


@pytest.mark.skipif(not ref_arm.available(), reason="neither /root/reference nor oracle/_ref is present")
def test_reference_arm_reproduces_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_b32.npz"))
    cfg = synth.B32
    model = ref_arm.build_model(cfg, synth.make_weights(cfg, seed=0))
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 8_791_812
    img = synth.make_images(cfg, 2, seed=2)
    with torch.no_grad():
        boxes, none1, sims, none2 = model(img[:1])
    assert none1 is None and none2 is None
    np.testing.assert_allclose(boxes[0].numpy(), g["boxes0"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(sims[0].numpy(), g["sims0"], rtol=0, atol=1e-6)
    # the matcher class is the reference's own, too
    T = 10
    s, p, lab, tgt = synth.make_matcher_inputs(6, T, seed=4)       # as tests/golden/make_golden.py drew them
    gm = np.load(os.path.join(golden_dir, f"matcher_T{T}.npz"))
    tc, ind, _ = ref_arm.matcher(80)({"pred_logits": s[:1], "pred_boxes": p[:1]}, [{"labels": lab[0], "boxes": tgt[0]}])
    assert np.array_equal(ind[0][0].numpy(), gm["pred_idx0"].astype(np.int64))
    assert np.array_equal(tc[0].numpy().astype(np.int16), gm["tc0"])
    # and our own drop-in `src` package is still what `import src` resolves to
    import src.models as ours
    assert ours.__file__.startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
