"""GPU parity of owl_preprocess_image / DevicePreprocessor (reference src/dataset.py:64-71 -> PIL bicubic resize,
rescale, normalise) against the oracle and the fixtures made with the real PIL resize: bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as pre  # noqa: E402  (checker only)
from owl_vit_object_detection_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))


@pytest.mark.parametrize("i", range(len(synth.PREPROCESS_CASES)))
def test_preprocess_matches_pil_fixture_and_oracle(i):
    from owl_vit_object_detection_b200.preprocess import DevicePreprocessor
    h, w, size = synth.PREPROCESS_CASES[i]
    img = synth.make_raw_image(h, w, seed=i)
    out = DevicePreprocessor(size)([torch.from_numpy(img)])[0].cpu().numpy()
    np.testing.assert_array_equal(out[:, ::61, ::53], GOLD[f"sample_{i}"])
    np.testing.assert_array_equal(out, pre.preprocess(img, size))


def test_forward_accepts_raw_uint8():
    """OwlViT.forward on raw uint8 [B,H,W,3] == forward on the preprocessed fp32 tensor (same kernels after the resize)."""
    from owl_vit_object_detection_b200.model import OwlViT
    from owl_vit_object_detection_b200.preprocess import DevicePreprocessor
    cfg = synth.TINY
    sd = synth.make_weights(cfg, seed=0)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device="cuda")
    raw = torch.from_numpy(np.stack([synth.make_raw_image(90, 120, seed=s) for s in range(2)])).cuda()
    with torch.no_grad():
        b0, _, s0, _ = model(raw)
        b1, _, s1, _ = model(DevicePreprocessor(cfg.image_size)(list(raw)))
    assert torch.equal(b0, b1) and torch.equal(s0, s1)
