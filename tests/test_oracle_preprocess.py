"""The preprocessing oracle (oracle/preprocess_oracle.py: Pillow's bicubic resample + transformers 4.30.2
rescale / normalize, restated) against fixtures produced by the real PIL resize (tests/golden/preprocess.npz): the
resized uint8 image must hash identically, the float output must be bit-identical on the sampled grid."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as pre  # noqa: E402
from owl_vit_object_detection_b200 import synth  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))


@pytest.mark.parametrize("i", range(len(synth.PREPROCESS_CASES)))
def test_oracle_matches_pil(i):
    h, w, size = synth.PREPROCESS_CASES[i]
    img = synth.make_raw_image(h, w, seed=i)
    r = pre.resize_bicubic_u8(img, size, size)
    sha = np.frombuffer(hashlib.sha256(r.tobytes()).digest(), dtype=np.uint8)
    np.testing.assert_array_equal(sha, GOLD[f"sha_{i}"])
    x = pre.preprocess(img, size)
    np.testing.assert_array_equal(x[:, ::61, ::53], GOLD[f"sample_{i}"])


def test_identity_axis_and_constant_image():
    img = np.full((50, 70, 3), 200, dtype=np.uint8)
    assert (pre.resize_bicubic_u8(img, 96, 96) == 200).all()          # coefficients sum to one
    img = synth.make_raw_image(96, 40, seed=3)
    r = pre.resize_bicubic_u8(img, 96, 40)
    np.testing.assert_array_equal(r, img)                             # scale 1 on both axes is the identity


def test_oracle_matches_installed_huggingface_processor():
    """The class the reference calls at src/dataset.py:64-71 (`OwlViTProcessor` -> the OWL-ViT image processor).  The
    reference pins transformers 4.30.2, whose processor is PIL-based; the installed 5.5.0 keeps that implementation as
    `OwlViTImageProcessorPil` (constructed from its defaults: no download) and the oracle - hence the device kernels,
    which tests/test_preprocess_gpu.py holds bit-equal to the oracle - reproduces it BIT FOR BIT.  The torchvision-
    backed `OwlViTImageProcessor` of 5.5.0 is a different resampler: it differs by one uint8 level after a resize."""
    from PIL import Image
    transformers = pytest.importorskip("transformers")
    if not hasattr(transformers, "OwlViTImageProcessorPil"):
        pytest.skip("this transformers version has no PIL-backed OWL-ViT image processor")
    proc = transformers.OwlViTImageProcessorPil()
    assert proc.size["height"] == 768 and proc.size["width"] == 768
    for k, (h, w) in enumerate([(480, 640), (333, 500), (768, 768), (90, 120), (1, 1)]):
        raw = synth.make_raw_image(h, w, seed=20 + k)
        hf = proc(images=Image.fromarray(raw), return_tensors="np")["pixel_values"][0]
        np.testing.assert_array_equal(hf, pre.preprocess(raw, 768), err_msg=f"{h} x {w}")
    fast = transformers.OwlViTImageProcessor()
    raw = synth.make_raw_image(480, 640, seed=20)
    hf = fast(images=Image.fromarray(raw), return_tensors="pt")["pixel_values"][0].numpy()
    one_level = 1.0 / 255.0 / 0.26130258
    assert np.abs(hf - pre.preprocess(raw, 768)).max() <= 2.1 * one_level
