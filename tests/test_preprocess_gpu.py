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


def test_raw_uint8_at_model_resolution_is_one_fused_pass_and_bit_equal():
    """Images that already have the model's resolution skip the resample (PIL's resize is the identity there) and are
    normalised inside the patch gather (owl_u8_patches_f16).  The patch rows must equal, bit for bit, what the
    reference's preprocessing (oracle: preprocess_oracle.preprocess, pinned to real PIL fixtures) followed by the fp32
    path's im2col produces - and so must the model outputs."""
    from owl_vit_object_detection_b200 import ops
    from owl_vit_object_detection_b200.model import OwlViT
    cfg = synth.TINY
    IS, ps = cfg.image_size, cfg.patch_size
    raw_np = np.stack([synth.make_raw_image(IS, IS, seed=10 + s) for s in range(3)])
    raw = torch.from_numpy(raw_np).cuda()
    ref32 = torch.from_numpy(np.stack([pre.preprocess(im, IS) for im in raw_np])).cuda()      # [B,3,IS,IS] fp32
    sd = synth.make_weights(cfg, seed=0)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device="cuda")
    K = 3 * ps * ps
    pa = torch.zeros((3 * cfg.patches, K), dtype=torch.float16, device="cuda")
    pb = torch.zeros_like(pa)
    ops.u8_patches_f16(raw, model.engine.pixel_lut(), pa, ps)
    ops.im2col_f16(ref32, pb, ps)
    assert torch.equal(pa, pb)
    with torch.no_grad():
        b0, _, s0, _ = model(raw)
        b1, _, s1, _ = model(ref32)
    assert torch.equal(b0, b1) and torch.equal(s0, s1)


def test_ragged_batch_in_three_launches_equals_per_image():
    """A batch of images of different sizes through ONE owl_preprocess_batch call (3 launches per 32 images) equals the
    oracle per image, and the single-image entry point, bit for bit; 40 images exercise the chunking."""
    from owl_vit_object_detection_b200 import _lib, ops
    from owl_vit_object_detection_b200.preprocess import DevicePreprocessor
    size = 96
    geoms = [(37 + 11 * (i % 7), 29 + 13 * (i % 5)) for i in range(40)] + [(1, 1), (200, 3), (96, 96)]
    raws = [synth.make_raw_image(h, w, seed=50 + i) for i, (h, w) in enumerate(geoms)]
    pre_dev = DevicePreprocessor(size)
    l0 = _lib.KERNEL_LAUNCHES
    out = pre_dev([torch.from_numpy(r) for r in raws])
    assert _lib.KERNEL_LAUNCHES - l0 == 3 * 2          # 43 images = two chunks of <= 32
    out = out.cpu().numpy()
    for i, r in enumerate(raws):
        np.testing.assert_array_equal(out[i], pre.preprocess(r, size), err_msg=f"image {i} {geoms[i]}")
    # single-image C entry point (kept for callers that hold one image)
    im = torch.from_numpy(raws[3]).cuda()
    ws = torch.empty(ops.preprocess_workspace_bytes(im.shape[0], im.shape[1], size), dtype=torch.uint8, device="cuda")
    one = torch.empty((3, size, size), dtype=torch.float32, device="cuda")
    ops.preprocess_image(im, pre_dev.lut, one, ws)
    np.testing.assert_array_equal(one.cpu().numpy(), out[3])
