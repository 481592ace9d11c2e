"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference's image preprocessing.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs may import this module.

The reference (`src/dataset.py:64-71`) hands every PIL image to HuggingFace's `OwlViTProcessor`
(transformers==4.30.2 pinned, `requirements.txt:1`; source not under /root/reference), whose image side is
`OwlViTImageProcessor.preprocess`: PIL bicubic resize to 768x768 (`image_transforms.resize` ->
`PIL.Image.resize(resample=BICUBIC)`), `rescale` by 1/255 (uint8 * python float -> float64 -> float32), `normalize`
with the CLIP mean / std in float32, channels first.  The installed transformers 5.5.0 replaced this by a torchvision
backend with different resize numerics, so the restated algorithm is the PINNED one:

  * `resize_bicubic_u8`  = Pillow's `ImagingResample` for 8-bit channels (src/libImaging/Resample.c, Pillow 12.2:
    `precompute_coeffs`, `normalize_coeffs_8bpc` with PRECISION_BITS = 22, horizontal pass then vertical pass, each
    rounded to uint8 through `clip8`), bicubic filter a = -0.5;
  * `rescale_normalize`  = transformers 4.30.2 `image_transforms.rescale` + `normalize`.

Parity pin: tests/golden/preprocess.npz holds SHA-256 digests of real `PIL.Image.resize` outputs and samples of the
real pipeline's float output (tests/golden/make_golden_preprocess.py); tests/test_oracle_preprocess.py checks this
file against them bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size): (ksize, bounds[out,2], kk[out,ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One pass along axis 0 of an [n, m, c] uint8 array (integer arithmetic exactly as ResampleVertical)."""
    _, bounds, kk = precompute_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for yy in range(out_size):
        ymin, ymax = bounds[yy]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for y in range(ymax):
            acc += src[ymin + y] * int(kk[yy, y])
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[H, W, C] uint8 -> [out_h, out_w, C] uint8, Pillow's two-pass order: horizontal first, then vertical."""
    assert img.dtype == np.uint8 and img.ndim == 3
    tmp = _resample_axis0(np.ascontiguousarray(img.transpose(1, 0, 2)), out_w).transpose(1, 0, 2)   # horizontal
    return _resample_axis0(np.ascontiguousarray(tmp), out_h)                                          # vertical


def rescale_normalize_lut() -> np.ndarray:
    """[3, 256] float32: value of channel c for byte v, computed with the 4.30.2 op sequence."""
    v = np.arange(256, dtype=np.uint8)
    x = (v * (1 / 255)).astype(np.float32)                     # uint8 * python float -> float64 -> float32
    mean = np.array(CLIP_MEAN, dtype=np.float32)
    std = np.array(CLIP_STD, dtype=np.float32)
    return np.stack([((x - mean[c]) / std[c]).astype(np.float32) for c in range(3)])


def preprocess(img: np.ndarray, size: int = 768) -> np.ndarray:
    """[H, W, 3] uint8 RGB -> [3, size, size] float32 (pixel_values of one image)."""
    r = resize_bicubic_u8(img, size, size)
    lut = rescale_normalize_lut()
    return np.stack([lut[c][r[:, :, c]] for c in range(3)])
