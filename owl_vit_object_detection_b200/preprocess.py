"""Device-side replacement of the reference's per-item CPU preprocessing (reference src/dataset.py:64-71: every PIL
image goes through HF `OwlViTProcessor` = bicubic resize to 768 x 768, /255, CLIP mean / std, channels first).

`DevicePreprocessor` takes raw uint8 RGB images (HWC, any size) and produces the model's `[B, 3, S, S]` fp32
`pixel_values` with `owl_preprocess_batch` (Pillow-exact resample, three launches for a whole batch of images of
different sizes, see csrc/preprocess.cu).  The host ships one
byte per channel instead of a 7 MB fp32 tensor per image and no DataLoader worker has to resize anything.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import ops

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def rescale_normalize_lut(mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD, rescale_factor: float = 1 / 255) -> np.ndarray:
    """[3, 256] float32: the value channel c takes for byte v, with the op sequence of transformers 4.30.2
    `image_transforms.rescale` (uint8 * python float -> float64 -> float32) and `normalize` (float32 mean / std)."""
    v = np.arange(256, dtype=np.uint8)
    x = (v * rescale_factor).astype(np.float32)
    m, s = np.array(mean, dtype=np.float32), np.array(std, dtype=np.float32)
    return np.stack([((x - m[c]) / s[c]).astype(np.float32) for c in range(3)])


class DevicePreprocessor:
    def __init__(self, size: int = 768, device="cuda", mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD):
        self.size = size
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePreprocessor runs on a CUDA device only (there is no CPU fallback)")
        self.lut = torch.from_numpy(rescale_normalize_lut(mean, std)).to(self.device).contiguous()
        self._ws = torch.empty(0, dtype=torch.uint8, device=self.device)

    def __call__(self, images: Sequence[torch.Tensor], out: torch.Tensor | None = None) -> torch.Tensor:
        """images: uint8 [H_i, W_i, 3] tensors (host or device).  Returns fp32 [B, 3, size, size] on the device."""
        B, S = len(images), self.size
        if out is None:
            out = torch.empty((B, 3, S, S), dtype=torch.float32, device=self.device)
        assert out.shape == (B, 3, S, S) and out.is_contiguous()
        dev_images = []
        for b, im in enumerate(images):
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
                raise ValueError(f"image {b}: expected uint8 [H, W, 3], got {im.dtype} {tuple(im.shape)}")
            dev_images.append(im.to(self.device, non_blocking=True).contiguous())
        with torch.cuda.device(self.device):
            need = ops.preprocess_batch_workspace_bytes(dev_images, S)
            if self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            # the whole ragged batch in three launches per 32 images (stream-ordered: the workspace is reused)
            ops.preprocess_batch(dev_images, self.lut, out, self._ws)
        return out
