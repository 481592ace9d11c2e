"""Generates tests/golden/preprocess.npz from the REAL libraries the reference's preprocessing calls
(`src/dataset.py:64-71` -> HF OwlViTImageProcessor, pinned transformers 4.30.2 = PIL bicubic resize + rescale +
normalize): `PIL.Image.resize(..., BICUBIC)` run here on seeded synthetic uint8 images, then the 4.30.2
rescale / normalize op sequence in numpy.  The fixtures hold a SHA-256 of every resized uint8 image and a strided
sample of the float output (the full tensors are 7 MB each).  Run: `python tests/golden/make_golden_preprocess.py`."""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from owl_vit_object_detection_b200 import synth  # noqa: E402


def main():
    out = {}
    mean = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
    std = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)
    for i, (h, w, size) in enumerate(synth.PREPROCESS_CASES):
        img = synth.make_raw_image(h, w, seed=i)
        r = np.asarray(Image.fromarray(img).resize((size, size), resample=Image.BICUBIC))
        out[f"sha_{i}"] = np.frombuffer(hashlib.sha256(r.tobytes()).digest(), dtype=np.uint8)
        x = (r * (1 / 255)).astype(np.float32)                # transformers 4.30.2 image_transforms.rescale
        x = ((x - mean) / std).astype(np.float32)             # ... normalize (float32 mean / std)
        x = x.transpose(2, 0, 1)                              # channels first
        out[f"sample_{i}"] = x[:, ::61, ::53].copy()
        print(i, (h, w, size), r.shape, float(x.mean()))
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)


if __name__ == "__main__":
    main()
