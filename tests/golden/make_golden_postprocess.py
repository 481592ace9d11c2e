"""Generates tests/golden/postprocess.npz by running the REAL reference `PostProcess` (read-only at
/root/reference, src/models.py:122-146) on the CPU over seeded inputs (owl_vit_object_detection_b200.synth).
Run here (the build container), never on the GPU box: `python tests/golden/make_golden_postprocess.py`."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from owl_vit_object_detection_b200 import synth  # noqa: E402


def main():
    # the reference's `src` is a namespace package (no __init__.py); ours is a regular package and would win the
    # import regardless of path order, so take the repo root off the path while importing theirs
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    import src.models as rmodels
    assert rmodels.__file__.startswith(REF), rmodels.__file__
    out = {}
    for name, (conf, iou) in synth.POSTPROCESS_CASES.items():
        boxes, sims = synth.make_postprocess_inputs(name)
        pp = rmodels.PostProcess(confidence_threshold=conf, iou_threshold=iou)
        for b in range(boxes.shape[0]):
            ob, oc, os_ = pp(boxes[b:b + 1].clone(), sims[b:b + 1].clone())
            out[f"{name}_boxes{b}"] = ob[0].numpy().copy()
            out[f"{name}_classes{b}"] = oc[0].numpy().copy()
            out[f"{name}_scores{b}"] = os_[0].numpy().copy()
            print(name, b, "kept", ob.shape[1])
    np.savez_compressed(os.path.join(HERE, "postprocess.npz"), **out)


if __name__ == "__main__":
    main()
