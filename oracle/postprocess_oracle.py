"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference `PostProcess`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs may import this module.  Restates

  * reference `src/models.py:127-146`  PostProcess.__call__: best class per prediction (torch.max over classes,
    first maximum wins), `scores > confidence_threshold` (threshold rounded to fp32, as torch compares a float32
    tensor with a Python scalar), class-aware NMS, outputs in decreasing-score order;
  * torchvision 0.26 `ops.boxes.batched_nms` -> `_batched_nms_coordinate_trick` (taken for <= 1000 boxes on the
    CPU): boxes + class * (max coordinate + 1), then `nms`;
  * torchvision's CPU `nms` kernel (csrc/ops/cpu/nms_kernel.cpp, source not in this container; restated from its
    published algorithm): stable descending sort of the scores, areas = (x2 - x1) * (y2 - y1), greedy sweep with
    ovr = inter / (area_i + area_j - inter), suppressed when ovr > iou_threshold (the threshold is a double: the
    fp32 ovr is promoted).  Every arithmetic step is a separately rounded fp32 operation (numpy float32 scalars).

Parity pin: tests/golden/postprocess.npz holds outputs of the REAL reference class run on the CPU
(tests/golden/make_golden_postprocess.py); tests/test_oracle_postprocess.py checks this file against them.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def postprocess_image(boxes: np.ndarray, sims: np.ndarray, confidence_threshold: float, iou_threshold: float
                      ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """boxes [P,4] f32 xyxy, sims [P,C] f32 -> (boxes [K,4] f32, classes [K] i64, scores [K] f32)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    sims = np.asarray(sims, dtype=np.float32)
    classes = sims.argmax(axis=1).astype(np.int64)            # first maximum, like torch.max(dim=1) on the CPU
    scores = sims[np.arange(sims.shape[0]), classes]
    keep = scores > np.float32(confidence_threshold)
    b, s, c = boxes[keep], scores[keep], classes[keep]
    n = b.shape[0]
    if n == 0:
        return b.reshape(0, 4), c, s
    one = np.float32(1)
    maxc = np.float32(b.max())
    off = c.astype(np.float32) * np.float32(maxc + one)
    ob = (b + off[:, None]).astype(np.float32)
    x1, y1, x2, y2 = ob[:, 0], ob[:, 1], ob[:, 2], ob[:, 3]
    areas = ((x2 - x1).astype(np.float32) * (y2 - y1).astype(np.float32)).astype(np.float32)
    order = np.argsort(-s, kind="stable")                      # stable, descending
    suppressed = np.zeros(n, dtype=bool)
    kept = []
    zero = np.float32(0)
    thr = float(iou_threshold)
    for a in range(n):
        i = order[a]
        if suppressed[i]:
            continue
        kept.append(i)
        for bb in range(a + 1, n):
            j = order[bb]
            if suppressed[j]:
                continue
            xx1, yy1 = max(x1[i], x1[j]), max(y1[i], y1[j])
            xx2, yy2 = min(x2[i], x2[j]), min(y2[i], y2[j])
            w = max(zero, np.float32(xx2 - xx1))
            h = max(zero, np.float32(yy2 - yy1))
            inter = np.float32(w * h)
            ovr = np.float32(inter / np.float32(np.float32(areas[i] + areas[j]) - inter))
            if float(ovr) > thr:
                suppressed[j] = True
    kept = np.asarray(kept, dtype=np.int64)
    return b[kept], c[kept], s[kept]
