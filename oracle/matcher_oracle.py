"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the reference matcher + loss.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  Restates, function by function:

  * reference `src/matcher.py:8-21`    box_iou            -> `box_iou_union`
  * reference `src/matcher.py:25-44`   generalized_box_iou -> `generalized_box_iou`
  * reference `src/matcher.py:103-131` cost matrix         -> `cost_matrix`
  * reference `src/matcher.py:134-159` assignment + target_classes -> `hungarian`
    (the solver: oracle/lsap.c restating SciPy's rectangular LSAP, see its header)
  * reference `src/losses.py:100-106`  IoU>0.85 label propagation -> `propagate_labels`
  * reference `src/losses.py:42-69`    loss_boxes          -> inside `push_pull_loss_image`
  * reference `src/losses.py:16-40`    class_loss          -> inside `push_pull_loss_image`
  * reference `src/losses.py:71-116`   PushPullLoss.forward -> `push_pull_loss`

fp32 torch CPU ops are used where the reference uses fp32 torch ops, in the same operation order
(every op rounds separately, which is what the CUDA kernels mirror with non-contracted arithmetic).
The reference is batch-1 only; the batched loss is DEFINED as the mean over images of the
reference's per-image loss, which is how `push_pull_loss` evaluates it.

Parity pin: tests/golden/matcher_*.npz and loss_*.npz are outputs of the real reference classes
(tests/golden/make_golden.py); tests/test_oracle_matcher.py checks this file against them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lsap_lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "liblsap_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["sh", os.path.join(_HERE, "build_oracle.sh")])
        lib = ctypes.CDLL(so)
        lib.lsap_oracle_f32.restype = ctypes.c_int
        lib.lsap_oracle_f32.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]
        _LIB = lib
    return _LIB


def lsap(cost: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """`scipy.optimize.linear_sum_assignment` semantics on a float32 [R, C] matrix (oracle/lsap.c)."""
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    r, c = cost.shape
    k = min(r, c)
    rows = np.empty(k, dtype=np.int64)
    cols = np.empty(k, dtype=np.int64)
    n = _lsap_lib().lsap_oracle_f32(r, c, cost.ctypes.data, rows.ctypes.data, cols.ctypes.data)
    if n < 0:
        raise ValueError("cost matrix is infeasible")
    return rows[:n], cols[:n]


def box_area(b: torch.Tensor) -> torch.Tensor:
    # torchvision.ops.box_area: (x2 - x1) * (y2 - y1)
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def box_iou_union(b1: torch.Tensor, b2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    # reference src/matcher.py:8-21
    a1, a2 = box_area(b1), box_area(b2)
    lt = torch.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = torch.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2[None, :] - inter
    return inter / union, union


def generalized_box_iou(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    # reference src/matcher.py:25-44 (degenerate boxes assert at :34-35)
    if not bool((b1[:, 2:] >= b1[:, :2]).all()) or not bool((b2[:, 2:] >= b2[:, :2]).all()):
        raise AssertionError("degenerate box (x1 < x0 or y1 < y0)")
    iou, union = box_iou_union(b1, b2)
    lt = torch.minimum(b1[:, None, :2], b2[None, :, :2])
    rb = torch.maximum(b1[:, None, 2:], b2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[..., 0] * wh[..., 1]
    return iou - (area - union) / area


def cost_matrix(sims: torch.Tensor, boxes: torch.Tensor, labels: torch.Tensor,
                tboxes: torch.Tensor, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1) -> torch.Tensor:
    """One image: sims [P,C] f32, boxes [P,4] xyxy, labels [T] i64, tboxes [T,4] -> C [P,T] f32.
    reference src/matcher.py:106-131 with the three weights = 1 (:58-60).  The L1 term is the
    sequential sum ((|dx0|+|dy0|)+|dx1|)+|dy1|, bit-identical to torch.cdist(p=1) on CPU.
    Sum order: (L1 + (-p)) + (-giou)  (:127-131)."""
    prob = sims.float().softmax(-1)
    d = (boxes[:, None, :] - tboxes[None, :, :]).abs()
    cost_bbox_m = ((d[..., 0] + d[..., 1]) + d[..., 2]) + d[..., 3]
    cost_giou_m = -generalized_box_iou(boxes, tboxes)
    w_class, w_bbox, w_giou = cost_class, cost_bbox, cost_giou
    cost_class = -prob[:, labels]
    # reference src/matcher.py:127-131: self.cost_bbox * cost_bbox + self.cost_class * cost_class + self.cost_giou * cost_giou
    return (w_bbox * cost_bbox_m + w_class * cost_class) + w_giou * cost_giou_m


def hungarian(sims: torch.Tensor, boxes: torch.Tensor, labels: Sequence[torch.Tensor],
              tboxes: Sequence[torch.Tensor], n_classes: int):
    """reference src/matcher.py:85-159 per image.  Returns (target_classes [B,P] i64,
    indices = list of (pred_idx i64 sorted asc, tgt_idx i64))."""
    B, P = sims.shape[:2]
    target_classes = torch.full((B, P), n_classes, dtype=torch.int64)
    indices = []
    for b in range(B):
        c = cost_matrix(sims[b], boxes[b], labels[b], tboxes[b])
        r, k = lsap(c.numpy())
        r, k = torch.from_numpy(r), torch.from_numpy(k)
        indices.append((r, k))
        target_classes[b, r] = labels[b][k]
    return target_classes, indices


def propagate_labels(tc: torch.Tensor, boxes: torch.Tensor, bg: int, thr: float = 0.85) -> torch.Tensor:
    """reference src/losses.py:100-106 — index-ordered sweep that mutates `tc` while reading it
    (labels written in iteration i are visible to iterations > i, never backwards)."""
    tc = tc.clone()
    for i in range(tc.shape[0]):
        lab = int(tc[i])
        if lab == bg:
            continue
        iou, _ = box_iou_union(boxes[i:i + 1], boxes)
        tc[iou[0] > thr] = lab
    return tc


def _weighted_bce(p: torch.Tensor, y: torch.Tensor, w: Optional[torch.Tensor]) -> torch.Tensor:
    # torch.nn.BCELoss(reduction="none", weight=w): -(w) * (y*max(log p,-100) + (1-y)*max(log(1-p),-100))
    lp = torch.log(p).clamp(min=-100.0)
    l1p = torch.log(1.0 - p).clamp(min=-100.0)
    loss = -(y * lp + (1.0 - y) * l1p)
    return loss if w is None else loss * w


def push_pull_loss_image(sims: torch.Tensor, boxes: torch.Tensor, labels: torch.Tensor,
                         tboxes: torch.Tensor, n_classes: int, scales: Optional[torch.Tensor]):
    """One image (the reference's only supported case).  sims [P,C], boxes [P,4] may require grad.
    Returns (dict of 4 scalars, target_classes after propagation [P], (pred_idx, tgt_idx))."""
    with torch.no_grad():
        tc, ind = hungarian(sims[None].detach(), boxes[None].detach(), [labels], [tboxes], n_classes)
        tc = tc[0]
        pi, ti = ind[0]
    # loss_boxes, reference src/losses.py:42-69
    src = boxes[pi]
    tgt = tboxes[ti]
    num_boxes = labels.shape[0]
    loss_bbox = (src - tgt).abs().sum() / num_boxes
    loss_giou = (1.0 - torch.diag(generalized_box_iou(src, tgt))).sum() / num_boxes
    # label propagation, reference src/losses.py:100-106
    with torch.no_grad():
        tc = propagate_labels(tc, boxes.detach(), n_classes)
    # class_loss, reference src/losses.py:16-40
    p = sims.abs()
    pos = tc != n_classes
    y = torch.nn.functional.one_hot(tc[pos], n_classes).to(p.dtype)
    pl = _weighted_bce(p[pos], y, scales)
    nl = _weighted_bce(p[~pos], torch.zeros_like(p[~pos]), scales)
    loss_ce = ((1.0 - torch.exp(-pl)) ** 2 * pl).sum(dim=1).mean()
    loss_bg = ((1.0 - torch.exp(-nl)) ** 2 * nl).sum(dim=1).mean()
    return ({"loss_ce": loss_ce, "loss_bg": loss_bg, "loss_bbox": loss_bbox, "loss_giou": loss_giou},
            tc, (pi, ti))


def push_pull_loss(sims: torch.Tensor, boxes: torch.Tensor, labels: Sequence[torch.Tensor],
                   tboxes: Sequence[torch.Tensor], n_classes: int, scales: Optional[torch.Tensor]):
    """Batched definition: mean over images of the reference per-image loss."""
    B = sims.shape[0]
    tot: Dict[str, torch.Tensor] = {}
    tcs, inds = [], []
    for b in range(B):
        l, tc, ind = push_pull_loss_image(sims[b], boxes[b], labels[b], tboxes[b], n_classes, scales)
        for k, v in l.items():
            tot[k] = v if k not in tot else tot[k] + v
        tcs.append(tc)
        inds.append(ind)
    return {k: v / B for k, v in tot.items()}, torch.stack(tcs), inds
