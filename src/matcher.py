"""Mirror of reference src/matcher.py: `HungarianMatcher`, `box_iou`, `generalized_box_iou`."""
import torch

from owl_vit_object_detection_b200.loss import HungarianMatcher  # noqa: F401


def box_iou(boxes1, boxes2):
    """reference src/matcher.py:8-21 -> (iou [N,M], union [N,M]).  Host-side utility for callers; the train path
    computes these inside the matcher / loss kernels."""
    a1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    a2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    wh = (torch.minimum(boxes1[:, None, 2:], boxes2[None, :, 2:])
          - torch.maximum(boxes1[:, None, :2], boxes2[None, :, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2[None, :] - inter
    return inter / union, union


def generalized_box_iou(boxes1, boxes2):
    """reference src/matcher.py:25-44 (asserts on degenerate boxes like the reference)."""
    assert (boxes1[:, 2:] >= boxes1[:, :2]).all()
    assert (boxes2[:, 2:] >= boxes2[:, :2]).all()
    iou, union = box_iou(boxes1, boxes2)
    wh = (torch.maximum(boxes1[:, None, 2:], boxes2[None, :, 2:])
          - torch.minimum(boxes1[:, None, :2], boxes2[None, :, :2])).clamp(min=0)
    hull = wh[..., 0] * wh[..., 1]
    return iou - (hull - union) / hull
