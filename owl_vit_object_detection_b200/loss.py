"""`HungarianMatcher` and `PushPullLoss` — the reference's matcher and loss (reference src/matcher.py:47-159,
src/losses.py:9-116) on device-resident kernels: no `.cpu()` round trip, no SciPy, no 576-iteration Python
loop, no per-step host synchronisation.

Batching: the reference supports batch size 1 only.  Here a batch of B images is scored as the MEAN
over images of the reference's per-image loss; targets of different lengths are padded with label -1 (or pass
`num_targets`).  With B = 1 and unpadded inputs the call is exactly the reference's.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops


class _Buffers:
    """Per-(B, P, C, Tmax) device buffers, reused across steps."""

    def __init__(self, B: int, P: int, C: int, Tmax: int, dev):
        i64, i32, f32 = torch.int64, torch.int32, torch.float32
        self.generation = 0                  # bumped by every HungarianMatcher.assign that writes these buffers
        self.costT = torch.zeros((B, Tmax, P), dtype=f32, device=dev)
        self.status = torch.zeros(1, dtype=i32, device=dev)
        self.match = torch.zeros((B, Tmax), dtype=i32, device=dev)
        self.tc_matched = torch.zeros((B, P), dtype=i64, device=dev)
        self.tc_final = torch.zeros((B, P), dtype=i64, device=dev)
        self.pred_sorted = torch.zeros((B, Tmax), dtype=i64, device=dev)
        self.tgt_sorted = torch.zeros((B, Tmax), dtype=i64, device=dev)
        self.losses_per_image = torch.zeros((B, ops.LOSS_WS), dtype=f32, device=dev)   # [:, :4] = the four losses
        self.dsims_unit = torch.zeros((B, P, C), dtype=f32, device=dev)
        self.dl1 = torch.zeros((B, Tmax, 4), dtype=f32, device=dev)
        self.dgiou = torch.zeros((B, Tmax, 4), dtype=f32, device=dev)


def _raise_status(status: int) -> None:
    """The matcher kernels flag what the reference raises inline (they cannot raise from the device)."""
    if status & 1:
        raise AssertionError("degenerate box: x1 < x0 or y1 < y0 (reference src/matcher.py:34-35)")
    if status & 8:
        raise IndexError("target label outside [0, n_classes) (reference src/matcher.py:118 gather)")
    if status & 4:
        raise ValueError("num_targets outside [0, Tmax]")
    if status & 2:
        raise ValueError("cost matrix is infeasible")


def _pad_targets(labels, boxes, num_targets, dev) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Accepts the reference's [B,T] / [B,T,4] tensors (optionally -1-padded) or per-image lists."""
    if isinstance(labels, (list, tuple)):
        B = len(labels)
        Tmax = max(1, max(int(l.shape[0]) for l in labels))
        lab = torch.full((B, Tmax), -1, dtype=torch.int64, device=dev)
        box = torch.zeros((B, Tmax, 4), dtype=torch.float32, device=dev)
        nt = torch.tensor([int(l.shape[0]) for l in labels], dtype=torch.int32, device=dev)
        for b in range(B):
            t = int(labels[b].shape[0])
            lab[b, :t] = labels[b].to(dev)
            box[b, :t] = boxes[b].to(dev)
        return lab, box, nt
    lab = labels.to(device=dev, dtype=torch.int64).contiguous()
    box = boxes.to(device=dev, dtype=torch.float32).contiguous()
    if lab.dim() == 1:
        lab, box = lab[None], box[None]
    if num_targets is None:
        # -1 padding is trailing: the targets of an image are the labels before its first negative entry
        nt = (lab >= 0).to(torch.int32).cumprod(dim=1).sum(dim=1).to(torch.int32)
    else:
        nt = num_targets.to(device=dev, dtype=torch.int32)
    return lab, box, nt.contiguous()


class HungarianMatcher(nn.Module):
    """reference src/matcher.py:47-159.  `forward(outputs, targets)` returns
    (target_classes [B,P] i64 on the prediction device, indices = list of (pred_idx, tgt_idx) int64 CPU tensors
    sorted by prediction index, idx = (batch_idx, src_idx))."""

    def __init__(self, n_classes: int, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"    # src/matcher.py:62-64
        self.cost_class, self.cost_bbox, self.cost_giou = float(cost_class), float(cost_bbox), float(cost_giou)
        self.n_classes = n_classes
        self._buf: Dict[tuple, _Buffers] = {}

    def buffers(self, B, P, C, Tmax, dev) -> _Buffers:
        key = (B, P, C, Tmax, str(dev))
        if key not in self._buf:
            self._buf[key] = _Buffers(B, P, C, Tmax, dev)
        return self._buf[key]

    @torch.no_grad()
    def assign(self, sims, boxes, lab, box, nt) -> _Buffers:
        """Cost matrix + LSAP on the device; results stay in the returned buffers (no host sync)."""
        B, P, C = sims.shape
        buf = self.buffers(B, P, C, lab.shape[1], sims.device)
        buf.generation += 1
        ops.zero(buf.status)
        with torch.cuda.device(sims.device):
            ops.matcher_cost(sims, boxes, lab, box, nt, buf.costT, buf.status, self.cost_class, self.cost_bbox,
                             self.cost_giou)
            ops.lsap(buf.costT, nt, buf.match, buf.status)
        return buf

    @torch.no_grad()
    def forward(self, outputs, targets):
        sims = outputs["pred_logits"].detach().float().contiguous()
        boxes = outputs["pred_boxes"].detach().float().contiguous()
        dev = sims.device
        lab, box, nt = _pad_targets([t["labels"] for t in targets], [t["boxes"] for t in targets], None, dev)
        buf = self.assign(sims, boxes, lab, box, nt)
        B, P, C = sims.shape
        ops.match_loss(sims, boxes, lab, box, nt, buf.match, None, self.n_classes, tc_matched=buf.tc_matched,
                       tc_final=buf.tc_final, pred_sorted=buf.pred_sorted, tgt_sorted=buf.tgt_sorted,
                       losses_per_image=buf.losses_per_image, losses_mean4=torch.zeros(4, device=dev),
                       dsims_unit=buf.dsims_unit, dl1=buf.dl1, dgiou=buf.dgiou)
        # this entry point returns host-side index lists like the reference does, so it synchronises here
        _raise_status(int(buf.status.item()))
        ps, ts, ntc = buf.pred_sorted.cpu(), buf.tgt_sorted.cpu(), nt.cpu()
        indices = [(ps[b, :int(ntc[b])].clone(), ts[b, :int(ntc[b])].clone()) for b in range(B)]
        batch_idx = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        src_idx = torch.cat([src for (src, _) in indices])
        return buf.tc_matched.clone(), indices, (batch_idx, src_idx)


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sims, boxes, module, lab, box, nt):
        B, P, C = sims.shape
        sims_c, boxes_c = sims.detach().float().contiguous(), boxes.detach().float().contiguous()
        buf = module.matcher.assign(sims_c, boxes_c, lab, box, nt)
        losses4 = torch.empty(4, dtype=torch.float32, device=sims.device)
        ops.match_loss(sims_c, boxes_c, lab, box, nt, buf.match, module.scales, module.background_label,
                       tc_matched=buf.tc_matched, tc_final=buf.tc_final, pred_sorted=buf.pred_sorted,
                       tgt_sorted=buf.tgt_sorted, losses_per_image=buf.losses_per_image, losses_mean4=losses4,
                       dsims_unit=buf.dsims_unit, dl1=buf.dl1, dgiou=buf.dgiou)
        ctx.buf, ctx.module, ctx.shape, ctx.generation = buf, module, (B, P, C), buf.generation
        return losses4

    @staticmethod
    def backward(ctx, g4):
        buf, module = ctx.buf, ctx.module
        if buf.generation != ctx.generation:
            raise RuntimeError(
                "PushPullLoss.backward: the criterion (or its matcher) was called again with the same shapes before "
                "this backward; the saved matcher / loss buffers were overwritten.  Call backward() first.")
        B, P, C = ctx.shape
        dev = g4.device
        dsims = torch.empty((B, P, C), dtype=torch.float32, device=dev)
        dboxes = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
        ops.loss_backward(buf.dsims_unit, buf.tc_final, buf.match, buf.dl1, buf.dgiou, g4.contiguous().float(),
                          module.background_label, dsims, dboxes)
        return dsims, dboxes, None, None, None, None


class PushPullLoss(nn.Module):
    """reference src/losses.py:9-116: `PushPullLoss(n_classes, scales)(pred_sims, labels, pred_boxes, boxes)`
    -> {"loss_ce", "loss_bg", "loss_bbox", "loss_giou"} (0-dim tensors with grad)."""

    def __init__(self, n_classes: int, scales=None):
        super().__init__()
        self.matcher = HungarianMatcher(n_classes)
        self.background_label = n_classes
        if scales is not None and not torch.is_tensor(scales):
            scales = torch.tensor(scales, dtype=torch.float32)
        self.register_buffer("scales", None if scales is None else scales.detach().float().contiguous())

    def forward(self, predicted_classes, target_classes, predicted_boxes, target_boxes, num_targets=None):
        dev = predicted_classes.device
        if self.scales is not None and self.scales.device != dev:
            self.scales = self.scales.to(dev)
        lab, box, nt = _pad_targets(target_classes, target_boxes, num_targets, dev)
        losses4 = _LossFn.apply(predicted_classes, predicted_boxes, self, lab, box, nt)
        return {"loss_ce": losses4[0], "loss_bg": losses4[1], "loss_bbox": losses4[2], "loss_giou": losses4[3]}

    def check_status(self) -> None:
        """Raises what the reference would have raised inline (degenerate boxes).  Synchronises; the train loop
        calls it where it synchronises anyway (e.g. next to `.item()` on the losses)."""
        for buf in self.matcher._buf.values():
            _raise_status(int(buf.status.item()))
