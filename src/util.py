"""Mirror of reference src/util.py: `BoxUtil`, `GeneralLossAccumulator`, `ProgressFormatter` (logging glue, not on
the accelerated path; same public behaviour incl. the in-place box scaling callers rely on)."""
import time
from collections import defaultdict
from datetime import timedelta

import torch


class GeneralLossAccumulator:
    def __init__(self):
        self.loss_values = defaultdict(float)
        self.n = 0

    def update(self, losses):
        # one device->host read for all four losses instead of four .item() syncs (reference src/util.py:21)
        keys = list(losses)
        vals = torch.stack([losses[k].detach().float() for k in keys]).tolist()
        for k, v in zip(keys, vals):
            self.loss_values[k] += v
        self.n += 1

    def get_values(self):
        return {k: round(v / self.n, 5) for k, v in self.loss_values.items()}

    def reset(self):
        # reference src/util.py:30-31 resets nothing that is read again: running averages continue
        self.value = 0


class ProgressFormatter:
    COLUMNS = ("epoch", "class loss", "bg loss", "box loss", "map", "map@0.5", "map (L/M/S)", "mar (L/M/S)",
               "time elapsed")

    def __init__(self):
        self.table = {c: [] for c in self.COLUMNS}
        self.start = time.time()

    def update(self, epoch, train_metrics, val_metrics):
        t = self.table
        t["epoch"].append(epoch)
        t["class loss"].append(train_metrics["loss_ce"])
        t["bg loss"].append(train_metrics["loss_bg"])
        t["box loss"].append(train_metrics["loss_bbox"] + train_metrics["loss_giou"])
        t["map"].append(round(val_metrics["map"].item(), 3))
        t["map@0.5"].append(round(val_metrics["map_50"].item(), 3))
        lms = lambda p: "/".join(str(round(val_metrics[f"{p}_{s}"].item(), 2)) for s in ("large", "medium", "small"))
        t["map (L/M/S)"].append(lms("map"))
        t["mar (L/M/S)"].append(lms("mar"))
        t["time elapsed"].append(str(timedelta(seconds=int(time.time() - self.start))))

    def print(self):
        from tabulate import tabulate
        print()
        print(tabulate(self.table, headers="keys"))
        print()


class BoxUtil:
    @classmethod
    def scale_bounding_box(cls, boxes_batch, imwidth, imheight, mode):
        """In place, like the reference (src/util.py:83-96): boxes [M,N,4] xyxy, mode "up" | "down"."""
        if mode == "down":
            boxes_batch[:, :, (0, 2)] /= imwidth
            boxes_batch[:, :, (1, 3)] /= imheight
        elif mode == "up":
            boxes_batch[:, :, (0, 2)] *= imwidth
            boxes_batch[:, :, (1, 3)] *= imheight
        else:
            return None
        return boxes_batch

    @classmethod
    def box_convert(cls, boxes_batch, in_format, out_format):
        from torchvision.ops import box_convert
        return box_convert(boxes_batch, in_format, out_format)

    @classmethod
    def draw_box_on_image(cls, image, boxes_batch, labels_batch=None, color=(0, 255, 0)):
        from torchvision.io import read_image
        from torchvision.utils import draw_bounding_boxes
        if isinstance(image, str):
            image = read_image(image)
        labels_iter = labels_batch if labels_batch is not None else [None] * len(boxes_batch)
        for boxes, labels in zip(boxes_batch, labels_iter):
            if len(boxes):
                image = draw_bounding_boxes(image, boxes, labels, width=2)
        return image
