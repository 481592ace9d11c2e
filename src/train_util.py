"""Mirror of reference src/train_util.py (five helpers, same signatures; glue, not on the accelerated path)."""
from src.util import BoxUtil


def coco_to_model_input(boxes, metadata):
    """absolute xywh -> relative xyxy (reference src/train_util.py:4-13)."""
    boxes = BoxUtil.box_convert(boxes, "xywh", "xyxy")
    return BoxUtil.scale_bounding_box(boxes, metadata["width"], metadata["height"], mode="down")


def model_output_to_image(boxes, metadata):
    """relative xyxy -> absolute pixels, in place (reference src/train_util.py:16-23)."""
    return BoxUtil.scale_bounding_box(boxes, metadata["width"], metadata["height"], mode="up")


def reverse_labelmap(labelmap):
    return {v["new_idx"]: {"actual_category": k, "name": v["name"]} for k, v in labelmap.items()}


def labels_to_classnames(labels, labelmap):
    return [[labelmap[str(int(l))] for l in labels[0]]]


def update_metrics(metric, metadata, pred_boxes, pred_classes, scores, boxes, labels):
    """Feeds torchmetrics' MeanAveragePrecision (reference src/train_util.py:37-64)."""
    w, h = metadata["width"], metadata["height"]
    pred_boxes = BoxUtil.scale_bounding_box(pred_boxes.cpu(), w, h, mode="up")
    boxes = BoxUtil.scale_bounding_box(boxes.cpu(), w, h, mode="up")
    preds = [{"boxes": b.cuda(), "scores": s.cuda(), "labels": c.cuda()}
             for b, c, s in zip(pred_boxes, pred_classes, scores)]
    targets = [{"boxes": b.cuda(), "labels": c.cuda()} for b, c in zip(boxes, labels)]
    metric.update(preds, targets)
