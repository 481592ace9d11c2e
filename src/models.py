"""Mirror of reference src/models.py: `OwlViT`, `PostProcess`, `load_model` (same names and signatures)."""
import os

import torch

from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT  # noqa: F401
from owl_vit_object_detection_b200.synth import trainable_names


class PostProcess:
    """reference src/models.py:122-146 (eval only, batch 1; SURVEY §8f row N2 — not on the accelerated path):
    best class per prediction, confidence threshold, class-aware NMS."""

    def __init__(self, confidence_threshold=0.75, iou_threshold=0.3):
        self.confidence_threshold = confidence_threshold
        self.iou_threshold = iou_threshold

    def __call__(self, all_pred_boxes, pred_classes):
        from torchvision.ops import batched_nms
        boxes, sims = all_pred_boxes[0], pred_classes[0]
        scores, classes = sims.max(dim=1)
        keep = scores > self.confidence_threshold
        boxes, scores, classes = boxes[keep], scores[keep], classes[keep]
        keep = batched_nms(boxes, scores, classes, iou_threshold=self.iou_threshold)
        return boxes[keep][None], classes[keep][None], scores[keep][None]


def load_model(labelmap, device):
    """reference src/models.py:149-191: HF owlvit-base-patch32 weights, query bank seeded from three prompts per
    class through the text tower (once), reference freeze rule, moved to `device`.  Needs the HF hub (network or
    cache), exactly like the reference."""
    from PIL import Image
    from transformers import AutoProcessor, OwlViTForObjectDetection
    os.environ["TOKENIZERS_PARALLELISM"] = "false"
    name = "google/owlvit-base-patch32"
    hf = OwlViTForObjectDetection.from_pretrained(name)
    processor = AutoProcessor.from_pretrained(name)
    prompts = []
    for label in labelmap.values():
        prompts += [label, "a photo of " + label, "a " + label + " in an environment"]
    print("Initializing priors from labels...")
    inputs = processor(text=[prompts], images=Image.new("RGB", (224, 224)), return_tensors="pt")
    with torch.no_grad():
        queries = hf(**inputs).text_embeds
    model = OwlViT(pretrained_model=hf, query_bank=queries)
    keep = set(trainable_names(model.cfg))
    print("Trainable parameters:")
    for pname, p in model.named_parameters():
        p.requires_grad = pname in keep
        if p.requires_grad:
            print(f"  {pname}")
    print()
    return model.to(device)
