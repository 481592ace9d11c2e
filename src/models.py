"""Mirror of reference src/models.py: `OwlViT`, `PostProcess`, `load_model` (same names and signatures)."""
import os

import torch

from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT  # noqa: F401
from owl_vit_object_detection_b200.synth import trainable_names
from owl_vit_object_detection_b200.text import text_query_bank


class PostProcess:
    """reference src/models.py:122-146 (eval path, main.py:110-118): best class per prediction, confidence threshold,
    class-aware NMS, survivors in decreasing-score order - one `owl_postprocess` launch (CTA per image) instead of
    boolean-mask indexing and torchvision's batched_nms.  Same call signature and return shapes as the reference
    (batch 1: boxes [1,K,4], classes [1,K] i64, scores [1,K]); `batched` returns the padded per-image results."""

    def __init__(self, confidence_threshold=0.75, iou_threshold=0.3):
        self.confidence_threshold = confidence_threshold
        self.iou_threshold = iou_threshold

    def batched(self, all_pred_boxes, pred_classes):
        """[B,P,4], [B,P,C] -> (boxes [B,P,4], classes [B,P], scores [B,P], count [B]) on the device, no sync."""
        from owl_vit_object_detection_b200 import ops
        if not all_pred_boxes.is_cuda:
            raise RuntimeError("PostProcess runs on the CUDA device only (there is no CPU fallback)")
        return ops.postprocess(all_pred_boxes.detach().float().contiguous(), pred_classes.detach().float().contiguous(),
                               self.confidence_threshold, self.iou_threshold)

    def __call__(self, all_pred_boxes, pred_classes):
        assert all_pred_boxes.shape[0] == 1, "the reference PostProcess is batch-1 (use .batched for more)"
        boxes, classes, scores, count = self.batched(all_pred_boxes, pred_classes)
        k = int(count[0].item())            # the reference syncs here too (boolean-mask indexing)
        return boxes[:, :k], classes[:, :k], scores[:, :k]


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
    # reference src/models.py:165-169 `queries = _model(**inputs).text_embeds`: the text tower on the device kernels
    # (owl_vit_object_detection_b200/text.py); [1, 3 * classes, 512], unit-norm rows
    queries = text_query_bank(hf, inputs["input_ids"], inputs.get("attention_mask"), device)
    model = OwlViT(pretrained_model=hf, query_bank=queries)
    keep = set(trainable_names(model.cfg))
    print("Trainable parameters:")
    for pname, p in model.named_parameters():
        p.requires_grad = pname in keep
        if p.requires_grad:
            print(f"  {pname}")
    print()
    return model.to(device)
