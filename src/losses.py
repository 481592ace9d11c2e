"""Mirror of reference src/losses.py: `PushPullLoss(n_classes, scales)`."""
from owl_vit_object_detection_b200.loss import HungarianMatcher, PushPullLoss  # noqa: F401
