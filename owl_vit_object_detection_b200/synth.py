"""Seeded synthetic weights and inputs for the OWL-ViT hot path.

There is no network on the build or GPU boxes, so neither pretrained weights nor COCO are
available.  Everything here is a pure function of a seed, so the oracle (run on the host),
the golden fixtures (generated once from the real reference, tests/golden/make_golden.py)
and the CUDA path (run on the GPU box) all see bit-identical fp32 inputs.

State-dict key names follow the reference wrapper `OwlViT` (reference src/models.py:48-61):
`queries`, `backbone.*`, `post_post_layernorm.*`, `class_predictor.dense0.*`, `box_head.dense{0,1,2}.*`.
"""
from __future__ import annotations

import dataclasses
import math
import zlib
from typing import Dict, List, Tuple

import torch


@dataclasses.dataclass(frozen=True)
class OwlConfig:
    """Dimensions of the vision tower + heads (defaults = google/owlvit-base-patch32 @ 768 px)."""
    image_size: int = 768
    patch_size: int = 32
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    ff: int = 3072
    embed: int = 512          # class-head projection dim == text embedding dim
    n_classes: int = 80
    variants: int = 3         # prompt variants per class (reference src/models.py:155-159)
    ln_eps: float = 1e-5

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def patches(self) -> int:
        return self.grid * self.grid

    @property
    def tokens(self) -> int:
        return self.patches + 1

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def n_queries(self) -> int:
        return self.n_classes * self.variants


B32 = OwlConfig()
# OWL-ViT-L/14 @ 840 (our extension, not a reference capability).
L14 = OwlConfig(image_size=840, patch_size=14, hidden=1024, layers=24, heads=16, ff=4096, embed=768)
# A tiny configuration for quick unit tests (same structure, every dimension a legal tile multiple).
TINY = OwlConfig(image_size=128, patch_size=32, hidden=128, layers=2, heads=2, ff=256, embed=128,
                 n_classes=8)


def param_shapes(cfg: OwlConfig) -> Dict[str, Tuple[int, ...]]:
    """Every parameter of the reference wrapper, in the reference's state_dict order."""
    D, F, E, P = cfg.hidden, cfg.ff, cfg.embed, cfg.patch_size
    s: Dict[str, Tuple[int, ...]] = {}
    s["queries"] = (1, cfg.n_queries, E)
    s["backbone.embeddings.class_embedding"] = (D,)
    s["backbone.embeddings.patch_embedding.weight"] = (D, 3, P, P)
    s["backbone.embeddings.position_embedding.weight"] = (cfg.tokens, D)
    s["backbone.pre_layernorm.weight"] = (D,)
    s["backbone.pre_layernorm.bias"] = (D,)
    for i in range(cfg.layers):
        p = f"backbone.encoder.layers.{i}."
        for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[p + f"self_attn.{proj}.weight"] = (D, D)
            s[p + f"self_attn.{proj}.bias"] = (D,)
        s[p + "layer_norm1.weight"] = (D,)
        s[p + "layer_norm1.bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (F, D)
        s[p + "mlp.fc1.bias"] = (F,)
        s[p + "mlp.fc2.weight"] = (D, F)
        s[p + "mlp.fc2.bias"] = (D,)
        s[p + "layer_norm2.weight"] = (D,)
        s[p + "layer_norm2.bias"] = (D,)
    s["backbone.post_layernorm.weight"] = (D,)
    s["backbone.post_layernorm.bias"] = (D,)
    s["post_post_layernorm.weight"] = (D,)
    s["post_post_layernorm.bias"] = (D,)
    s["class_predictor.dense0.weight"] = (E, D)
    s["class_predictor.dense0.bias"] = (E,)
    for j, out in ((0, D), (1, D), (2, 4)):
        s[f"box_head.dense{j}.weight"] = (out, D)
        s[f"box_head.dense{j}.bias"] = (out,)
    return s


def trainable_names(cfg: OwlConfig) -> List[str]:
    """The reference freeze rule (reference src/models.py:173-184), generalised from the literal
    substring "layers.11" to "the last encoder layer" so that it also makes sense for L/14 (D5)."""
    last = f"layers.{cfg.layers - 1}."
    out = []
    for name in param_shapes(cfg):
        if (last in name or "box" in name or "post_layernorm" in name
                or "class_predictor" in name or "queries" in name):
            out.append(name)
    return out


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def make_weights(cfg: OwlConfig = B32, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 CPU state dict.  Linear/conv weights ~ N(0, 1/fan_in) (so the heads do not saturate,
    SURVEY Q11), biases ~ N(0, 0.02), LayerNorm gamma ~ 1 + 0.1 N(0,1), beta ~ 0.05 N(0,1) so that
    every affine term is exercised by the parity tests (HF's own init leaves biases at zero)."""
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        g = _gen(seed, name)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if name == "queries":
            # HF text_embeds are unit-norm
            t = torch.nn.functional.normalize(r, dim=-1)
        elif name.endswith("class_embedding"):
            t = r * cfg.hidden ** -0.5
        elif name.endswith("position_embedding.weight"):
            t = r * 0.02
        elif name.endswith("patch_embedding.weight"):
            t = r * (3 * cfg.patch_size * cfg.patch_size) ** -0.5
        elif "layer_norm" in name or "layernorm" in name:
            t = 1.0 + 0.1 * r if name.endswith("weight") else 0.05 * r
        elif name.endswith(".weight"):
            t = r * shape[1] ** -0.5
        else:  # biases
            t = r * 0.02
        sd[name] = t.contiguous()
    return sd


def make_images(cfg: OwlConfig, batch: int, seed: int = 2) -> torch.Tensor:
    """`randn` clamped to the CLIP-normalised pixel range [-1.8, 2.2]."""
    g = _gen(seed, "images")
    x = torch.randn((batch, 3, cfg.image_size, cfg.image_size), generator=g, dtype=torch.float32)
    return x.clamp_(-1.8, 2.2)


def make_images_u8(cfg: OwlConfig, batch: int, seed: int = 2) -> torch.Tensor:
    """Raw RGB bytes [B, IS, IS, 3] (HWC, what an image decoder yields) at the model's resolution: the input of the
    device-side preprocessing path (reference src/dataset.py:64-71 does rescale + normalise on the CPU instead)."""
    g = _gen(seed, "images_u8")
    return torch.randint(0, 256, (batch, cfg.image_size, cfg.image_size, 3), generator=g, dtype=torch.uint8)


def make_targets(cfg: OwlConfig, batch: int, seed: int = 3, fixed_t: int | None = None,
                 max_t: int = 100) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """COCO-shaped targets.  Returns (labels [B,max_t] i64, boxes [B,max_t,4] f32 rel-xyxy,
    num_targets [B] i32); rows >= num_targets[b] are padding (label -1, box 0)."""
    g = _gen(seed, "targets")
    labels = torch.full((batch, max_t), -1, dtype=torch.int64)
    boxes = torch.zeros((batch, max_t, 4), dtype=torch.float32)
    nt = torch.zeros((batch,), dtype=torch.int32)
    for b in range(batch):
        if fixed_t is None:
            u = torch.rand((), generator=g).item()
            t = int(round(-7.3 * math.log(max(1.0 - u, 1e-12))))
            t = max(1, min(max_t, t))
        else:
            t = fixed_t
        cxy = 0.1 + 0.8 * torch.rand((t, 2), generator=g)
        wh = 0.02 + 0.48 * torch.rand((t, 2), generator=g)
        lo = (cxy - wh / 2).clamp(0.0, 1.0)
        hi = (cxy + wh / 2).clamp(0.0, 1.0)
        hi = torch.maximum(hi, lo + 1e-3)
        boxes[b, :t] = torch.cat([lo, hi], dim=-1)
        labels[b, :t] = torch.randint(0, cfg.n_classes, (t,), generator=g)
        nt[b] = t
    return labels, boxes, nt


def make_class_scales(cfg: OwlConfig, seed: int = 5) -> torch.Tensor:
    """`round(log(max_n / n) + 3, 1)` from a Zipf-like count vector (reference src/dataset.py:97-98)."""
    g = _gen(seed, "scales")
    counts = (1000.0 / (1.0 + torch.arange(cfg.n_classes, dtype=torch.float64))).round()
    counts = counts[torch.randperm(cfg.n_classes, generator=g)].clamp_(min=1)
    scales = torch.log(counts.max() / counts) + 3.0
    return (scales * 10).round().div(10).to(torch.float32)


def make_matcher_inputs(n_images: int, t: int, seed: int = 4, patches: int = 576, n_classes: int = 80):
    """Matcher microbench inputs (SURVEY §8d row "matcher bench"): sims ~ U(-0.1, 0.3), predicted
    boxes drawn like the targets, fixed T."""
    g = _gen(seed, f"matcher{t}")
    sims = torch.rand((n_images, patches, n_classes), generator=g) * 0.4 - 0.1

    def boxes(n):
        cxy = 0.1 + 0.8 * torch.rand((n_images, n, 2), generator=g)
        wh = 0.02 + 0.48 * torch.rand((n_images, n, 2), generator=g)
        lo = (cxy - wh / 2).clamp(0.0, 1.0)
        hi = torch.maximum((cxy + wh / 2).clamp(0.0, 1.0), lo + 1e-3)
        return torch.cat([lo, hi], dim=-1)

    pred = boxes(patches)
    tgt = boxes(t)
    labels = torch.randint(0, n_classes, (n_images, t), generator=g)
    return sims, pred, labels, tgt


def subsample(t: torch.Tensor) -> torch.Tensor:
    """Deterministic small view of a (gradient) tensor for golden fixtures: the full tensor when it
    has <= 4096 elements, else rows ::37 / cols ::29 of its [-1, last_dim] view."""
    if t.numel() <= 4096:
        return t
    t2 = t.reshape(-1, t.shape[-1]) if t.dim() > 1 else t.reshape(-1, 1)
    return t2[::37, ::29]


# ------------------------------------------------------------------------------------------ PostProcess cases
# name -> (confidence_threshold, iou_threshold); inputs from make_postprocess_inputs(name)
POSTPROCESS_CASES = {
    "dense": (0.01, 0.6),      # reference config.yaml thresholds: almost every prediction passes, NMS does the work
    "ties": (0.2, 0.5),        # quantised scores / duplicated boxes: stable-sort and first-maximum tie rules
    "sparse": (0.75, 0.3),     # PostProcess defaults: few or no survivors (image 2 has none)
}


def make_postprocess_inputs(name: str, n_images: int = 3, patches: int = 576, n_classes: int = 80, seed: int = 6):
    """(pred_boxes [B,P,4] xyxy in [0,1], pred_sims [B,P,C]) for the PostProcess parity cases: boxes jittered around
    a few dozen centres so that class-aware NMS has real work, a handful of dominant classes."""
    g = _gen(seed, "postprocess_" + name)
    centres = 0.15 + 0.7 * torch.rand((n_images, 40, 2), generator=g)
    sizes = 0.05 + 0.3 * torch.rand((n_images, 40, 2), generator=g)
    which = torch.randint(0, 40, (n_images, patches), generator=g)
    cxy = torch.gather(centres, 1, which[..., None].expand(-1, -1, 2)) + 0.02 * torch.randn((n_images, patches, 2), generator=g)
    wh = torch.gather(sizes, 1, which[..., None].expand(-1, -1, 2)) * (1 + 0.1 * torch.randn((n_images, patches, 2), generator=g))
    wh = wh.clamp_min(0.01)
    lo = (cxy - wh / 2).clamp(0.0, 1.0)
    hi = torch.maximum((cxy + wh / 2).clamp(0.0, 1.0), lo + 1e-3)
    boxes = torch.cat([lo, hi], dim=-1).float()
    sims = torch.rand((n_images, patches, n_classes), generator=g) * 0.2 - 0.1
    hot = torch.randint(0, 6, (n_images, patches), generator=g)                  # six dominant classes
    boost = torch.rand((n_images, patches), generator=g)
    sims.scatter_(2, hot[..., None], boost[..., None])
    if name == "ties":
        sims = torch.round(sims * 16) / 16
        boxes[:, 1::7] = boxes[:, 0:-1:7][:, : boxes[:, 1::7].shape[1]]          # exact duplicates
    if name == "sparse":
        sims[2] = sims[2].clamp_max(0.5)                                          # nothing above 0.75 in image 2
    return boxes.contiguous(), sims.float().contiguous()


# ------------------------------------------------------------------------------------------ preprocessing cases
# (height, width, output size): COCO-like down-scaling, up-scaling, identity on one axis, odd sizes, strong reduction
PREPROCESS_CASES = [(480, 640, 768), (427, 640, 768), (333, 500, 768), (768, 1024, 768), (1200, 1600, 768),
                    (97, 131, 768), (2001, 1503, 768), (64, 48, 96)]


def make_raw_image(h: int, w: int, seed: int = 0):
    """Synthetic uint8 RGB image [h, w, 3] (numpy): smooth gradients + blocks + noise, so that resampling sees both
    flat regions and edges."""
    import numpy as np
    rng = np.random.default_rng(1000 + seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([127 + 120 * np.sin(xx / (7.0 + seed)) * np.cos(yy / 11.0), 255.0 * xx / max(w - 1, 1),
                     255.0 * ((xx // 16 + yy // 16) % 2)], axis=-1)
    noise = rng.integers(-40, 41, size=(h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)
