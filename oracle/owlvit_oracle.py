"""ORACLE (test infrastructure, NOT product code) — fp32 CPU restatement of the reference forward.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product path (owl_vit_object_detection_b200, src/) never does.

This is a floating-point kernel, so per the task contract the oracle is a plain torch fp32
functional restatement, op for op, of

  * reference `src/models.py:98-119`  OwlViT.forward
  * reference `src/models.py:77-96`   OwlViT.image_embedder
  * reference `src/models.py:65-73`   OwlViT.box_predictor
  * reference `src/models.py:24-38`   PatchedOwlViTClassPredictionHead.forward
  * the third-party HuggingFace vision tower those call (transformers==4.30.2 pinned by the reference's
    requirements.txt:1; transformers 5.5.0 installed here — "HF:" line numbers below refer to the
    installed transformers/models/owlvit/modeling_owlvit.py, as SURVEY.md does).

Parity pin: tests/golden/*.npz hold outputs of the REAL reference classes (imported from
/root/reference with the SURVEY D7 shim) on the seeded weights of owl_vit_object_detection_b200.synth;
tests/test_oracle_model.py checks this restatement against them (generator: tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    # transformers/activations.py:122-123  x * sigmoid(1.702 x)
    return x * torch.sigmoid(1.702 * x)


def embeddings(sd: Dict[str, torch.Tensor], cfg, pixel_values: torch.Tensor) -> torch.Tensor:
    # HF:334-344  conv k=s=patch (no bias) -> flatten -> CLS cat -> + position table
    w = sd["backbone.embeddings.patch_embedding.weight"]
    x = F.conv2d(pixel_values, w, bias=None, stride=cfg.patch_size)          # [B,D,g,g]
    x = x.flatten(2).transpose(1, 2)                                          # [B,P,D]
    cls = sd["backbone.embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1)                                            # [B,S,D]
    return x + sd["backbone.embeddings.position_embedding.weight"][None]


def layer_norm(sd, prefix: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def attention(sd, prefix: str, cfg, x: torch.Tensor) -> torch.Tensor:
    # HF:430-461 (projections, head split) + HF:379-404 (eager: scale q.k^T, fp32 softmax, @v)
    B, S, D = x.shape
    H, dh = cfg.heads, cfg.head_dim
    q = F.linear(x, sd[prefix + "q_proj.weight"], sd[prefix + "q_proj.bias"])
    k = F.linear(x, sd[prefix + "k_proj.weight"], sd[prefix + "k_proj.bias"])
    v = F.linear(x, sd[prefix + "v_proj.weight"], sd[prefix + "v_proj.bias"])
    q = q.view(B, S, H, dh).transpose(1, 2)
    k = k.view(B, S, H, dh).transpose(1, 2)
    v = v.view(B, S, H, dh).transpose(1, 2)
    att = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5)
    att = torch.softmax(att, dim=-1, dtype=torch.float32)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, S, D)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def encoder_layer(sd, i: int, cfg, x: torch.Tensor) -> torch.Tensor:
    # HF:490-511 pre-LN block: x + attn(LN1(x)); x + fc2(quick_gelu(fc1(LN2(x))))
    p = f"backbone.encoder.layers.{i}."
    h = layer_norm(sd, p + "layer_norm1", x, cfg.ln_eps)
    x = x + attention(sd, p + "self_attn.", cfg, h)
    h = layer_norm(sd, p + "layer_norm2", x, cfg.ln_eps)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])        # HF:474
    h = quick_gelu(h)                                                         # HF:475
    h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])        # HF:476
    return x + h


def vision_tower(sd, cfg, pixel_values: torch.Tensor) -> torch.Tensor:
    # HF:757-782: embeddings -> pre_layernorm -> encoder; returns last_hidden_state [B,S,D]
    x = embeddings(sd, cfg, pixel_values)
    x = layer_norm(sd, "backbone.pre_layernorm", x, cfg.ln_eps)
    for i in range(cfg.layers):
        x = encoder_layer(sd, i, cfg, x)
    return x


def image_embedder(sd, cfg, pixel_values: torch.Tensor) -> torch.Tensor:
    # reference src/models.py:77-96 — post_layernorm on ALL tokens (Q6), patches * CLS, second LN
    last = vision_tower(sd, cfg, pixel_values)
    e = layer_norm(sd, "backbone.post_layernorm", last, cfg.ln_eps)          # :80
    cls = e[:, :1, :]                                                         # :82-83 (broadcast)
    e = e[:, 1:, :] * cls                                                     # :85
    e = layer_norm(sd, "post_post_layernorm", e, cfg.ln_eps)                  # :86
    return e                                                                  # [B,P,D] (grid reshape is a view, :88-94)


def box_bias(cfg) -> torch.Tensor:
    # HF:1097-1130 compute_box_bias / normalize_grid_corner_coordinates; x varies fastest
    g = cfg.grid
    xs = torch.arange(1, g + 1, dtype=torch.float32)
    xx, yy = torch.meshgrid(xs, xs, indexing="xy")
    coords = torch.stack((xx, yy), dim=-1)
    coords[..., 0] /= g
    coords[..., 1] /= g
    coords = coords.view(-1, 2).clip(0.0, 1.0)
    coord_bias = torch.log(coords + 1e-4) - torch.log1p(-coords + 1e-4)
    size = torch.full_like(coord_bias, 1.0)
    size[..., 0] /= g
    size[..., 1] /= g
    size_bias = torch.log(size + 1e-4) - torch.log1p(-size + 1e-4)
    return torch.cat([coord_bias, size_bias], dim=-1)                         # [P,4]


def box_predictor(sd, cfg, feats: torch.Tensor) -> torch.Tensor:
    # reference src/models.py:65-73; HF:1019-1025 box head (exact erf GELU twice)
    h = F.gelu(F.linear(feats, sd["box_head.dense0.weight"], sd["box_head.dense0.bias"]))
    h = F.gelu(F.linear(h, sd["box_head.dense1.weight"], sd["box_head.dense1.bias"]))
    z = F.linear(h, sd["box_head.dense2.weight"], sd["box_head.dense2.bias"])
    z = z + box_bias(cfg)                                                     # :71
    s = torch.sigmoid(z)                                                      # :72
    cx, cy, w, hh = s.unbind(-1)                                              # image_transforms.py:529-536
    return torch.stack([cx - 0.5 * w, cy - 0.5 * hh, cx + 0.5 * w, cy + 0.5 * hh], dim=-1)


def class_predictor(sd, cfg, feats: torch.Tensor) -> torch.Tensor:
    # reference src/models.py:24-38 (Q1 asymmetric normalisation, Q2 no shift/scale, Q3 maxpool3)
    e = F.linear(feats, sd["class_predictor.dense0.weight"], sd["class_predictor.dense0.bias"])
    e = e / (torch.linalg.norm(e, dim=-1, keepdim=True) + 1e-6)               # :28-30
    q = sd["queries"]
    q = q / torch.linalg.norm(q, dim=-1, keepdim=True) + 1e-6                 # :31-33 (precedence quirk)
    sims = e @ q.transpose(1, 2)                                              # :35
    return F.max_pool1d(sims, kernel_size=cfg.variants, stride=cfg.variants)  # :36


def forward(sd, cfg, image: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference src/models.py:98-119 -> (pred_boxes xyxy [B,P,4], pred_sims [B,P,C])."""
    feats = image_embedder(sd, cfg, image)
    return box_predictor(sd, cfg, feats), class_predictor(sd, cfg, feats)


def flops_per_image(cfg) -> Dict[str, float]:
    """Algorithmic FLOPs (2*M*N*K per GEMM), SURVEY.md §8d."""
    S, P, D, Fd, E, L = cfg.tokens, cfg.patches, cfg.hidden, cfg.ff, cfg.embed, cfg.layers
    patch = 2.0 * P * D * 3 * cfg.patch_size ** 2
    qkv = 2.0 * S * D * 3 * D
    core = 2.0 * 2.0 * S * S * D
    out = 2.0 * S * D * D
    mlp = 2.0 * 2.0 * S * D * Fd
    layer = qkv + core + out + mlp
    cls = 2.0 * P * D * E + 2.0 * P * E * cfg.n_queries
    box = 2.0 * 2.0 * P * D * D + 2.0 * P * D * 4
    fwd = patch + L * layer + cls + box
    bwd_ref_policy = 2.0 * (layer + cls + box)
    return {"patch": patch, "layer": layer, "attn_core": core, "cls": cls, "box": box, "fwd": fwd,
            "fwd_bwd_ref_policy": fwd + bwd_ref_policy, "fwd_bwd_full": 3.0 * fwd - patch}
