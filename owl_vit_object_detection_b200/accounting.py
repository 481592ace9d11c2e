"""Algorithmic work of the hot path: the figures every roofline fraction in bench.py is computed from.
FLOPs are 2*M*N*K per contraction; bytes are what an ideal single pass has to move."""
from __future__ import annotations

from typing import Dict

from .synth import OwlConfig


def flops_per_image(cfg: OwlConfig) -> Dict[str, float]:
    """reference src/models.py:98-119 (+ HF vision tower) per image; `fwd_bwd_ref_policy` adds the backward of what the
    reference freeze rule (src/models.py:173-184) leaves trainable: last encoder layer + heads = 2x their forward."""
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


def matcher_cost_bytes_per_image(P: int, C: int, T: int) -> int:
    """reference src/matcher.py:103-131, one pass: read sims [P,C] f32 + boxes [P,4] f32 + T x (4 f32 + i64 label),
    write cost [P,T] f32  (216,816 / 309,936 / 426,336 bytes for T = 10 / 50 / 100 at P = 576, C = 80)."""
    return P * C * 4 + P * 4 * 4 + T * 24 + P * T * 4
