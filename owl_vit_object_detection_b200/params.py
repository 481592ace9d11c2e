"""Flat parameter storage for the OWL-ViT hot path.

All parameters of the reference wrapper (`OwlViT`, reference src/models.py:48-61) live in ONE
contiguous fp32 buffer; the tensors the reference freeze rule leaves trainable (reference
src/models.py:173-184: last encoder layer, heads, both post layer norms, query bank) sit together at
its end, so that
  * their gradients form one flat fp32 buffer (a single NCCL all-reduce per step),
  * the fused AdamW kernel walks one range,
  * q/k/v projection weights of a layer are adjacent in q,k,v order and are used as one [3D, D] GEMM operand.
An fp16 shadow of the same layout feeds the tensor-core GEMMs.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .synth import OwlConfig, param_shapes, trainable_names

ALIGN = 64  # elements; keeps every tensor 128-byte aligned in the fp16 shadow (TMA needs 16 bytes)


def _order(cfg: OwlConfig) -> List[str]:
    """Storage order: frozen tensors first, trainable last; inside a layer q,k,v are adjacent."""
    names = list(param_shapes(cfg))
    train = set(trainable_names(cfg))

    def layer_sorted(ns: List[str]) -> List[str]:
        out: List[str] = []
        seen = set()
        for n in ns:
            if n in seen:
                continue
            if ".self_attn.k_proj.weight" in n:
                # reference/HF state-dict order is k, v, q, out: emit q, k, v weights then q, k, v biases
                pre = n[: n.index("k_proj.weight")]
                group = [pre + f"{p}_proj.weight" for p in "qkv"] + [pre + f"{p}_proj.bias" for p in "qkv"]
                for gname in group:
                    out.append(gname)
                    seen.add(gname)
                continue
            out.append(n)
            seen.add(n)
        return out

    ordered = layer_sorted(names)
    return [n for n in ordered if n not in train] + [n for n in ordered if n in train]


class ParamLayout:
    def __init__(self, cfg: OwlConfig):
        self.cfg = cfg
        self.shapes = param_shapes(cfg)
        self.trainable = trainable_names(cfg)
        tset = set(self.trainable)
        self.offsets: Dict[str, int] = {}
        off = 0
        self.train_begin = None
        prev = None
        for n in _order(cfg):
            if n in tset and self.train_begin is None:
                off = (off + ALIGN - 1) // ALIGN * ALIGN
                self.train_begin = off
            numel = 1
            for s in self.shapes[n]:
                numel *= s
            # q,k,v weights (and biases) must be exactly adjacent: every such tensor is a multiple of ALIGN
            # for the supported widths, so plain alignment keeps them contiguous.
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            if prev is not None and "_proj." in n and "q_proj" not in n and "out_proj" not in n:
                assert off == self.offsets[prev] + self._numel(prev), (prev, n)
            self.offsets[n] = off
            off += numel
            prev = n
        self.total = (off + ALIGN - 1) // ALIGN * ALIGN
        if self.train_begin is None:
            self.train_begin = self.total
        self.n_trainable_padded = self.total - self.train_begin

    def _numel(self, n: str) -> int:
        k = 1
        for s in self.shapes[n]:
            k *= s
        return k

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o = self.offsets[name]
        return flat[o:o + self._numel(name)].view(self.shapes[name])

    def span(self, first: str, last: str) -> Tuple[int, int]:
        """[begin, end) element range covering tensors first..last (which must be adjacent)."""
        return self.offsets[first], self.offsets[last] + self._numel(last)

    def grad_buckets(self) -> List[Tuple[int, int]]:
        """Three contiguous element ranges of the flat gradient buffer (offsets relative to the trainable tail) that
        tile it, in the order the backward pass COMPLETES them, so a data-parallel driver can reduce one range while
        the kernels of the next still run: heads + both post layer norms; the MLP half of the last layer (fc1, fc2,
        layer_norm2); the attention half + layer_norm1 + the query bank."""
        p = f"backbone.encoder.layers.{self.cfg.layers - 1}."
        t0 = self.train_begin
        a = self.offsets["backbone.post_layernorm.weight"] - t0
        b = self.offsets[p + "mlp.fc1.weight"] - t0
        assert 0 < b < a <= self.n_trainable_padded and self.offsets["queries"] == t0, "unexpected trainable layout"
        return [(a, self.n_trainable_padded), (b, a), (0, b)]

    def pack(self, sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
        flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        for n in self.shapes:
            self.view(flat, n).copy_(sd[n].to(device=device, dtype=torch.float32))
        return flat
