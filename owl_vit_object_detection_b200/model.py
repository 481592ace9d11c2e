"""`OwlViT` — the reference's model wrapper (reference src/models.py:41-119) on the sm_100a kernels.

Same constructor arguments, same `forward(image) -> (pred_boxes_xyxy, None, pred_sims, None)`,
same parameter / state-dict key names (`queries`, `backbone.*`, `post_post_layernorm.*`,
`class_predictor.dense0.*`, `box_head.dense{0,1,2}.*`), so HF weights load unchanged and
`torch.optim.AdamW(model.parameters())` works as in reference main.py:56-60.

Every parameter is a view into ONE flat fp32 buffer and every gradient a view into ONE flat fp32 gradient
buffer (see params.py), which is what the fused AdamW step and the NCCL gradient all-reduce operate on.

Supported training policy: the reference freeze rule (reference src/models.py:173-184) — last encoder layer,
both post layer norms, heads and query bank trainable.  The constructor applies it (the reference applies the
same rule right after construction in `load_model`).
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Optional, Tuple

import torch
from torch import nn

from .engine import Engine
from .params import ParamLayout
from .synth import OwlConfig, trainable_names


def config_from_hf(hf_cfg, n_queries: int, variants: int = 3) -> OwlConfig:
    """OwlConfig from a HuggingFace `OwlViTConfig` (reference src/models.py:152 loads owlvit-base-patch32)."""
    v = hf_cfg.vision_config
    return OwlConfig(image_size=v.image_size, patch_size=v.patch_size, hidden=v.hidden_size,
                     layers=v.num_hidden_layers, heads=v.num_attention_heads, ff=v.intermediate_size,
                     embed=hf_cfg.text_config.hidden_size, n_classes=n_queries // variants, variants=variants,
                     ln_eps=v.layer_norm_eps)


def state_dict_from_hf(hf_model) -> Dict[str, torch.Tensor]:
    """The tensors the reference wrapper keeps from `OwlViTForObjectDetection` (reference src/models.py:52-59),
    under the wrapper's own key names."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in hf_model.owlvit.vision_model.state_dict().items():
        if "position_ids" in k:
            continue
        out["backbone." + k] = v
    for k, v in hf_model.layer_norm.state_dict().items():
        out["post_post_layernorm." + k] = v
    out["class_predictor.dense0.weight"] = hf_model.class_head.dense0.weight
    out["class_predictor.dense0.bias"] = hf_model.class_head.dense0.bias
    for k, v in hf_model.box_head.state_dict().items():
        out["box_head." + k] = v
    return {k: v.detach() for k, v in out.items()}


class _Node(nn.Module):
    """Anonymous container so that dotted reference names become real sub-module paths."""


class _ForwardFn(torch.autograd.Function):
    """Connects the kernel-sequenced forward / backward to autograd (reference main.py:82-90).

    The activations the backward needs live in the engine's per-batch-size workspace, i.e. they belong to the LAST
    forward of that batch size.  Each forward stamps a generation number; a backward whose forward is no longer
    the latest (two forwards before one backward, an eval forward in between) raises instead of silently using
    the wrong activations."""

    @staticmethod
    def forward(ctx, anchor, image, model):
        boxes, sims = model.engine.forward(image, save_for_backward=True)
        ctx.model = model
        ctx.batch = image.shape[0]
        ctx.generation = model.engine.workspace(ctx.batch).generation
        return boxes, sims

    @staticmethod
    def backward(ctx, dboxes, dsims):
        model = ctx.model
        B = ctx.batch
        cfg = model.cfg
        dev = model._flat.device
        ws = model.engine.workspace(B)
        if ws.generation != ctx.generation:
            raise RuntimeError(
                "OwlViT.backward: another forward with the same batch size ran after the forward this backward "
                "belongs to; the saved activations (one workspace per batch size) were overwritten.  Call "
                "backward() before the next forward (reference main.py:82-90 order).")
        if dboxes is None:
            dboxes = torch.zeros((B, cfg.patches, 4), device=dev)
        if dsims is None:
            dsims = torch.zeros((B, cfg.patches, cfg.n_classes), device=dev)
        model._prepare_grads()
        model.engine.backward(dsims.contiguous().float(), dboxes.contiguous().float(), model._flat_grad)
        return None, None, None


class OwlViT(nn.Module):
    """
    Drop-in for the reference `OwlViT(pretrained_model, query_bank)`.

    pretrained_model: a HuggingFace `OwlViTForObjectDetection` (as in the reference), or a mapping
                      {reference state-dict key: tensor} together with `cfg`.
    query_bank:       [1, 3*C, E] text embeddings (3 prompt variants per class, class-major).
    """

    def __init__(self, pretrained_model, query_bank: torch.Tensor, cfg: Optional[OwlConfig] = None,
                 device: Optional[torch.device] = None):
        super().__init__()
        if isinstance(pretrained_model, Mapping):
            assert cfg is not None, "a state-dict needs an explicit OwlConfig"
            sd = dict(pretrained_model)
        else:
            cfg = cfg or config_from_hf(pretrained_model.config, query_bank.shape[1])
            sd = state_dict_from_hf(pretrained_model)
        sd["queries"] = query_bank.detach()
        assert query_bank.dim() == 3 and query_bank.shape[1] == cfg.n_queries and query_bank.shape[2] == cfg.embed, \
            f"query bank {tuple(query_bank.shape)} does not match [1, {cfg.n_queries}, {cfg.embed}]"
        self.cfg = cfg
        self.layout = ParamLayout(cfg)
        missing = [k for k in self.layout.shapes if k not in sd]
        assert not missing, f"missing tensors: {missing[:4]} ..."
        device = torch.device(device) if device is not None else query_bank.device
        self._flat = self.layout.pack(sd, device)
        self._flat_grad: Optional[torch.Tensor] = None
        self._symm_grad = None          # collective.SymmetricGrad once use_symmetric_grads() succeeded
        self._engine: Optional[Engine] = None
        self._anchor = None
        self._pre = None            # DevicePreprocessor, created on the first uint8 input
        self._names = list(self.layout.shapes)
        train = set(trainable_names(cfg))
        for name in self._names:
            p = nn.Parameter(self.layout.view(self._flat, name), requires_grad=name in train)
            self._register(name, p)

    # ------------------------------------------------------------------ parameter plumbing
    def _register(self, dotted: str, p: nn.Parameter) -> None:
        mod: nn.Module = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            if part not in mod._modules:
                mod.add_module(part, _Node())
            mod = mod._modules[part]
        mod.register_parameter(parts[-1], p)

    def _param(self, dotted: str) -> nn.Parameter:
        mod: nn.Module = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            mod = mod._modules[part]
        return mod._parameters[parts[-1]]

    def _apply(self, fn, recurse=True):
        """.to() / .cuda() / .float(): move the flat buffer once and re-point every parameter view at it."""
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise TypeError("OwlViT keeps fp32 master parameters (the kernels make their own fp16 operands)")
        moved = new_flat is not self._flat
        self._flat = new_flat
        if moved:
            self._engine = None
            self._flat_grad = None
            self._symm_grad = None
            self._anchor = None
            for name in self._names:
                p = self._param(name)
                p.data = self.layout.view(self._flat, name)
                p.grad = None
        return self

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            if not self._flat.is_cuda:
                raise RuntimeError("OwlViT runs on hand-written sm_100a kernels only: move the model to a CUDA "
                                   "device first (there is no CPU fallback)")
            self._engine = Engine(self.cfg, self.layout, self._flat)
            # the parameters alias the flat buffer but may carry version counters of their own (after .to())
            self._engine.watch([self._param(n) for n in self._names])
        return self._engine

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    @property
    def flat_trainable(self) -> torch.Tensor:
        return self._flat[self.layout.train_begin:]

    @property
    def flat_grad(self) -> torch.Tensor:
        if self._flat_grad is None:
            self._flat_grad = torch.zeros(self.layout.n_trainable_padded, dtype=torch.float32,
                                          device=self._flat.device)
        return self._flat_grad

    @staticmethod
    def _zero(t: torch.Tensor) -> None:
        """Clear a (slice of the) gradient buffer: a memset on the stream for device memory (owl_zero), torch otherwise."""
        if t.is_cuda and t.is_contiguous():
            from . import ops
            ops.zero(t)
        else:
            t.zero_()

    def _check_policy(self) -> None:
        train = set(trainable_names(self.cfg))
        for name in self._names:
            if self._param(name).requires_grad and name not in train:
                raise NotImplementedError(
                    f"{name} requires grad, but only the reference freeze policy (reference src/models.py:173-184) "
                    "has a backward pass in this build")

    def _prepare_grads(self) -> None:
        """Make every trainable parameter's .grad a view of the flat gradient buffer.  A parameter whose .grad
        is None (optimizer.zero_grad(set_to_none=True), reference main.py:74) gets its slice zeroed first - per
        parameter, so an optimizer that owns only a subset of the trainable tensors (and therefore only clears
        that subset) still starts each of its steps from zero.  The backward kernels ACCUMULATE into the buffer."""
        g = self.flat_grad
        L = self.layout
        names = [n for n in trainable_names(self.cfg) if self._param(n).requires_grad]
        wholesale = all(self._param(n).grad is None for n in names)
        if wholesale:
            self._zero(g)
        for n in names:
            p = self._param(n)
            o = L.offsets[n] - L.train_begin
            view = g[o:o + L._numel(n)].view(L.shapes[n])
            if p.grad is None:
                if not wholesale:
                    self._zero(view)
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                raise RuntimeError(f"{n}.grad was replaced by a foreign tensor; use model.zero_grad() or "
                                   "optimizer.zero_grad() instead")

    def zero_grad(self, set_to_none: bool = True) -> None:
        if self._flat_grad is not None:
            self._zero(self._flat_grad)
        if set_to_none:
            for n in trainable_names(self.cfg):
                self._param(n).grad = None

    def use_symmetric_grads(self, group=None) -> bool:
        """Move the flat gradient buffer into symmetric memory so that `allreduce_grads` can reduce it inside the
        NVSwitch (collective.SymmetricGrad).  Returns False (and changes nothing) without multicast support.  Call
        once, on every rank, before the first backward."""
        from .collective import SymmetricGrad
        sg = SymmetricGrad.create(self.layout.n_trainable_padded, self._flat.device, group)
        if sg is None:
            return False
        self._symm_grad = sg
        self._flat_grad = sg.buf
        for n in trainable_names(self.cfg):
            self._param(n).grad = None          # re-pointed at the new buffer by the next backward
        return True

    def allreduce_grads(self, group=None) -> None:
        """Data-parallel gradient sync over the flat fp32 gradient buffer: the in-switch multimem
        all-reduce when the buffer is symmetric (`use_symmetric_grads`), else ONE NCCL all-reduce.  The 1/world
        averaging is folded into the optimizer step (FusedAdamW(grad_mul=1/world))."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        if getattr(self, "_symm_grad", None) is not None and self._flat_grad is self._symm_grad.buf:
            self._symm_grad.all_reduce()
        else:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)

    # ------------------------------------------------------------------ reference src/models.py:98-119
    def forward(self, image: torch.Tensor) -> Tuple[torch.Tensor, None, torch.Tensor, None]:
        if image.dim() != 4:
            raise ValueError(f"expected [B,3,H,W], got {tuple(image.shape)}")
        if image.device != self._flat.device:
            raise RuntimeError(f"image on {image.device}, model on {self._flat.device}")
        if image.dtype == torch.uint8:
            # raw RGB pixels [B,H,W,3]: the reference's CPU preprocessing (src/dataset.py:64-71) runs on the device
            if image.shape[3] != 3:
                raise ValueError(f"uint8 input must be raw RGB [B,H,W,3], got {tuple(image.shape)}")
            IS = self.cfg.image_size
            if image.shape[1] == IS and image.shape[2] == IS and self.cfg.patch_size % 8 == 0:
                pass    # already at the model's resolution: normalised inside the patch gather (owl_u8_patches_f16)
            else:
                if self._pre is None:
                    from .preprocess import DevicePreprocessor
                    self._pre = DevicePreprocessor(IS, self._flat.device)
                image = self._pre(list(image))
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            self._check_policy()
            if self._anchor is None:
                self._anchor = torch.zeros(1, device=self._flat.device, requires_grad=True)
            boxes, sims = _ForwardFn.apply(self._anchor, image if image.dtype == torch.uint8 else image.float(), self)
        else:
            boxes, sims = self.engine.forward(image if image.dtype == torch.uint8 else image.float(),
                                              save_for_backward=False)
        return boxes, None, sims, None


class FusedAdamW:
    """torch.optim.AdamW semantics (reference main.py:56-60,91) as ONE kernel over the flat trainable range;
    the same kernel refreshes the fp16 GEMM operands.  `grad_mul` folds the 1/world of the DDP average."""

    def __init__(self, model: OwlViT, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, grad_mul: float = 1.0):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.grad_mul = lr, betas, eps, weight_decay, grad_mul
        n = model.layout.n_trainable_padded
        dev = model.flat_params.device
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)   # step count + bias corrections, on device

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.model.zero_grad(set_to_none)

    def step(self) -> None:
        from . import ops
        m = self.model
        eng = m.engine
        lo = m.layout.train_begin
        ops.adamw(m.flat_params[lo:], m.flat_grad, self.exp_avg, self.exp_avg_sq, eng.flat16[lo:], lr=self.lr,
                  beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay,
                  state=self.state, grad_mul=self.grad_mul)
        # the kernel wrote params through a raw pointer: tell the engine its fp16 shadow is already current
        eng._shadow_version = eng.version()
