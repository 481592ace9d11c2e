"""Forward / backward of the OWL-ViT hot path as a fixed sequence of libowl_b200.so kernel launches.

Mirrors, step for step (HF = transformers/models/owlvit/modeling_owlvit.py as cited in SURVEY.md):
  reference src/models.py:98-119  OwlViT.forward          -> Engine.forward
  reference src/models.py:77-96   image_embedder + HF:757-782 vision tower
  reference src/models.py:65-73   box_predictor  + HF:1009-1025
  reference src/models.py:24-38   PatchedOwlViTClassPredictionHead.forward
  reference main.py:90            loss.backward() under the reference freeze rule -> Engine.backward

This module only sequences kernels and owns buffers (torch is used for device memory and streams);
every arithmetic op runs in a hand-written sm_100a kernel.  There is no CPU / torch fallback.

Data layout in HBM (B images, S = P + 1 tokens, D hidden):
  residual stream   x      fp32 [B*S, D]   (row = image-major token index, CLS first)
  GEMM operands     *16    fp16, row-major, produced by the LayerNorm / previous GEMM epilogue
  QKV               qkv16  fp16 [B*S, 3D]  (q | k | v column blocks, head h at columns h*dh inside a block)
  attention scores / probabilities never leave the SM (fused forward and backward kernels)
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .params import ParamLayout
from .synth import OwlConfig


def _box_bias(cfg: OwlConfig, device) -> torch.Tensor:
    """HF:1097-1130 compute_box_bias: a constant of the grid size (evaluated once, on the host, in fp32)."""
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
    return torch.cat([coord_bias, size_bias], dim=-1).contiguous().to(device)


class Workspace:
    """All activation buffers for one batch size (allocated once, reused every step)."""

    def __init__(self, cfg: OwlConfig, B: int, device, train: bool):
        S, P, D, F, E, H = cfg.tokens, cfg.patches, cfg.hidden, cfg.ff, cfg.embed, cfg.heads
        Q, C = cfg.n_queries, cfg.n_classes
        f16, f32 = torch.float16, torch.float32
        self.B = B
        self.generation = 0                  # bumped by every Engine.forward that writes this workspace
        self.Kp = (3 * cfg.patch_size ** 2 + 7) // 8 * 8
        self.Sp = (S + 7) // 8 * 8

        def z(shape, dt):
            return torch.zeros(shape, dtype=dt, device=device)

        self.patches16 = z((B * P, self.Kp), f16)
        self.emb = z((B * S, D), f32)
        self.x = z((B * S, D), f32)          # residual stream (input of the last layer after the loop)
        self.x_mid = z((B * S, D), f32)      # last layer: after attention
        self.x_out = z((B * S, D), f32)      # last layer: after MLP
        self.h1 = z((B * S, D), f16)         # LN1 output (kept for the last layer's wgrad)
        self.h2 = z((B * S, D), f16)         # LN2 output
        self.qkv = z((B * S, 3 * D), f16)
        self.lse = z((B * H, S), f32)        # last layer: log-sum-exp of the scaled scores (for the backward)
        self.ctx = z((B * S, D), f16)
        self.m = z((B * S, F), f16)
        self.mpre = z((B * S, F), f16)
        self.ecls = z((B, D), f32)
        self.feats = z((B * P, D), f16)
        self.e32 = z((B * P, E), f32)
        self.en16 = z((B * P, E), f16)
        self.qn16 = z((Q, E), f16)
        self.argmax = z((B * P, C), torch.uint8)
        self.bh0 = z((B * P, D), f16)
        self.bh0pre = z((B * P, D), f16)
        self.bh1 = z((B * P, D), f16)
        self.bh1pre = z((B * P, D), f16)
        self.sig = z((B * P, 4), f32)
        self._bw = None
        self._cfg, self._device = cfg, device

    def backward_buffers(self):
        """Buffers only the backward pass needs (allocated on first use)."""
        if self._bw is None:
            cfg, B, device = self._cfg, self.B, self._device
            S, P, D, F, E, H = cfg.tokens, cfg.patches, cfg.hidden, cfg.ff, cfg.embed, cfg.heads
            Q = cfg.n_queries
            f16, f32 = torch.float16, torch.float32

            def z(shape, dt):
                return torch.zeros(shape, dtype=dt, device=device)

            class _B:
                pass
            b = _B()
            b.gscale = z((4,), f32)
            b.dfull = z((B * P, Q), f16)
            b.den32 = z((B * P, E), f32)
            b.de16 = z((B * P, E), f16)
            b.dqn32 = z((Q, E), f32)
            b.dz = z((B * P, 4), f32)
            b.dpre1 = z((B * P, D), f16)
            b.dpre0 = z((B * P, D), f16)
            b.dfeats = z((B * P, D), f32)
            b.dx_out = z((B * S, D), f32)
            b.dx_mid = z((B * S, D), f32)
            b.dcl = z((B, D), f32)
            b.g16 = z((B * S, D), f16)
            b.dmpre = z((B * S, F), f16)
            b.dh32 = z((B * S, D), f32)
            b.dctx = z((B * S, D), f16)
            b.delta = z((B * H, S), f32)
            b.dq32 = z((B * S, D), f32)                # fp32 accumulation of dQ across key blocks
            b.dqkv = z((B * S, 3 * D), f16)
            self._bw = b
        return self._bw


def alias_version(flat: torch.Tensor, aliases) -> int:
    """Sum of the autograd version counters of a flat buffer and of tensors that alias it (Engine.version)."""
    v = flat._version
    for t in aliases:
        v += t._version
    return v


class Engine:
    def __init__(self, cfg: OwlConfig, layout: ParamLayout, flat32: torch.Tensor):
        assert flat32.is_cuda and flat32.dtype == torch.float32 and flat32.numel() == layout.total
        assert cfg.hidden % 128 == 0 and cfg.embed % 128 == 0 and cfg.head_dim == 64, \
            "kernels are written for hidden/embed multiples of 128 and 64-wide heads"
        self.cfg, self.layout = cfg, layout
        self.flat32 = flat32
        self.device = flat32.device
        self.flat16 = torch.zeros(layout.total, dtype=torch.float16, device=self.device)
        self.Kp = (3 * cfg.patch_size ** 2 + 7) // 8 * 8
        self.patch_w16 = torch.zeros((cfg.hidden, self.Kp), dtype=torch.float16, device=self.device)
        self.box_bias = _box_bias(cfg, self.device)
        self._ws: Dict[int, Workspace] = {}
        self._watch: list = []          # tensors aliasing flat32 that carry version counters of their own (watch())
        self._shadow_version = None
        self._lut: Optional[torch.Tensor] = None
        self.refresh_shadow()

    def watch(self, tensors) -> None:
        """Tensors aliasing `flat32` whose own version counters must also invalidate the fp16 shadow.  The
        nn.Parameter views share `flat32`'s counter only until `Module.to()` re-points them through `.data =`
        (set_data leaves a tensor with a counter of its own): from then on an in-place update by torch.optim
        (reference main.py:56-60,91) bumps the parameter's counter, not the flat buffer's."""
        self._watch = list(tensors)
        self._shadow_version = self.version()

    def version(self) -> int:
        """Changes whenever flat32 or a watched alias was modified in place through torch."""
        return alias_version(self.flat32, self._watch)

    # ------------------------------------------------------------------ parameters
    def p32(self, name: str) -> torch.Tensor:
        return self.layout.view(self.flat32, name)

    def p16(self, name: str) -> torch.Tensor:
        return self.layout.view(self.flat16, name)

    def refresh_shadow(self, trainable_only: bool = False) -> None:
        """fp32 master -> fp16 GEMM operands (owl_cast_f16).  Cheap: 85 us for all of B/32, 8 us trainable."""
        lo = self.layout.train_begin if trainable_only else 0
        with torch.cuda.device(self.device):
            ops.cast_f16(self.flat32[lo:], self.flat16[lo:])
        if not trainable_only:
            K = 3 * self.cfg.patch_size ** 2
            w = self.p16("backbone.embeddings.patch_embedding.weight").view(self.cfg.hidden, K)
            if self.Kp == K:
                self.patch_w16 = w
            else:
                self.patch_w16[:, :K].copy_(w)
        self._shadow_version = self.version()

    def sync_shadow(self) -> None:
        if self.version() != self._shadow_version:
            self.refresh_shadow()

    def pixel_lut(self) -> torch.Tensor:
        """[3, 256] fp32: rescale (/255) + CLIP normalise of a raw byte, as the reference's processor does it
        (reference src/dataset.py:64-71; built by preprocess.rescale_normalize_lut)."""
        if self._lut is None:
            from .preprocess import rescale_normalize_lut
            self._lut = torch.from_numpy(rescale_normalize_lut()).to(self.device).contiguous()
        return self._lut

    def workspace(self, B: int) -> Workspace:
        ws = self._ws.get(B)
        if ws is None:
            ws = Workspace(self.cfg, B, self.device, True)
            self._ws[B] = ws
        return ws

    # ------------------------------------------------------------------ forward
    def _attention(self, ws: Workspace, B: int, keep_lse: bool) -> None:
        """Fused tcgen05 attention (HF:379-404): scores stay in TMEM.  keep_lse=True (last layer when a backward
        follows) also stores the per-row log-sum-exp, from which the backward pass recomputes the probabilities."""
        cfg = self.cfg
        ops.flash_attn_fwd(ws.qkv, ws.ctx, B=B, S=cfg.tokens, H=cfg.heads, head_dim=cfg.head_dim,
                           scale=cfg.head_dim ** -0.5, lse=ws.lse if keep_lse else None)

    def forward(self, image: torch.Tensor, save_for_backward: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """image [B,3,IS,IS] fp32 CUDA -> (pred_boxes [B,P,4] xyxy fp32, pred_sims [B,P,C] fp32)."""
        with torch.cuda.device(self.device):     # kernels launch on the current device's current stream
            return self._forward(image, save_for_backward)

    def _forward(self, image: torch.Tensor, save_for_backward: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        cfg, L = self.cfg, self.layout
        assert image.is_cuda and image.dim() == 4
        raw = image.dtype == torch.uint8      # raw RGB bytes [B,IS,IS,3] at the model's resolution (see ops.u8_patches_f16)
        if raw:
            assert image.shape[1] == cfg.image_size and image.shape[2] == cfg.image_size and image.shape[3] == 3, \
                f"expected uint8 [B,{cfg.image_size},{cfg.image_size},3], got {tuple(image.shape)}"
            assert cfg.patch_size % 8 == 0, "the raw-pixel path needs a patch size that is a multiple of 8"
        else:
            assert image.dtype == torch.float32
            assert image.shape[1] == 3 and image.shape[2] == cfg.image_size and image.shape[3] == cfg.image_size, \
                f"expected [B,3,{cfg.image_size},{cfg.image_size}], got {tuple(image.shape)}"
        image = image.contiguous()
        B = image.shape[0]
        S, P, D, F, E = cfg.tokens, cfg.patches, cfg.hidden, cfg.ff, cfg.embed
        Q, C, eps = cfg.n_queries, cfg.n_classes, cfg.ln_eps
        self.sync_shadow()
        ws = self.workspace(B)
        ws.generation += 1
        ops.l2_persist(ws.x)      # opt-in (OWL_L2_PERSIST=1): keep the fp32 residual stream in the L2 set-aside
        M, MP = B * S, B * P

        # ---- embeddings HF:334-344 + pre_layernorm HF:768
        if raw:
            ops.u8_patches_f16(image, self.pixel_lut(), ws.patches16, cfg.patch_size)
        else:
            ops.im2col_f16(image, ws.patches16, cfg.patch_size)
        ops.gemm(ws.patches16, self.patch_w16, ws.emb, M=MP, N=D, K=ws.Kp,
                 pos=self.p32("backbone.embeddings.position_embedding.weight"), rows_per_img=P)
        ops.layernorm(ws.emb, self.p32("backbone.pre_layernorm.weight"), self.p32("backbone.pre_layernorm.bias"),
                      ws.x, rows=M, D=D, eps=eps, cls_emb=self.p32("backbone.embeddings.class_embedding"),
                      pos0=self.p32("backbone.embeddings.position_embedding.weight"), tokens=S)

        # ---- encoder HF:490-511
        for i in range(cfg.layers):
            p = f"backbone.encoder.layers.{i}."
            last = i == cfg.layers - 1
            x_in = ws.x
            x_mid = ws.x_mid if last else ws.x
            x_out = ws.x_out if last else ws.x
            lo, hi = L.span(p + "self_attn.q_proj.weight", p + "self_attn.v_proj.weight")
            wqkv = self.flat16[lo:hi].view(3 * D, D)
            lo, hi = L.span(p + "self_attn.q_proj.bias", p + "self_attn.v_proj.bias")
            bqkv = self.flat32[lo:hi]
            ops.layernorm(x_in, self.p32(p + "layer_norm1.weight"), self.p32(p + "layer_norm1.bias"), ws.h1,
                          rows=M, D=D, eps=eps)
            ops.gemm(ws.h1, wqkv, ws.qkv, M=M, N=3 * D, K=D, bias=bqkv)
            self._attention(ws, B, keep_lse=last and save_for_backward)
            ops.gemm(ws.ctx, self.p16(p + "self_attn.out_proj.weight"), x_mid, M=M, N=D, K=D,
                     bias=self.p32(p + "self_attn.out_proj.bias"), resid=x_in)
            ops.layernorm(x_mid, self.p32(p + "layer_norm2.weight"), self.p32(p + "layer_norm2.bias"), ws.h2,
                          rows=M, D=D, eps=eps)
            ops.gemm(ws.h2, self.p16(p + "mlp.fc1.weight"), ws.m, M=M, N=F, K=D, bias=self.p32(p + "mlp.fc1.bias"),
                     act="quick_gelu", pre_out=ws.mpre if (last and save_for_backward) else None)
            ops.gemm(ws.m, self.p16(p + "mlp.fc2.weight"), x_out, M=M, N=D, K=F, bias=self.p32(p + "mlp.fc2.bias"),
                     resid=x_mid)

        # ---- reference src/models.py:80-86: post LN (all tokens) x CLS, second LN
        g1, b1 = self.p32("backbone.post_layernorm.weight"), self.p32("backbone.post_layernorm.bias")
        ops.layernorm(ws.x_out, g1, b1, ws.ecls, rows=B, D=D, eps=eps, x_stride=S * D)
        ops.post_fuse(ws.x_out, ws.ecls, g1, b1, self.p32("post_post_layernorm.weight"),
                      self.p32("post_post_layernorm.bias"), ws.feats, B=B, P=P, D=D, eps=eps)

        # ---- class head, reference src/models.py:24-38
        sims = torch.empty((B, P, C), dtype=torch.float32, device=self.device)
        ops.gemm(ws.feats, self.p16("class_predictor.dense0.weight"), ws.e32, M=MP, N=E, K=D,
                 bias=self.p32("class_predictor.dense0.bias"))
        ops.rownorm_f16(ws.e32, ws.en16, rows=MP, E=E, query_mode=False)
        ops.rownorm_f16(self.p32("queries").view(Q, E), ws.qn16, rows=Q, E=E, query_mode=True)
        ops.gemm(ws.en16, ws.qn16, sims.view(MP, C), M=MP, N=Q, K=E, pool3=True, argmax=ws.argmax)

        # ---- box head, reference src/models.py:65-73 + HF:1019-1025
        boxes = torch.empty((B, P, 4), dtype=torch.float32, device=self.device)
        ops.gemm(ws.feats, self.p16("box_head.dense0.weight"), ws.bh0, M=MP, N=D, K=D,
                 bias=self.p32("box_head.dense0.bias"), act="gelu", pre_out=ws.bh0pre if save_for_backward else None)
        ops.gemm(ws.bh0, self.p16("box_head.dense1.weight"), ws.bh1, M=MP, N=D, K=D,
                 bias=self.p32("box_head.dense1.bias"), act="gelu", pre_out=ws.bh1pre if save_for_backward else None)
        ops.box_tail(ws.bh1, self.p32("box_head.dense2.weight"), self.p32("box_head.dense2.bias"), self.box_bias,
                     boxes.view(MP, 4), ws.sig, M=MP, P=P, D=D)
        return boxes, sims

    # ------------------------------------------------------------------ backward (reference freeze policy)
    def _wgrad(self, dy16: torch.Tensor, x16: torch.Tensor, dw: torch.Tensor, *, rows: int, n_out: int, n_in: int,
               gscale: torch.Tensor) -> None:
        """dw [n_out, n_in] += (1/S) dy^T x   (split-K tcgen05 GEMM, both operands read as stored)."""
        tiles = ((n_out + 127) // 128) * ((n_in + 255) // 256)
        split = max(1, min(16, 148 // max(tiles, 1), (rows + 511) // 512))
        ops.gemm(dy16, x16, dw, M=n_out, N=n_in, K=rows, a_mn=True, b_mn=True, a_ld=dy16.stride(0),
                 b_ld=x16.stride(0), ldo=n_in, split_k=split, out_mode=2, alpha_dev=gscale[1:])

    def grad_buckets(self):
        """See ParamLayout.grad_buckets: the ranges `backward(..., on_ready=)` reports, in completion order."""
        return self.layout.grad_buckets()

    def backward(self, dsims: torch.Tensor, dboxes: torch.Tensor, grad_flat: torch.Tensor, on_ready=None) -> None:
        """Accumulates d(loss)/d(trainable parameters) into grad_flat (fp32, the trainable tail of the flat
        layout), given d(loss)/d(pred_sims) [B,P,C] and d(loss)/d(pred_boxes) [B,P,4] of the LAST forward.
        `on_ready(i)` is called right after the kernels that complete `grad_buckets()[i]` have been enqueued."""
        with torch.cuda.device(self.device):
            self._backward(dsims, dboxes, grad_flat, on_ready)

    def _backward(self, dsims: torch.Tensor, dboxes: torch.Tensor, grad_flat: torch.Tensor, on_ready=None) -> None:
        cfg, L = self.cfg, self.layout
        B = dsims.shape[0]
        ws = self.workspace(B)
        bw = ws.backward_buffers()
        S, P, D, F, E, H, dh = cfg.tokens, cfg.patches, cfg.hidden, cfg.ff, cfg.embed, cfg.heads, cfg.head_dim
        Q, C, eps, Sp = cfg.n_queries, cfg.n_classes, cfg.ln_eps, ws.Sp
        M, MP = B * S, B * P
        assert dsims.is_cuda and dsims.dtype == torch.float32 and dsims.shape == (B, P, C)
        assert dboxes.dtype == torch.float32 and dboxes.shape == (B, P, 4)
        assert grad_flat.dtype == torch.float32 and grad_flat.numel() == L.n_trainable_padded
        dsims, dboxes = dsims.contiguous(), dboxes.contiguous()
        gs = bw.gscale

        def gview(name: str) -> torch.Tensor:
            o = L.offsets[name] - L.train_begin
            return grad_flat[o:o + L._numel(name)].view(L.shapes[name])

        def gspan(first: str, last: str) -> torch.Tensor:
            lo, hi = L.span(first, last)
            return grad_flat[lo - L.train_begin:hi - L.train_begin]

        ops.grad_scale(dsims, dboxes, gs, target=64.0)

        # ---- box head (reference src/models.py:65-73, HF:1019-1025)
        ops.box_tail_bwd(dboxes, ws.sig, self.p32("box_head.dense2.weight"), ws.bh1pre, ws.bh1, gs, bw.dz, bw.dpre1,
                         gview("box_head.dense2.weight"), gview("box_head.dense2.bias"), M=MP, D=D)
        self._wgrad(bw.dpre1, ws.bh0, gview("box_head.dense1.weight"), rows=MP, n_out=D, n_in=D, gscale=gs)
        ops.colsum(bw.dpre1, gview("box_head.dense1.bias"), M=MP, N=D, gscale=gs)
        ops.gemm(bw.dpre1, self.p16("box_head.dense1.weight"), bw.dpre0, M=MP, N=D, K=D, b_mn=True,
                 act="gelu_grad", act_src=ws.bh0pre)
        self._wgrad(bw.dpre0, ws.feats, gview("box_head.dense0.weight"), rows=MP, n_out=D, n_in=D, gscale=gs)
        ops.colsum(bw.dpre0, gview("box_head.dense0.bias"), M=MP, N=D, gscale=gs)
        ops.gemm(bw.dpre0, self.p16("box_head.dense0.weight"), bw.dfeats, M=MP, N=D, K=D, b_mn=True)

        # ---- class head (reference src/models.py:24-38)
        ops.pool3_bwd(dsims, ws.argmax, gs, bw.dfull)
        ops.gemm(bw.dfull, ws.qn16, bw.den32, M=MP, N=E, K=Q, b_mn=True)                      # d(en) = dfull qn
        ops.zero(bw.dqn32)
        ops.gemm(bw.dfull, ws.en16, bw.dqn32, M=Q, N=E, K=MP, a_mn=True, b_mn=True, a_ld=Q, b_ld=E, ldo=E,
                 split_k=max(1, min(16, (MP + 511) // 512)), out_mode=2)                       # d(qn) = dfull^T en
        ops.rownorm_bwd(self.p32("queries").view(Q, E), bw.dqn32, gview("queries").view(Q, E), rows=Q, E=E,
                        query_mode=True, gscale=gs)
        ops.rownorm_bwd(ws.e32, bw.den32, bw.de16, rows=MP, E=E, query_mode=False, gscale=gs)
        self._wgrad(bw.de16, ws.feats, gview("class_predictor.dense0.weight"), rows=MP, n_out=E, n_in=D, gscale=gs)
        ops.colsum(bw.de16, gview("class_predictor.dense0.bias"), M=MP, N=E, gscale=gs)
        ops.gemm(bw.de16, self.p16("class_predictor.dense0.weight"), bw.dfeats, M=MP, N=D, K=E, b_mn=True,
                 out_mode=1)                                                                   # dfeats += de Wc

        # ---- reference src/models.py:80-86 backward
        g1n, b1n = "backbone.post_layernorm.weight", "backbone.post_layernorm.bias"
        ops.zero(bw.dcl)
        ops.post_fuse_bwd(ws.x_out, ws.ecls, self.p32(g1n), self.p32(b1n), self.p32("post_post_layernorm.weight"),
                          bw.dfeats, bw.dx_out, bw.dcl, gview(g1n), gview(b1n), gview("post_post_layernorm.weight"),
                          gview("post_post_layernorm.bias"), B=B, P=P, D=D, eps=eps, gscale=gs)
        ops.layernorm_bwd(ws.x_out, bw.dcl, self.p32(g1n), gview(g1n), gview(b1n), rows=B, D=D, eps=eps, gscale=gs,
                          dx=bw.dx_out, x_stride=S * D, dx_stride=S * D)                       # CLS rows

        if on_ready is not None:
            on_ready(0)

        # ---- last encoder layer (HF:490-511), MLP half
        p = f"backbone.encoder.layers.{cfg.layers - 1}."
        # fc2 bias gradient + the fp16 operand of the two fc2 backward GEMMs in ONE pass over dx_out
        ops.colsum(bw.dx_out, gview(p + "mlp.fc2.bias"), M=M, N=D, gscale=gs, cast_to=bw.g16)
        self._wgrad(bw.g16, ws.m, gview(p + "mlp.fc2.weight"), rows=M, n_out=D, n_in=F, gscale=gs)
        ops.gemm(bw.g16, self.p16(p + "mlp.fc2.weight"), bw.dmpre, M=M, N=F, K=D, b_mn=True,
                 act="quick_gelu_grad", act_src=ws.mpre)
        self._wgrad(bw.dmpre, ws.h2, gview(p + "mlp.fc1.weight"), rows=M, n_out=F, n_in=D, gscale=gs)
        ops.colsum(bw.dmpre, gview(p + "mlp.fc1.bias"), M=M, N=F, gscale=gs)
        ops.gemm(bw.dmpre, self.p16(p + "mlp.fc1.weight"), bw.dh32, M=M, N=D, K=F, b_mn=True)
        ops.layernorm_bwd(ws.x_mid, bw.dh32, self.p32(p + "layer_norm2.weight"), gview(p + "layer_norm2.weight"),
                          gview(p + "layer_norm2.bias"), rows=M, D=D, eps=eps, gscale=gs, dx=bw.dx_mid,
                          dx_add=bw.dx_out)

        if on_ready is not None:
            on_ready(1)

        # ---- attention half
        # out-proj bias gradient + the fp16 operand of the two out-proj backward GEMMs in ONE pass over dx_mid
        ops.colsum(bw.dx_mid, gview(p + "self_attn.out_proj.bias"), M=M, N=D, gscale=gs, cast_to=bw.g16)
        self._wgrad(bw.g16, ws.ctx, gview(p + "self_attn.out_proj.weight"), rows=M, n_out=D, n_in=D, gscale=gs)
        ops.gemm(bw.g16, self.p16(p + "self_attn.out_proj.weight"), bw.dctx, M=M, N=D, K=D, b_mn=True)
        qkv, dqkv = ws.qkv, bw.dqkv
        scale = dh ** -0.5
        # attention backward (autograd of HF:393-404) in one kernel: P is recomputed from the saved log-sum-exp,
        # dS = P (scale dctx v^T - delta) with delta = scale * rowsum(dctx . ctx); S / P / dP / dS stay on the SM
        ops.attn_delta(ws.ctx, bw.dctx, bw.delta, B=B, S=S, H=H, head_dim=dh, alpha=scale)
        ops.attn_bwd(qkv, bw.dctx, ws.lse, bw.delta, dqkv, bw.dq32, B=B, S=S, H=H, head_dim=dh, scale=scale)
        gw = gspan(p + "self_attn.q_proj.weight", p + "self_attn.v_proj.weight").view(3 * D, D)
        gb = gspan(p + "self_attn.q_proj.bias", p + "self_attn.v_proj.bias")
        self._wgrad(dqkv, ws.h1, gw, rows=M, n_out=3 * D, n_in=D, gscale=gs)
        ops.colsum(dqkv, gb, M=M, N=3 * D, gscale=gs)
        lo, hi = L.span(p + "self_attn.q_proj.weight", p + "self_attn.v_proj.weight")
        ops.gemm(dqkv, self.flat16[lo:hi].view(3 * D, D), bw.dh32, M=M, N=D, K=3 * D, b_mn=True)
        ops.layernorm_bwd(ws.x, bw.dh32, self.p32(p + "layer_norm1.weight"), gview(p + "layer_norm1.weight"),
                          gview(p + "layer_norm1.bias"), rows=M, D=D, eps=eps, gscale=gs)
        if on_ready is not None:
            on_ready(2)
