"""Query-bank initialisation on the device (reference src/models.py:155-169).

The reference seeds its learned query bank with `OwlViTForObjectDetection(**inputs).text_embeds`: the HuggingFace CLIP
text tower (12 pre-LN layers, hidden 512, 8 heads, 16 tokens, causal mask) over three prompts per class, the
end-of-text token pooled, projected to 512 and L2-normalised (HF:644-690, 978, 984).  `TextTower` runs exactly that on
the library's kernels: the encoder layers are the same owl_layernorm / owl_gemm sequence as a vision layer
(HF:490-511), plus the four text-only kernels of csrc/text.cu.  fp16 GEMM operands, fp32 accumulation, fp32 residual
stream, like the vision path.  There is no CPU path: it raises without a CUDA device.
"""
from __future__ import annotations

from typing import Mapping, Optional

import torch

from . import ops


class TextTower:
    """text_embeds = TextTower(hf_model_or_state_dict, device)(input_ids, attention_mask)

    `source`: a HuggingFace `OwlViTForObjectDetection` / `OwlViTModel`, or a mapping with the keys of
    `OwlViTModel.state_dict()` (`text_model.*`, `text_projection.weight`)."""

    def __init__(self, source, device, *, heads: Optional[int] = None, eps: Optional[float] = None):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("TextTower runs on a CUDA device only (there is no CPU fallback)")
        if isinstance(source, Mapping):
            sd = dict(source)
            assert heads is not None, "a state-dict needs `heads`"
            eps = 1e-5 if eps is None else eps
        else:
            clip = source.owlvit if hasattr(source, "owlvit") else source
            sd = {k: v.detach() for k, v in clip.state_dict().items()
                  if k.startswith("text_model.") or k.startswith("text_projection.")}
            tc = clip.config.text_config
            heads = heads or tc.num_attention_heads
            eps = tc.layer_norm_eps if eps is None else eps
            assert tc.hidden_act == "quick_gelu", tc.hidden_act
        self.device, self.heads, self.eps = device, heads, float(eps)

        def f32(k):
            return sd[k].to(device=device, dtype=torch.float32).contiguous()

        def f16(t):
            return t.to(device=device, dtype=torch.float16).contiguous()

        e = "text_model.embeddings."
        self.tok, self.pos = f32(e + "token_embedding.weight"), f32(e + "position_embedding.weight")
        self.D = int(self.tok.shape[1])
        assert self.D % (64 * heads) == 0 and self.D // heads == 64, "head_dim 64 only"
        self.layers = []
        i = 0
        while f"text_model.encoder.layers.{i}.layer_norm1.weight" in sd:
            p = f"text_model.encoder.layers.{i}."
            a = p + "self_attn."
            self.layers.append(dict(
                ln1_w=f32(p + "layer_norm1.weight"), ln1_b=f32(p + "layer_norm1.bias"),
                ln2_w=f32(p + "layer_norm2.weight"), ln2_b=f32(p + "layer_norm2.bias"),
                # q, k, v stacked into one [3D, D] operand (one GEMM, packed q|k|v output like the vision path)
                wqkv=f16(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0)),
                bqkv=torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0)
                .to(device=device, dtype=torch.float32).contiguous(),
                wo=f16(sd[a + "out_proj.weight"]), bo=f32(a + "out_proj.bias"),
                w1=f16(sd[p + "mlp.fc1.weight"]), b1=f32(p + "mlp.fc1.bias"),
                w2=f16(sd[p + "mlp.fc2.weight"]), b2=f32(p + "mlp.fc2.bias")))
            i += 1
        assert self.layers, "no text encoder layers found"
        self.ff = int(self.layers[0]["w1"].shape[0])
        self.lnf_w, self.lnf_b = f32("text_model.final_layer_norm.weight"), f32("text_model.final_layer_norm.bias")
        self.proj = f16(sd["text_projection.weight"])          # [E, D], no bias (HF:852)
        self.E = int(self.proj.shape[0])

    @torch.no_grad()
    def __call__(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """input_ids [N, S] (any int dtype, any device), attention_mask [N, S] or None -> text_embeds [N, E] fp32,
        unit-norm rows (what `OwlViTForObjectDetection(**inputs).text_embeds` holds, reshaped to [1, N, E] by the
        caller)."""
        dev, D, H = self.device, self.D, self.heads
        with torch.cuda.device(dev):
            ids = input_ids.reshape(-1, input_ids.shape[-1]).to(device=dev, dtype=torch.int64).contiguous()
            N, S = int(ids.shape[0]), int(ids.shape[1])
            if S > int(self.pos.shape[0]):
                raise ValueError(f"{S} tokens per prompt, the model has {int(self.pos.shape[0])} positions")
            mask = None
            if attention_mask is not None:
                mask = attention_mask.reshape(N, S).to(device=dev, dtype=torch.int32).contiguous()
            M = N * S
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            x = torch.empty((M, D), dtype=torch.float32, device=dev)
            y16 = torch.empty((M, D), dtype=torch.float16, device=dev)
            qkv16 = torch.empty((M, 3 * D), dtype=torch.float16, device=dev)
            ctx16 = torch.empty((M, D), dtype=torch.float16, device=dev)
            h16 = torch.empty((M, self.ff), dtype=torch.float16, device=dev)
            ops.text_embed(ids, self.tok, self.pos, x, S=S, status=status)
            for L in self.layers:                                               # HF:490-511
                ops.layernorm(x, L["ln1_w"], L["ln1_b"], y16, rows=M, D=D, eps=self.eps)
                ops.gemm(y16, L["wqkv"], qkv16, M=M, N=3 * D, K=D, bias=L["bqkv"])
                ops.text_attn(qkv16, mask, ctx16, N=N, S=S, H=H, head_dim=64, scale=0.125)
                ops.gemm(ctx16, L["wo"], x, M=M, N=D, K=D, bias=L["bo"], resid=x)
                ops.layernorm(x, L["ln2_w"], L["ln2_b"], y16, rows=M, D=D, eps=self.eps)
                ops.gemm(y16, L["w1"], h16, M=M, N=self.ff, K=D, bias=L["b1"], act="quick_gelu")
                ops.gemm(h16, L["w2"], x, M=M, N=D, K=self.ff, bias=L["b2"], resid=x)
            pooled16 = torch.empty((N, D), dtype=torch.float16, device=dev)
            ops.text_pool_ln(x, ids, self.lnf_w, self.lnf_b, pooled16, N=N, S=S, D=D, eps=self.eps)
            emb = torch.empty((N, self.E), dtype=torch.float32, device=dev)
            ops.gemm(pooled16, self.proj, emb, M=N, N=self.E, K=D)
            ops.l2norm_rows(emb, emb)
            if int(status.item()) & 8:
                raise IndexError("TextTower: a token id lies outside the vocabulary (nn.Embedding raises here too)")
        return emb


def text_query_bank(hf_model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], device) -> torch.Tensor:
    """reference src/models.py:165-169: `queries = _model(**inputs).text_embeds` -> [1, N, E] on `device`."""
    return TextTower(hf_model, device)(input_ids, attention_mask).unsqueeze(0)
