"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (fp32 torch, op for op) of what reference src/models.py:165-169 computes to seed the query bank:
`OwlViTForObjectDetection(**inputs).text_embeds`, i.e. the HuggingFace OWL-ViT TEXT tower.  "HF:" lines refer to the
installed `transformers/models/owlvit/modeling_owlvit.py` (5.5.0; the reference pins 4.30.2, same arithmetic).

Pinned by tests/test_oracle_text.py against the real HuggingFace classes (the third-party code the reference calls).
"""
import torch
import torch.nn.functional as F


def text_embeds(sd, input_ids, attention_mask, *, heads, eps=1e-5):
    """sd: `OwlViTModel.state_dict()` entries (text_model.*, text_projection.weight); input_ids [N, S] int64;
    attention_mask [N, S] (1 = token, 0 = padding) or None.  Returns unit-norm text_embeds [N, E] fp32."""
    ids = input_ids.reshape(-1, input_ids.shape[-1]).long()
    N, S = ids.shape
    e = "text_model.embeddings."
    x = sd[e + "token_embedding.weight"][ids] + sd[e + "position_embedding.weight"][:S][None]     # HF:370-373
    D = x.shape[-1]
    dh = D // heads
    # HF:661-666 create_causal_mask: key j visible to query i iff j <= i and attention_mask[n, j] != 0
    visible = torch.ones(S, S, dtype=torch.bool).tril()[None].expand(N, S, S).clone()
    if attention_mask is not None:
        visible &= attention_mask.reshape(N, 1, S).bool()
    bias = torch.zeros(N, 1, S, S).masked_fill(~visible[:, None], float("-inf"))
    i = 0
    while f"text_model.encoder.layers.{i}.layer_norm1.weight" in sd:
        p = f"text_model.encoder.layers.{i}."
        a = p + "self_attn."
        h = F.layer_norm(x, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)   # HF:498
        q = F.linear(h, sd[a + "q_proj.weight"], sd[a + "q_proj.bias"])                           # HF:439-441
        k = F.linear(h, sd[a + "k_proj.weight"], sd[a + "k_proj.bias"])
        v = F.linear(h, sd[a + "v_proj.weight"], sd[a + "v_proj.bias"])
        q, k, v = (t.view(N, S, heads, dh).transpose(1, 2) for t in (q, k, v))
        w = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5 + bias, -1)                        # HF:393-399
        w = torch.nan_to_num(w)                      # a query with no visible key (never the pooled row)
        ctx = (w @ v).transpose(1, 2).reshape(N, S, D)                                            # HF:401-404
        x = x + F.linear(ctx, sd[a + "out_proj.weight"], sd[a + "out_proj.bias"])                 # HF:459, 504
        h = F.layer_norm(x, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)   # HF:507
        h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])                         # HF:474
        h = h * torch.sigmoid(1.702 * h)                                                         # quick_gelu
        x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])                     # HF:476, 509
        i += 1
    x = F.layer_norm(x, (D,), sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"], eps)
    pooled = x[torch.arange(N), ids.argmax(-1)]                                                  # HF:677-684
    emb = F.linear(pooled, sd["text_projection.weight"])                                         # HF:978
    return emb / torch.linalg.norm(emb, ord=2, dim=-1, keepdim=True)                             # HF:984


def synthetic_prompts(n_prompts, seq=16, vocab=49408, seed=5):
    """Token ids shaped like the CLIP tokenizer's output: <bos> words... <eos> then padding 0 (the OWL-ViT processor
    pads with id 0), with the matching attention mask.  <eos> = vocab - 1 is the largest id of every row (HF:680)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(n_prompts, seq, dtype=torch.int64)
    mask = torch.zeros(n_prompts, seq, dtype=torch.int64)
    for n in range(n_prompts):
        words = int(torch.randint(1, seq - 1, (1,), generator=g))
        ids[n, 0] = vocab - 2
        ids[n, 1:1 + words] = torch.randint(1, vocab - 2, (words,), generator=g)
        ids[n, 1 + words] = vocab - 1
        mask[n, :words + 2] = 1
    return ids, mask
