"""The text-tower oracle (oracle/text_oracle.py) against the REAL HuggingFace classes the reference calls at
src/models.py:152-169 (`OwlViTForObjectDetection(**inputs).text_embeds`).  CPU only."""
import torch


def _hf_model(seed=0, sharpen=4.0):
    """Default OWL-ViT text tower (12 layers, hidden 512, 8 heads) next to a tiny vision tower (text_embeds do not
    depend on the image).  Random init leaves the attention almost uniform, so q / k projections are scaled up: the
    causal + padding mask then decides which keys carry the weight."""
    from transformers import OwlViTConfig, OwlViTForObjectDetection
    torch.manual_seed(seed)
    cfg = OwlViTConfig(vision_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1,
                                          num_attention_heads=1, image_size=64, patch_size=32))
    hf = OwlViTForObjectDetection._from_config(cfg, attn_implementation="eager").eval()
    with torch.no_grad():
        for layer in hf.owlvit.text_model.encoder.layers:
            layer.self_attn.q_proj.weight.mul_(sharpen)
            layer.self_attn.k_proj.weight.mul_(sharpen)
            layer.self_attn.q_proj.bias.normal_(0, 0.5)
            layer.mlp.fc1.bias.normal_(0, 0.5)
    return hf


def hf_text_embeds(hf, ids, mask):
    with torch.no_grad():
        out = hf(input_ids=ids, attention_mask=mask, pixel_values=torch.zeros(1, 3, 64, 64))
    return out.text_embeds[0]


def test_text_oracle_matches_huggingface():
    from oracle import text_oracle
    hf = _hf_model()
    ids, mask = text_oracle.synthetic_prompts(24)
    ref = hf_text_embeds(hf, ids, mask)
    sd = {k: v.detach() for k, v in hf.owlvit.state_dict().items()}
    got = text_oracle.text_embeds(sd, ids, mask, heads=8, eps=hf.config.text_config.layer_norm_eps)
    assert ref.shape == got.shape == (24, 512)
    assert torch.allclose(ref.norm(dim=-1), torch.ones(24), atol=1e-5)
    assert (ref - got).abs().max().item() <= 2e-6
    # the mask matters in this fixture: ignoring it must move the result
    nomask = text_oracle.text_embeds(sd, ids, None, heads=8)
    assert (nomask - got).abs().max().item() <= 2e-6      # padding sits AFTER <eos>: causality already hides it
    shuffled = ids.clone()
    shuffled[:, 1] = ids[:, 1].roll(1)
    assert (text_oracle.text_embeds(sd, shuffled, mask, heads=8) - got).abs().max().item() > 1e-3
