"""Generates tests/golden/*.npz by running the REAL reference (read-only at /root/reference).

Run here (the build container), never on the GPU box: `python tests/golden/make_golden.py`.
The reference classes are imported unmodified; the only shim is SURVEY D7 (transformers 5.5.0
changed `compute_box_bias(feature_map)` to `compute_box_bias(h, w)`).  Inputs are the seeded
tensors of owl_vit_object_detection_b200.synth, so the fixtures only need to hold OUTPUTS.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from owl_vit_object_detection_b200 import synth  # noqa: E402


def load_reference():
    """Import the REFERENCE's `src` package (a namespace package: it has no __init__.py, so our own drop-in mirror
    /root/repo/src - a regular package - would win on sys.path no matter the order).  Bind `src` to the reference
    directory explicitly."""
    import importlib
    import types
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    pkg = types.ModuleType("src")
    pkg.__path__ = [os.path.join(REF, "src")]
    sys.modules["src"] = pkg
    rmodels = importlib.import_module("src.models")
    rlosses = importlib.import_module("src.losses")
    rmatcher = importlib.import_module("src.matcher")
    for m in (rmodels, rlosses, rmatcher):
        assert m.__file__.startswith(REF), m.__file__
    return rmodels, rlosses, rmatcher


def build_reference_model(rmodels, cfg, sd):
    from transformers import OwlViTConfig, OwlViTForObjectDetection
    assert cfg == synth.B32
    hf = OwlViTForObjectDetection._from_config(OwlViTConfig(), attn_implementation="eager")
    orig = hf.compute_box_bias
    hf.compute_box_bias = lambda fm: orig(fm.shape[1], fm.shape[2]).to(fm.device)      # D7 shim
    model = rmodels.OwlViT(hf, sd["queries"].clone())
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing
    # reference freeze loop, reference src/models.py:173-184, verbatim conditions
    for name, p in model.named_parameters():
        if any(["layers.11" in name, "box" in name, "post_layernorm" in name,
                "class_predictor" in name, "queries" in name]):
            continue
        p.requires_grad = False
    return model


def sub(t: torch.Tensor) -> np.ndarray:
    """Full tensor if small, else a strided subsample of its 2-D view [shape[0]-or-rows, -1]."""
    return synth.subsample(t.detach()).numpy().copy()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    rmodels, rlosses, rmatcher = load_reference()
    cfg = synth.B32
    sd = synth.make_weights(cfg, seed=0)

    # ---------------- model forward + train-step gradients (config 1: batch 1, CPU fp32) -------------
    model = build_reference_model(rmodels, cfg, sd)
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert n_train == 8_791_812, n_train
    image = synth.make_images(cfg, 2, seed=2)
    labels, tboxes, nt = synth.make_targets(cfg, 2, seed=3)
    scales = synth.make_class_scales(cfg)
    out = {}
    crit = rlosses.PushPullLoss(cfg.n_classes, scales)
    grads_acc = None
    for b in range(2):
        model.zero_grad()
        t = int(nt[b])
        boxes, _, sims, _ = model(image[b:b + 1])
        out[f"boxes{b}"] = boxes.detach().numpy()[0].copy()
        out[f"sims{b}"] = sims.detach().numpy()[0].copy()
        losses = crit(sims, labels[b:b + 1, :t], boxes, tboxes[b:b + 1, :t])
        for k, v in losses.items():
            out[f"{k}{b}"] = np.float32(v.item())
        (losses["loss_ce"] + losses["loss_bg"] + losses["loss_bbox"] + losses["loss_giou"]).backward()
        if b == 0:
            for name, p in model.named_parameters():
                if p.requires_grad:
                    out["grad0." + name] = sub(p.grad)
                    out["gnorm0." + name] = np.float32(p.grad.norm().item())
    np.savez_compressed(os.path.join(HERE, "model_b32.npz"), **out)
    print("model_b32.npz:", {k: v for k, v in out.items() if k.startswith("loss")})

    # ---------------- matcher (config 5 inputs, 6 images per T) --------------------------------------
    for T in (10, 50, 100):
        n = 6
        sims, pred, lab, tgt = synth.make_matcher_inputs(n, T, seed=4)
        matcher = rmatcher.HungarianMatcher(80)
        mo = {}
        for b in range(n):
            tc, ind, _ = matcher({"pred_logits": sims[b:b + 1], "pred_boxes": pred[b:b + 1]},
                                 [{"labels": lab[b], "boxes": tgt[b]}])
            mo[f"tc{b}"] = tc[0].numpy().astype(np.int16)
            mo[f"pred_idx{b}"] = ind[0][0].numpy().astype(np.int16)
            mo[f"tgt_idx{b}"] = ind[0][1].numpy().astype(np.int16)
        # full reference loss (+ grads wrt sims / boxes) on the first two images
        for b in range(2):
            s = sims[b:b + 1].clone().requires_grad_(True)
            p = pred[b:b + 1].clone().requires_grad_(True)
            crit = rlosses.PushPullLoss(80, scales)
            losses = crit(s, lab[b:b + 1], p, tgt[b:b + 1])
            for k, v in losses.items():
                mo[f"{k}{b}"] = np.float32(v.item())
            (losses["loss_ce"] + losses["loss_bg"] + losses["loss_bbox"] + losses["loss_giou"]).backward()
            if b == 0:
                mo[f"dsims{b}"] = s.grad[0].numpy().astype(np.float32)
                mo[f"dboxes{b}"] = p.grad[0].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"matcher_T{T}.npz"), **mo)
        print(f"matcher_T{T}.npz written")


if __name__ == "__main__":
    main()
