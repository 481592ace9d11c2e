"""ORACLE-side helper (test infrastructure, NOT product code): the REAL reference classes as a baseline arm.

The reference (stevebottos/owl-vit-object-detection) is pure Python with no setup.py / pyproject, so there is nothing
`pip install` could build.  Its equivalent of "compile the reference from the sources where they lie" is byte-compiling
the three modules of the hot path,

    /root/reference/src/models.py   /root/reference/src/matcher.py   /root/reference/src/losses.py

into sourceless byte-code files (`*.refbin`, the `.pyc` format) under `oracle/_ref/src/` (`install()`, called by `__graft_entry__.build()` when
/root/reference is present).  `oracle/_ref/` is git-ignored (no reference source or binary enters the history) but not
gpurun-ignored, so the compiled modules travel to the GPU box like our own built `.so`; the box has the same image,
hence the same interpreter and the same third-party packages the reference imports (torch, transformers, scipy,
torchvision).  Nothing here is imported by the product path.

Only `bench.py`'s baseline legs (`--impl reference`, `cpu_baseline`, `torch_cuda_baseline`, the matcher reference
timing) and `tests/` use this module.  `available()` says whether the real classes can be loaded; when they cannot,
bench.py falls back to the oracle port and says `kind: "port"`.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import os
import py_compile
import sys
import types
from typing import Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
COMPILED = os.path.join(_HERE, "_ref", "src")
MODULES = ("matcher", "losses", "models")        # import order: losses imports matcher
EXT = ".refbin"                                  # CPython byte-code (.pyc format); not named *.pyc so that no
                                                 # "ignore compiled python" rule of a sync tool drops it on the way
_PKG = "_owl_reference_src"                      # private package name: never collides with our own `src` mirror
_loaded = None


def install(verbose: bool = False) -> bool:
    """Byte-compile the reference's hot-path modules into oracle/_ref/src/*.refbin.  No-op without /root/reference."""
    if not os.path.isdir(REF_SRC):
        return False
    os.makedirs(COMPILED, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(REF_SRC, m + ".py"), cfile=os.path.join(COMPILED, m + EXT),
                           dfile=f"reference/src/{m}.py", doraise=True)
        if verbose:
            print(f"oracle/_ref/src/{m}{EXT} <- {REF_SRC}/{m}.py")
    return True


def _where() -> Optional[Tuple[str, str]]:
    if os.path.isdir(REF_SRC) and all(os.path.exists(os.path.join(REF_SRC, m + ".py")) for m in MODULES):
        return REF_SRC, ".py"
    if all(os.path.exists(os.path.join(COMPILED, m + EXT)) for m in MODULES):
        return COMPILED, EXT
    return None


def available() -> bool:
    return _where() is not None


def load():
    """(models, losses, matcher) modules of the REAL reference.  The reference's modules import each other as
    `src.matcher`; they are loaded under a private package and `src.matcher` is aliased only while `losses` is being
    imported, so our own drop-in mirror (/root/repo/src) stays importable as `src` before and after."""
    global _loaded
    if _loaded is not None:
        return _loaded
    w = _where()
    if w is None:
        raise RuntimeError("the reference is not available: neither /root/reference/src nor oracle/_ref/src/*.refbin")
    folder, ext = w
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    pkg = types.ModuleType("src")
    pkg.__path__ = []
    sys.modules["src"] = pkg
    mods = {}
    try:
        for m in MODULES:
            path = os.path.join(folder, m + ext)
            name = "src." + m
            loader = (importlib.machinery.SourceFileLoader(name, path) if ext == ".py"
                      else importlib.machinery.SourcelessFileLoader(name, path))
            spec = importlib.util.spec_from_loader(name, loader, origin=path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            loader.exec_module(mod)
            setattr(pkg, m, mod)
            mods[m] = mod
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    for m, mod in mods.items():
        sys.modules[f"{_PKG}.{m}"] = mod
    _loaded = (mods["models"], mods["losses"], mods["matcher"])
    return _loaded


def build_model(cfg, sd, attn_implementation: str = "eager"):
    """The reference `OwlViT` (src/models.py:41-119) around a HuggingFace `OwlViTForObjectDetection` carrying the
    synthetic weights `sd`, with the reference freeze loop (src/models.py:173-184) applied verbatim.  The only shim is
    SURVEY D7 (transformers 5.5.0: `compute_box_bias(h, w)` instead of 4.30.2's `compute_box_bias(feature_map)`)."""
    import torch
    from transformers import OwlViTConfig, OwlViTForObjectDetection
    rmodels, _, _ = load()
    vc = dict(image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_size=cfg.hidden,
              num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads, intermediate_size=cfg.ff)
    hf_cfg = OwlViTConfig(vision_config=vc, text_config=dict(hidden_size=cfg.embed), projection_dim=cfg.embed)
    hf = OwlViTForObjectDetection._from_config(hf_cfg, attn_implementation=attn_implementation)
    orig = hf.compute_box_bias
    # D7 shim: 4.30.2's `compute_box_bias(feature_map)` built the same [P, 4] constant and moved it to feature_map.device
    hf.compute_box_bias = lambda fm: orig(fm.shape[1], fm.shape[2]).to(fm.device)
    model = rmodels.OwlViT(hf, sd["queries"].clone())
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing
    for name, p in model.named_parameters():                                  # reference src/models.py:173-184
        if any(["layers.11" in name, "box" in name, "post_layernorm" in name,
                "class_predictor" in name, "queries" in name]):
            continue
        p.requires_grad = False
    return model


def criterion(n_classes: int, scales):
    _, rlosses, _ = load()
    return rlosses.PushPullLoss(n_classes, scales)


def matcher(n_classes: int):
    _, _, rmatcher = load()
    return rmatcher.HungarianMatcher(n_classes)
