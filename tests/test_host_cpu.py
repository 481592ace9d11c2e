"""CPU-only checks (no GPU needed): the C-ABI library loads and exports every symbol include/owl_b200.h declares,
the flat parameter layout, the reference-facing module surface, and the data-parallel gradient all-reduce over
gloo with world_size 2."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from owl_vit_object_detection_b200 import synth  # noqa: E402
from owl_vit_object_detection_b200.params import ParamLayout  # noqa: E402


def test_cabi_exports_every_declared_symbol():
    import __graft_entry__ as g
    from owl_vit_object_detection_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "owl_b200.h")).read()
    names = sorted(set(re.findall(r"\b(owl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20, names
    for n in names:
        assert hasattr(lib, n), f"libowl_b200.so does not export {n}"
    assert lib.owl_abi_version() == 6
    # argument validation works without a GPU and reports through owl_last_error
    assert lib.owl_gemm(None, None) != 0
    assert b"null" in lib.owl_last_error()


def test_header_is_plain_c_and_new_entry_points_validate_without_a_gpu():
    """include/owl_b200.h must be consumable from C (the FFI of a reference maintainer: cgo / ctypes / cffi), and the
    entry points added in round 2 reject bad arguments before touching a device (so this runs without a GPU)."""
    import ctypes
    from owl_vit_object_detection_b200 import _lib, ops
    hdr = os.path.join(ROOT, "include", "owl_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    lib = _lib.lib()
    one = ctypes.c_void_p(16)          # a non-null dummy: validation fails before it is dereferenced
    assert lib.owl_text_attn(one, None, one, 2, 33, 8, 64, ctypes.c_float(0.125), None) != 0
    assert b"33 tokens" in lib.owl_last_error()
    assert lib.owl_text_attn(one, None, one, 2, 16, 8, 32, ctypes.c_float(0.125), None) != 0
    assert b"head_dim" in lib.owl_last_error()
    assert lib.owl_text_embed(one, one, one, one, 4, 16, 510, 49408, None, None) != 0
    assert b"multiple of 4" in lib.owl_last_error()
    assert lib.owl_colsum(one, 1, ctypes.c_longlong(8), 4, 8, None, one, one, None) != 0
    assert b"fp32 input" in lib.owl_last_error()
    # host-only helper: workspace of a ragged batch = sum over the images; an empty side is refused
    imgs = (ops.PreImage * 2)()
    imgs[0].pixels, imgs[0].H, imgs[0].W, imgs[0].row_stride_bytes = 16, 480, 640, 1920
    imgs[1].pixels, imgs[1].H, imgs[1].W, imgs[1].row_stride_bytes = 16, 90, 120, 360
    fn, one_img = lib.owl_preprocess_batch_workspace_bytes, lib.owl_preprocess_workspace_bytes
    fn.restype = one_img.restype = ctypes.c_longlong
    assert fn(imgs, 2, 768) == one_img(480, 640, 768) + one_img(90, 120, 768) > 3 * 480 * 768
    imgs[1].W = 0
    assert fn(imgs, 2, 768) == -1
    assert lib.owl_preprocess_batch(imgs, 2, one, one, 768, one, ctypes.c_longlong(1 << 30), None) != 0


def test_layout_groups_trainables_and_qkv():
    for cfg in (synth.B32, synth.TINY, synth.L14):
        L = ParamLayout(cfg)
        train = synth.trainable_names(cfg)
        assert all(L.offsets[n] >= L.train_begin for n in train)
        assert all(L.offsets[n] < L.train_begin for n in L.shapes if n not in train)
        for i in range(cfg.layers):
            p = f"backbone.encoder.layers.{i}.self_attn."
            lo, hi = L.span(p + "q_proj.weight", p + "v_proj.weight")
            assert hi - lo == 3 * cfg.hidden * cfg.hidden
            assert L.offsets[p + "k_proj.weight"] == lo + cfg.hidden * cfg.hidden
        assert L.train_begin % 64 == 0 and L.total % 64 == 0
    L = ParamLayout(synth.B32)
    assert len(synth.trainable_names(synth.B32)) == 29
    assert sum(L._numel(n) for n in synth.trainable_names(synth.B32)) == 8_791_812


def test_module_surface_matches_reference_names():
    from src.models import OwlViT
    cfg = synth.TINY
    sd = synth.make_weights(cfg, seed=1)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg)
    assert list(model.state_dict().keys()).sort() == list(synth.param_shapes(cfg).keys()).sort()
    assert {n for n, p in model.named_parameters() if p.requires_grad} == set(synth.trainable_names(cfg))
    for n, p in model.named_parameters():
        assert torch.equal(p.detach(), sd[n]), n
    # in-place optimizer updates write through to the flat buffer
    with torch.no_grad():
        model._param("queries").add_(1.0)
    assert torch.equal(model.layout.view(model.flat_params, "queries"), sd["queries"] + 1.0)
    # load_state_dict with reference keys works
    model.load_state_dict(sd)
    assert torch.equal(model.layout.view(model.flat_params, "queries"), sd["queries"])
    # no CPU fallback: the product path fails loudly
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, cfg.image_size, cfg.image_size))


def test_hf_constructor_path():
    """`OwlViT(pretrained_model=<HF OwlViTForObjectDetection>, query_bank=...)` like reference src/models.py:171."""
    from transformers import OwlViTConfig, OwlViTForObjectDetection
    from src.models import OwlViT
    hc = OwlViTConfig(vision_config=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                                         num_attention_heads=2, image_size=128, patch_size=32),
                      text_config=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1,
                                       num_attention_heads=2), projection_dim=128)
    hf = OwlViTForObjectDetection(hc)
    q = torch.nn.functional.normalize(torch.randn(1, 24, 128), dim=-1)
    model = OwlViT(pretrained_model=hf, query_bank=q)
    assert model.cfg.hidden == 128 and model.cfg.layers == 2 and model.cfg.n_classes == 8
    w = hf.owlvit.vision_model.encoder.layers[1].mlp.fc1.weight
    assert torch.equal(model._param("backbone.encoder.layers.1.mlp.fc1.weight").detach(), w.detach())
    assert torch.equal(model._param("class_predictor.dense0.weight").detach(), hf.class_head.dense0.weight.detach())


_DDP_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from owl_vit_object_detection_b200 import synth
from src.models import OwlViT
rank = int(os.environ["RANK"])
dist.init_process_group("gloo")
cfg = synth.TINY
sd = synth.make_weights(cfg, seed=1)
model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg)
model._prepare_grads()                                          # what backward() does first: .grad = views
g = model.flat_grad
g.copy_(torch.arange(g.numel(), dtype=torch.float32) * (rank + 1))
model.allreduce_grads()
expect = torch.arange(g.numel(), dtype=torch.float32) * 3      # ranks 1x + 2x
assert torch.equal(model.flat_grad, expect), "all-reduce(sum) over the flat gradient buffer"
p = model._param("box_head.dense2.bias")
o = model.layout.offsets["box_head.dense2.bias"] - model.layout.train_begin
assert torch.equal(p.grad, expect[o:o + 4])
# the bucketed reduction TrainStep overlaps with the backward pass: three ranges, in completion order, tile the buffer
buckets = model.layout.grad_buckets()
assert sorted(buckets) == [(buckets[2][0], buckets[2][1]), (buckets[1][0], buckets[1][1]), (buckets[0][0], buckets[0][1])]
assert buckets[2][0] == 0 and buckets[2][1] == buckets[1][0] and buckets[1][1] == buckets[0][0] and buckets[0][1] == g.numel()
g.copy_(torch.arange(g.numel(), dtype=torch.float32) * (rank + 1))
for lo, hi in buckets:
    dist.all_reduce(model.flat_grad[lo:hi], op=dist.ReduceOp.SUM)
assert torch.equal(model.flat_grad, expect), "bucket by bucket == one all-reduce"
dist.destroy_process_group()
sys.stdout.write("rank " + str(rank) + " ok\n")
"""


def test_allreduce_flat_grads_gloo_world2(tmp_path):
    script = tmp_path / "ddp_worker.py"
    script.write_text(_DDP_WORKER % ROOT)
    port = 29500 + (os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2, r.stdout


def test_shadow_version_tracks_parameters_after_module_to():
    """The fp16 GEMM operands are refreshed when `Engine.version()` changes.  After `Module.to()` the parameters are
    re-pointed at the moved flat buffer through `.data =`, which leaves them with version counters of their own: an
    in-place torch.optim.AdamW step (the reference's optimizer, main.py:56-60,91) then changes the flat buffer's
    contents without touching its counter, so the version must also watch the parameters."""
    import torch
    from owl_vit_object_detection_b200 import synth
    from owl_vit_object_detection_b200.engine import alias_version
    from owl_vit_object_detection_b200.model import OwlViT
    cfg = synth.TINY
    sd = synth.make_weights(cfg, seed=1)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg)
    model._apply(lambda t: t.clone())                       # what .to(device) does: a new flat buffer
    flat = model.flat_params
    params = [model._param(n) for n in model._names]
    assert all(p.data_ptr() >= flat.data_ptr() and p.data_ptr() < flat.data_ptr() + flat.numel() * 4 for p in params)
    before_values = flat.clone()
    v_flat, v_all = flat._version, alias_version(flat, params)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.1)
    for p in model.parameters():
        if p.requires_grad:
            p.grad = torch.ones_like(p)
    opt.step()
    assert not torch.equal(flat, before_values), "the optimizer writes through the views into the flat buffer"
    assert alias_version(flat, params) != v_all, "Engine.version() must see an in-place torch.optim step"
    if flat._version == v_flat:
        # the case the watch exists for: set_data() un-shared the counters
        assert any(p._version > 0 for p in params)
    # a model that was never moved shares one counter between the flat buffer and its parameter views
    model2 = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg)
    p0 = model2._param("queries")
    v0 = alias_version(model2.flat_params, [p0])
    with torch.no_grad():
        p0.mul_(0.5)
    assert alias_version(model2.flat_params, [p0]) != v0


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, smoke() and bench.py's baseline legs may import it.  Static check of
    every product module (the package except its smoke test, and the reference-facing mirrors under src/), plus a
    dynamic one: importing the whole product leaves no `oracle` module loaded."""
    import ast
    offenders = []
    for base in ("owl_vit_object_detection_b200", "src"):
        for fn in sorted(os.listdir(os.path.join(ROOT, base))):
            if not fn.endswith(".py") or (base, fn) == ("owl_vit_object_detection_b200", "smoke.py"):
                continue
            tree = ast.parse(open(os.path.join(ROOT, base, fn)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                offenders += [f"{base}/{fn}: {m}" for m in mods if m.split(".")[0] == "oracle"]
    assert not offenders, offenders
    code = ("import sys; sys.path.insert(0, %r); import src.models, src.losses, src.matcher, src.train_util, src.util; "
            "import owl_vit_object_detection_b200.train, owl_vit_object_detection_b200.text, "
            "owl_vit_object_detection_b200.preprocess, owl_vit_object_detection_b200.collective; "
            "bad = [m for m in sys.modules if m.split('.')[0] == 'oracle']; assert not bad, bad" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])
