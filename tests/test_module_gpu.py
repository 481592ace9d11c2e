"""End-to-end through the reference-facing API (src.models.OwlViT / src.losses.PushPullLoss, the classes
reference main.py:42-91 drives): forward, loss, backward, optimizer step, CUDA-graph capture, smoke()."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import matcher_oracle as mo  # noqa: E402  (checker only)
from owl_vit_object_detection_b200 import synth  # noqa: E402


def _model(cfg, seed=1):
    from src.models import OwlViT
    sd = synth.make_weights(cfg, seed=seed)
    return OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg).to("cuda"), sd


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_train_step_like_reference_main():
    """reference main.py:74-91 with torch.optim.AdamW over model.parameters(), batch 1, unpadded targets."""
    from src.losses import PushPullLoss
    cfg = synth.TINY
    model, sd = _model(cfg)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    assert names == synth.trainable_names(cfg) or set(names) == set(synth.trainable_names(cfg))
    assert set(model.state_dict().keys()) == set(synth.param_shapes(cfg).keys())
    crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).cuda())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.1)
    img = synth.make_images(cfg, 1, seed=5).cuda()
    labels, tboxes, nt = synth.make_targets(cfg, 1, seed=3, max_t=8)
    t = int(nt[0])
    model.train()
    first = None
    for step in range(3):
        opt.zero_grad()
        boxes, _, sims, _ = model(img)
        losses = crit(sims, labels[:, :t].cuda(), boxes, tboxes[:, :t].cuda())
        loss = losses["loss_ce"] + losses["loss_bg"] + losses["loss_bbox"] + losses["loss_giou"]
        loss.backward()
        if step == 0:
            ref, _, _ = mo.push_pull_loss(sims.detach().cpu(), boxes.detach().cpu(), [labels[0, :t]], [tboxes[0, :t]],
                                          cfg.n_classes, synth.make_class_scales(cfg))
            for k in ref:
                np.testing.assert_allclose(losses[k].item(), ref[k].item(), rtol=1e-4, atol=1e-6)
            first = loss.item()
            g = model._param("box_head.dense0.weight").grad
            assert g is not None and g.abs().max().item() > 0
        opt.step()
    assert loss.item() < first, "three AdamW steps on one image should reduce its loss"
    crit.check_status()


def test_eval_forward_no_grad_and_postprocess():
    from src.models import PostProcess
    cfg = synth.TINY
    model, _ = _model(cfg)
    model.eval()
    img = synth.make_images(cfg, 1, seed=6).cuda()
    with torch.no_grad():
        boxes, none1, sims, none2 = model(img)
    assert none1 is None and none2 is None and not boxes.requires_grad
    assert boxes.shape == (1, cfg.patches, 4) and sims.shape == (1, cfg.patches, cfg.n_classes)
    b, c, s = PostProcess(confidence_threshold=-1.0, iou_threshold=0.6)(boxes, sims)
    assert b.shape[0] == 1 and b.shape[2] == 4 and c.shape == s.shape


def test_cuda_graph_step_matches_eager():
    """The whole step (forward, matcher, loss, backward, fused AdamW) is capturable: no syncs, no allocations."""
    from src.losses import PushPullLoss
    from src.models import FusedAdamW
    cfg = synth.TINY
    B = 2
    img = synth.make_images(cfg, B, seed=5).cuda()
    labels, tboxes, nt = [x.cuda() for x in synth.make_targets(cfg, B, seed=3, max_t=8)]
    scales = synth.make_class_scales(cfg).cuda()

    def build():
        model, _ = _model(cfg)
        return model, PushPullLoss(cfg.n_classes, scales), FusedAdamW(model, lr=1e-3, weight_decay=0.1)

    def step(model, crit, opt):
        opt.zero_grad(set_to_none=False)
        boxes, _, sims, _ = model(img)
        l = crit(sims, labels, boxes, tboxes, num_targets=nt)
        (l["loss_ce"] + l["loss_bg"] + l["loss_bbox"] + l["loss_giou"]).backward()
        opt.step()
        return torch.stack([l[k].detach() for k in ("loss_ce", "loss_bg", "loss_bbox", "loss_giou")])

    m1, c1, o1 = build()
    eager = [step(m1, c1, o1).cpu() for _ in range(4)]
    m2, c2, o2 = build()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        first = step(m2, c2, o2).cpu()            # warm-up (allocates buffers), counts as step 1
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            out = step(m2, c2, o2)
    torch.cuda.current_stream().wait_stream(s)
    got = [first]
    for _ in range(3):
        graph.replay()
        got.append(out.cpu().clone())
    torch.cuda.synchronize()
    np.testing.assert_allclose(got[0].numpy(), eager[0].numpy(), rtol=1e-5)
    np.testing.assert_allclose(got[1].numpy(), eager[1].numpy(), rtol=1e-3)


def test_trainstep_host_buffers_graph_matches_eager():
    """`TrainStep` (the bench / training driver): batches arrive from pinned HOST memory into rotating input slots,
    the single-GPU step is ONE graph replay (forward, matcher + loss, backward, AdamW), and the four losses of step i
    are read from pinned memory (`result`) after step i + 1 has been queued.  Same numbers as the eager step."""
    from src.losses import PushPullLoss
    from src.models import FusedAdamW
    from owl_vit_object_detection_b200.train import TrainStep
    cfg = synth.TINY
    B, n_slots, n_steps = 2, 2, 5
    scales = synth.make_class_scales(cfg).cuda()
    host = []
    for s in range(3):
        img = synth.make_images(cfg, B, seed=40 + s).pin_memory()
        lab, box, nt = synth.make_targets(cfg, B, seed=50 + s, max_t=8)
        host.append((img, lab.pin_memory(), box.pin_memory(), nt.pin_memory()))

    def drive(use_graph):
        model, _ = _model(cfg)
        crit = PushPullLoss(cfg.n_classes, scales)
        opt = FusedAdamW(model, lr=1e-3, weight_decay=0.1)
        step = TrainStep(model, crit, opt, batch=B, max_targets=host[0][1].shape[1], use_graph=use_graph,
                         n_input_slots=n_slots)
        for s in range(n_slots):
            step.load(*host[s], slot=s)
        torch.cuda.synchronize()
        step.warmup()                 # captures on whatever the slots hold and restores the optimizer state
        out, pending = [], None
        nxt = step.load(*host[0])
        for i in range(n_steps):
            cur = nxt
            if i + 1 < n_steps:
                nxt = step.load(*host[(i + 1) % len(host)])
            step.run(slot=cur, readback=True)
            if pending is not None:
                out.append(step.result(pending))
            pending = cur
        out.append(step.result(pending))
        torch.cuda.synchronize()
        crit.check_status()
        return np.array(out), model.flat_params.detach().float().cpu().numpy().copy()

    l_graph, p_graph = drive(True)
    l_eager, p_eager = drive(False)
    assert l_graph.shape == (n_steps, 4) and np.isfinite(l_graph).all()
    np.testing.assert_allclose(l_graph[0], l_eager[0], rtol=1e-5)
    np.testing.assert_allclose(l_graph, l_eager, rtol=2e-3)           # atomics order differs from step 2 on
    np.testing.assert_allclose(p_graph, p_eager, rtol=0, atol=2e-4)
    assert not np.allclose(l_graph[0], l_graph[1]), "different batches must give different losses"


def _per_tensor_report(model, a, b, tol):
    """[(name, elements, elements off by more than tol, max |diff|)] over the trainable tensors of two flat parameter vectors."""
    L = model.layout
    rows = []
    for n in L.trainable:
        o, k = L.offsets[n], L._numel(n)
        d = np.abs(a[o:o + k] - b[o:o + k])
        rows.append((n, k, int((d > tol).sum()), float(d.max())))
    return rows


def test_fused_adamw_matches_torch_adamw():
    """SURVEY R14 (reference main.py:56-60,91): the fused AdamW kernel over the flat buffers against
    torch.optim.AdamW over model.parameters() AFTER `.to("cuda")` (the reference's own construction order) - same
    gradients (our backward), same hyper-parameters, three steps.  Strict, with a per-tensor report.

    Bars: every trainable tensor within 2e-5 after three steps at lr 1e-3 (measured <= 1.5e-5: the two optimizers
    round differently in the last bit, the next forward amplifies that through fp16 rounding of the operands).
    `k_proj.bias` is bounded separately: its gradient is mathematically zero (softmax is invariant to a per-query
    shift of the scores, tests/test_backward_gpu.py docstring), so what reaches Adam is rounding noise whose SIGN
    decides a full +-lr step; the reference's autograd has the same property.  It must stay within the 3 * lr an
    Adam step sequence can move a weight (+ weight decay), nothing tighter is meaningful.
    Also asserted: the fp16 GEMM operands equal the fp32 parameters at every forward (no stale shadow)."""
    from src.losses import PushPullLoss
    from src.models import FusedAdamW
    cfg = synth.TINY
    B, lr, steps = 2, 1e-3, 3
    img = synth.make_images(cfg, B, seed=15).cuda()
    labels, tboxes, nt = [x.cuda() for x in synth.make_targets(cfg, B, seed=13, max_t=8)]
    scales = synth.make_class_scales(cfg).cuda()

    def run(fused):
        model, _ = _model(cfg)
        crit = PushPullLoss(cfg.n_classes, scales)
        opt = (FusedAdamW(model, lr=lr, weight_decay=0.1) if fused
               else torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=0.1))
        start = model.flat_params.detach().clone()
        lo = model.layout.train_begin
        for _ in range(steps):
            if fused:
                opt.zero_grad(set_to_none=False)
            else:
                opt.zero_grad()
            boxes, _, sims, _ = model(img)
            stale = int((model.engine.flat16[lo:] != model.flat_params[lo:].half()).sum().item())
            assert stale == 0, f"{stale} fp16 GEMM operand elements do not follow the fp32 parameters"
            l = crit(sims, labels, boxes, tboxes, num_targets=nt)
            (l["loss_ce"] + l["loss_bg"] + l["loss_bbox"] + l["loss_giou"]).backward()
            opt.step()
        torch.cuda.synchronize()
        return model, start.cpu().numpy(), model.flat_params.detach().cpu().numpy().copy()

    model, s0, p_fused = run(True)
    _, s1, p_torch = run(False)
    assert np.array_equal(s0, s1)
    moved = np.abs(p_torch - s1).max()
    assert moved > 1e-3, "three steps at lr 1e-3 must move the trainable parameters"
    bad = []
    for name, k, n_off, mx in _per_tensor_report(model, p_fused, p_torch, 2e-5):
        print(f"  {name:58s} n={k:7d} off={n_off:6d} max|diff|={mx:.3e}")
        if name.endswith("k_proj.bias"):
            if mx > 2 * steps * lr * 1.1:
                bad.append((name, n_off, mx))
        elif n_off:
            bad.append((name, n_off, mx))
    assert not bad, bad
    frozen = np.abs(p_torch - s1) == 0
    assert np.array_equal(p_fused[frozen], s1[frozen]), "frozen parameters must not move"


def test_main_replay_three_steps_vs_reference_golden(golden_dir):
    """reference main.py:42-91 replayed literally on B/32: `OwlViT(...).to("cuda")`, `torch.optim.AdamW(
    model.parameters(), lr, weight_decay)`, then three times zero_grad / forward / PushPullLoss / backward / step
    (batch 1, images 8, 9, 8 of the synthetic set: chosen in make_golden_train3.py so that no discrete decision of the
    loss sits on its threshold) - against tests/golden/train3_b32.npz, produced by the REAL reference doing the same
    three CPU fp32 steps (tests/golden/make_golden_train3.py).

    What is compared, and why in this form.  The matcher and the IoU > 0.85 label sweep are DISCRETE: after an
    optimizer step our parameters differ from the reference's in the last bits (fp16 operands, atomics order), and a
    decision that sits on its threshold can then fall the other way and move `loss_ce` by 10 % (seen in ~1 of 4 runs with
    images 0, 1, 0, whose forward outputs nevertheless matched the golden to 8e-5 / 3e-4).  So from step 1 on the losses are checked against the ORACLE evaluated on our own forward outputs (which
    pins matcher + loss exactly), and the comparison with the real reference goes through the continuous quantities:
      * step 0 (identical parameters): the four losses within 2e-2 of the golden;
      * every step: `pred_sims` / `pred_boxes` within 1.5e-3 / 2e-3 (step 0), 2e-3 / 8e-3 (step 1), 3e-3 / 1.2e-2 (step 2)
        of the golden forward outputs of that step (measured 9e-5 / 3.5e-4, 6.6e-4 / 3.4e-3, 1.25e-3 / 5.9e-3: Adam turns
        fp16-level gradient noise into +-lr parameter steps, so the deviation grows with the step count).  For scale:
        the golden outputs of the revisited image move by up to 1.6e-2 / 0.26 between step 0 and step 2, which is
        what a forward that kept reading stale fp16 weights would be off by;
      * every step: our four losses == oracle(our sims, our boxes) to 2e-4;
      * the first image is revisited at step 2: its total loss must have dropped by >= 1.5 % (golden: 28.70 -> 27.22, -5.2 %; measured -3.2 %);
      * per trainable tensor, the displacement p_3 - p_0.  Adam's update is lr * m / (sqrt(v) + eps) ~ lr * sign(g) for
        the first steps: the MAGNITUDE of a gradient element is normalised away, so every element whose gradient is
        small against the fp16 path's gradient error (2e-3 of the tensor's max, i.e. 5-10 % of a typical element) takes
        a full +-lr step in a random direction.  The displacement therefore agrees with the reference only
        statistically: measured cosine 0.99 for the box head and the query bank, 0.76-0.82 for the tensors fed by the
        class loss, norms within 11 %.  Stated bars (sanity, not parity: a missing / doubled / wrong-sign update
        fails them): cosine >= 0.7, norm within 15 %, mean |diff| <= 0.45 * (3 lr), no element further than 2 * (3 lr).
        Parity of the optimizer itself is the strict per-tensor test above (identical gradients in, 2e-5), parity of
        the gradients is tests/test_backward_gpu.py.  `k_proj.bias` (zero gradient, pure noise) only has to stay
        inside 3 lr (1 + wd);
      * after every optimizer step the fp16 GEMM operands equal the fp32 parameters."""
    import os
    from src.losses import PushPullLoss
    gold = np.load(os.path.join(golden_dir, "train3_b32.npz"))
    lr, wd, steps = float(gold["lr"]), float(gold["weight_decay"]), int(gold["steps"])
    cfg = synth.B32
    model, _ = _model(cfg, seed=0)                                  # OwlViT(...).to("cuda")
    L = model.layout
    start = {n: model._param(n).detach().clone() for n in L.trainable}
    scales = synth.make_class_scales(cfg)
    crit = PushPullLoss(cfg.n_classes, scales.cuda())
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
    seq, n_img = [int(x) for x in gold["seq"]], int(gold["n_images"])
    image = synth.make_images(cfg, n_img, seed=2)
    labels, tboxes, nt = synth.make_targets(cfg, n_img, seed=3)
    model.train()
    lo = L.train_begin
    names = ("loss_ce", "loss_bg", "loss_bbox", "loss_giou")
    totals = []
    for step in range(steps):
        b = seq[step]
        t = int(nt[b])
        opt.zero_grad()
        boxes, _, sims, _ = model(image[b:b + 1].cuda())
        assert torch.equal(model.engine.flat16[lo:], model.flat_params[lo:].half()), "stale fp16 GEMM operands"
        losses = crit(sims, labels[b:b + 1, :t].cuda(), boxes, tboxes[b:b + 1, :t].cuda())
        (losses["loss_ce"] + losses["loss_bg"] + losses["loss_bbox"] + losses["loss_giou"]).backward()
        opt.step()
        sc, bc = sims.detach().cpu(), boxes.detach().cpu()
        es = np.abs(synth.subsample(sc[0]).numpy() - gold[f"sims{step}"]).max()
        eb = np.abs(synth.subsample(bc[0]).numpy() - gold[f"boxes{step}"]).max()
        ref, _, _ = mo.push_pull_loss(sc, bc, [labels[b, :t]], [tboxes[b, :t]], cfg.n_classes, scales)
        got = {k: losses[k].item() for k in names}
        print(f"  step {step}: forward err vs golden sims {es:.2e} boxes {eb:.2e}; losses "
              + ", ".join(f"{k} {got[k]:.4f} (golden {float(gold[f'{k}{step}']):.4f})" for k in names))
        assert es <= (1.5e-3, 2e-3, 3e-3)[step] and eb <= (2e-3, 8e-3, 1.2e-2)[step], (step, es, eb)
        for k in names:
            np.testing.assert_allclose(got[k], ref[k].item(), rtol=2e-4, atol=1e-6, err_msg=f"{k} step {step} vs oracle")
            if step == 0:
                np.testing.assert_allclose(got[k], float(gold[f"{k}{step}"]), rtol=2e-2, err_msg=f"{k} step 0 vs golden")
        totals.append(sum(got.values()))
    crit.check_status()
    assert totals[2] <= 0.985 * totals[0], f"revisited image: total loss {totals[0]:.3f} -> {totals[2]:.3f} after two updates"
    bad = []
    for n in L.trainable:
        d = synth.subsample((model._param(n).detach() - start[n]).cpu()).numpy().astype(np.float64).ravel()
        r = gold["disp." + n].astype(np.float64).ravel()
        diff = np.abs(d - r)
        if n.endswith("k_proj.bias"):
            ok = np.abs(d).max() <= steps * lr * (1 + wd) * 1.05
            print(f"  {n:58s} max|disp| {np.abs(d).max():.3e} (noise gradient: bounded only)")
        else:
            cos = float(d @ r / (np.linalg.norm(d) * np.linalg.norm(r) + 1e-30))
            full_norm = float((model._param(n).detach() - start[n]).norm().item())
            nrel = abs(full_norm - float(gold["dispnorm." + n])) / float(gold["dispnorm." + n])
            mean_rel, mx_rel = diff.mean() / (steps * lr), diff.max() / (steps * lr)
            ok = cos >= 0.7 and nrel <= 0.15 and mean_rel <= 0.45 and mx_rel <= 2.0
            print(f"  {n:58s} cos {cos:.5f} norm rel {nrel:.2e} mean|diff|/(3lr) {mean_rel:.3e} max {mx_rel:.3e}")
        if not ok:
            bad.append(n)
    assert not bad, bad


def test_subset_optimizer_grads_do_not_accumulate():
    """An optimizer that owns only the heads clears only their .grad (set_to_none): every step must still start
    those gradients from zero, although the backward kernels accumulate into the flat buffer."""
    from src.losses import PushPullLoss
    cfg = synth.TINY
    model, _ = _model(cfg)
    crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).cuda())
    heads = [p for n, p in model.named_parameters() if p.requires_grad and n.startswith("box_head")]
    opt = torch.optim.SGD(heads, lr=0.0)
    img = synth.make_images(cfg, 1, seed=5).cuda()
    labels, tboxes, nt = [x.cuda() for x in synth.make_targets(cfg, 1, seed=3, max_t=8)]
    grads = []
    for _ in range(2):
        opt.zero_grad()                           # clears the heads only
        boxes, _, sims, _ = model(img)
        l = crit(sims, labels, boxes, tboxes, num_targets=nt)
        (l["loss_ce"] + l["loss_bg"] + l["loss_bbox"] + l["loss_giou"]).backward()
        grads.append((model._param("box_head.dense0.weight").grad.clone(),
                      model._param("class_predictor.dense0.weight").grad.clone()))
    np.testing.assert_allclose(grads[1][0].cpu().numpy(), grads[0][0].cpu().numpy(), rtol=1e-4, atol=1e-7)
    # parameters nobody cleared keep accumulating (torch semantics)
    np.testing.assert_allclose(grads[1][1].cpu().numpy(), 2 * grads[0][1].cpu().numpy(), rtol=1e-3, atol=1e-7)


def test_backward_of_an_overwritten_forward_raises():
    """Saved activations live in one workspace per batch size: a backward whose forward is no longer the latest must
    raise instead of using the wrong activations; same for the criterion's buffers."""
    from src.losses import PushPullLoss
    cfg = synth.TINY
    model, _ = _model(cfg)
    crit = PushPullLoss(cfg.n_classes, None)
    img = synth.make_images(cfg, 1, seed=5).cuda()
    labels, tboxes, nt = [x.cuda() for x in synth.make_targets(cfg, 1, seed=3, max_t=8)]
    boxes, _, sims, _ = model(img)
    model(img)                                    # second forward, same batch size
    with pytest.raises(RuntimeError, match="another forward"):
        (sims.sum() + boxes.sum()).backward()
    boxes, _, sims, _ = model(img)
    l = crit(sims, labels, boxes, tboxes, num_targets=nt)
    crit(sims.detach(), labels, boxes.detach(), tboxes, num_targets=nt)
    with pytest.raises(RuntimeError, match="called again"):
        l["loss_ce"].backward()


def test_matcher_flags_out_of_range_labels_and_counts():
    """reference src/matcher.py:118 raises IndexError on a label >= n_classes; the kernels clamp and flag."""
    from src.matcher import HungarianMatcher
    g = torch.Generator().manual_seed(0)
    sims = (torch.rand((1, 64, 8), generator=g) * 0.4 - 0.1).cuda()
    xy = torch.rand((1, 64, 2), generator=g) * 0.5
    boxes = torch.cat([xy, xy + 0.1 + 0.3 * torch.rand((1, 64, 2), generator=g)], dim=-1).cuda()
    tb = torch.tensor([[0.1, 0.1, 0.4, 0.5], [0.3, 0.2, 0.9, 0.8]]).cuda()
    m = HungarianMatcher(8)
    m({"pred_logits": sims, "pred_boxes": boxes}, [{"labels": torch.tensor([1, 7]).cuda(), "boxes": tb}])
    with pytest.raises(IndexError):
        m({"pred_logits": sims, "pred_boxes": boxes}, [{"labels": torch.tensor([1, 8]).cuda(), "boxes": tb}])


def test_trainstep_raw_uint8_slots_equal_fp32_slots():
    """`TrainStep(raw_u8=True)` (raw RGB bytes in the input slots, rescale + normalise fused into the patch gather) must
    produce exactly the losses and parameters of the fp32 path fed with the reference's preprocessing of the same
    bytes (oracle: preprocess_oracle, pinned to PIL fixtures; at the model's resolution the resize is the identity)."""
    from oracle import preprocess_oracle as pre   # checker only
    from src.losses import PushPullLoss
    from src.models import FusedAdamW
    from owl_vit_object_detection_b200.train import TrainStep
    cfg = synth.TINY
    B = 2
    raw = synth.make_images_u8(cfg, B, seed=31)
    f32 = torch.from_numpy(np.stack([pre.preprocess(im.numpy(), cfg.image_size) for im in raw]))
    lab, box, nt = synth.make_targets(cfg, B, seed=32, max_t=8)
    scales = synth.make_class_scales(cfg).cuda()

    def drive(raw_u8, img):
        model, _ = _model(cfg)
        step = TrainStep(model, PushPullLoss(cfg.n_classes, scales), FusedAdamW(model, lr=1e-3, weight_decay=0.1),
                         batch=B, max_targets=lab.shape[1], n_input_slots=1, raw_u8=raw_u8)
        step.load(img.pin_memory(), lab.pin_memory(), box.pin_memory(), nt.pin_memory(), slot=0)
        torch.cuda.synchronize()
        step.warmup()
        out = [step.run(slot=0).clone() for _ in range(2)]
        torch.cuda.synchronize()
        return torch.stack(out).cpu(), model.flat_params.detach().cpu()

    l_u8, p_u8 = drive(True, raw)
    l_f32, p_f32 = drive(False, f32)
    assert torch.equal(l_u8[0], l_f32[0]), "same patch rows -> same first-step losses, bit for bit"
    np.testing.assert_allclose(l_u8.numpy(), l_f32.numpy(), rtol=2e-3)
    np.testing.assert_allclose(p_u8.numpy(), p_f32.numpy(), rtol=0, atol=2e-4)
