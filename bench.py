#!/usr/bin/env python
"""bench.py — images/sec of the OWL-ViT-B/32 768 px fine-tuning step + matcher us/image (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation of the path (the REAL
                                                           reference classes from oracle/_ref or /root/reference when
                                                           present, else the oracle port), fp32 torch on the host cores

One "step" = reference main.py:74-91 on one batch: zero_grad, forward, PushPullLoss (matcher + losses), backward under
the reference freeze policy, gradient all-reduce (N > 1), AdamW.  Workload = BASELINE.json configs[1]: OWL-ViT-B/32,
synthetic 768x768 images, batch 16 per GPU (weak scaling: global batch 16 N).

The ONE JSON line of our arm (rank 0) carries, besides the contract keys:
  roofline / roofline_fc1 / roofline_attention / roofline_out_proj / roofline_step   kernels timed live with CUDA events
  matcher              BASELINE.json configs[4]: cost + assignment us/image over >= 1 M synthetic images for T = 10 / 50 /
                       100, algorithmic GB/s against the measured HBM peak, index-exactness of >= 10 000 images per T
                       against the reference's arithmetic (oracle cost ops + the REAL scipy.linear_sum_assignment, on
                       the host cores), and the reference HungarianMatcher.forward timed on the same host
  torch_cuda_baseline  the reference's stock torch-CUDA step on this GPU (fp32; and bf16-autocast + SDPA): the
                       denominator of the north star's ">= 4x"
  cpu_baseline         the reference's CPU step on the host cores (bounded sample)
Keys are described in DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec OWL-ViT-B/32 768px fwd+bwd"
UNIT = "images/s"
BATCH_PER_GPU = 16
WORKLOAD = "OWL-ViT-B/32 fine-tune step (fwd + matcher/loss + bwd[reference freeze policy] + AdamW), synthetic 768x768"
LR, WD = 3e-6, 0.1           # reference config.yaml


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: 16 for b32, 4 for l14)")
    ap.add_argument("--workload", default="b32", choices=["b32", "l14"],
                    help="b32 = BASELINE.json configs[1] (the metric's configuration); l14 = configs[3], OWL-ViT-L/14 "
                         "840x840, batch 4 per GPU (an extension over the reference, SURVEY D5)")
    ap.add_argument("--input", default="u8", choices=["u8", "f32"],
                    help="u8: raw RGB bytes, normalised on the device (default for b32); f32: the reference's fp32 pixel_values")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-cuda-baseline", action="store_true")
    ap.add_argument("--no-matcher", action="store_true")
    ap.add_argument("--matcher-images", type=int, default=1_000_000, help="images per T in the matcher microbench")
    ap.add_argument("--matcher-check", type=int, default=10_000, help="images per T checked index-exact against SciPy")
    ap.add_argument("--matcher-ref-images", type=int, default=1000, help="images per T for the reference matcher timing")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_facts():
    """Per-launch DRAM traffic etc. read from the committed ncu summary (profiles/ncu_facts.json, written by
    tools/ncu_facts.py from `ncu --set full` captures of the same kernels at batch 16)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_facts.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "10", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)          # let the sampler come up before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(load), "sm_min_mhz": min(load), "sm_max_mhz": max(mx),
                "power_w_median": statistics.median(pw), "reasons": sorted(reasons), "samples": len(sm)}


# ====================================================================================== reference arm (CPU)
def reference_cpu_step_time(n_images: int, iters: int, warmup: int = 1):
    """The reference's own implementation of the path on the host cores, fp32, batch-1 loop exactly like reference
    main.py:70-93 (zero_grad, forward, PushPullLoss, backward, AdamW).  Uses the REAL reference classes when they are
    available (oracle/ref_arm.py: /root/reference or the byte-compiled copy under oracle/_ref), else the oracle port.
    Returns (s / image, threads, kind)."""
    import torch
    from oracle import ref_arm
    from owl_vit_object_detection_b200 import synth
    cfg = synth.B32
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_weights(cfg, seed=0)
    imgs = synth.make_images(cfg, n_images, seed=2)
    labels, tboxes, nt = synth.make_targets(cfg, n_images, seed=3)
    scales = synth.make_class_scales(cfg)
    if ref_arm.available():
        kind = "reference"
        model = ref_arm.build_model(cfg, sd)
        crit = ref_arm.criterion(cfg.n_classes, scales)
        opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD)          # reference main.py:56-60
        model.train()

        def one(b):
            t = int(nt[b])
            opt.zero_grad()
            boxes, _, sims, _ = model(imgs[b:b + 1])
            l = crit(sims, labels[b:b + 1, :t], boxes, tboxes[b:b + 1, :t])
            (l["loss_ce"] + l["loss_bg"] + l["loss_bbox"] + l["loss_giou"]).backward()
            opt.step()
    else:
        kind = "port"
        from oracle import matcher_oracle as mo
        from oracle import owlvit_oracle as oo
        train = synth.trainable_names(cfg)
        for n in train:
            sd[n].requires_grad_(True)
        opt = torch.optim.AdamW([sd[n] for n in train], lr=LR, weight_decay=WD)

        def one(b):
            t = int(nt[b])
            opt.zero_grad()
            boxes, sims = oo.forward(sd, cfg, imgs[b:b + 1])
            l, _, _ = mo.push_pull_loss(sims, boxes, [labels[b, :t]], [tboxes[b, :t]], cfg.n_classes, scales)
            sum(l.values()).backward()
            opt.step()
    times = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        for b in range(n_images):
            one(b)
        if it >= warmup:
            times.append((time.perf_counter() - t0) / n_images)
    return statistics.mean(times), torch.get_num_threads(), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = 2                                     # one "step" of this arm = the reference's batch-1 loop over 2 images
    steps = max(1, min(args.steps, 600))          # ~0.22 s per step on 16 host cores: K steps stay within minutes
    warm = max(1, min(args.warmup, 10))
    t0 = time.time()
    sec_per_img, cores, kind = reference_cpu_step_time(n_img, iters=steps, warmup=warm)
    val = 1.0 / sec_per_img
    sample = (f"{steps} timed steps of {n_img} images each (batch-1 loop as reference main.py:70-93: fwd + "
              f"PushPullLoss + bwd + AdamW), fp32 torch CPU, {cores} threads on {cpu_model()}; "
              f"{'the REAL reference classes (src/models.py, src/losses.py, src/matcher.py)' if kind == 'reference' else 'oracle port'}; "
              f"wall {time.time() - t0:.0f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec_per_img * 1e3 * n_img, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": n_img,
                   "note": "CPU arm does not use the GPUs; n_gpus echoes the launch; the reference is batch-1 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "cpu": cpu_model()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ====================================================================================== matcher microbench (configs[4])
def _matcher_ref_worker(job):
    """Host worker: the reference's assignment for a chunk of images = the reference's fp32 torch cost ops (restated in
    oracle.matcher_oracle.cost_matrix, pinned to the real HungarianMatcher by tests/golden/matcher_T*.npz) + the REAL
    scipy.optimize.linear_sum_assignment, exactly as reference src/matcher.py:135-137 calls it.  Returns match[n, T]
    (prediction index per target) and the oracle totals."""
    import numpy as np
    import torch
    from scipy.optimize import linear_sum_assignment
    from oracle import matcher_oracle as mo
    torch.set_num_threads(1)
    sims, pred, lab, tgt = (torch.from_numpy(x) for x in job)
    n, T = lab.shape
    match = np.full((n, T), -1, dtype=np.int32)
    totals = np.zeros(n, dtype=np.float64)
    for b in range(n):
        c = mo.cost_matrix(sims[b], pred[b], lab[b], tgt[b]).numpy()
        rows, cols = linear_sum_assignment(c)
        match[b, cols] = rows
        totals[b] = c[rows, cols].astype(np.float64).sum()
    return match, totals


def matcher_inputs(n, T, seed, dev, P=576, C=80):
    """SURVEY §8d "matcher bench": sims ~ U(-0.1, 0.3), predicted boxes drawn like the targets, fixed T."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    sims = torch.rand((n, P, C), generator=g, device=dev) * 0.4 - 0.1

    def boxes(k):
        cxy = 0.1 + 0.8 * torch.rand((n, k, 2), generator=g, device=dev)
        wh = 0.02 + 0.48 * torch.rand((n, k, 2), generator=g, device=dev)
        lo = (cxy - wh / 2).clamp(0.0, 1.0)
        hi = torch.maximum((cxy + wh / 2).clamp(0.0, 1.0), lo + 1e-3)
        return torch.cat([lo, hi], dim=-1).contiguous()
    return sims, boxes(P), torch.randint(0, C, (n, T), generator=g, device=dev), boxes(T)


def matcher_block(args, pool, n_workers, dev, pk):
    """BASELINE.json configs[4] on one GPU.  Device times are CUDA events around each kernel, per chunk of 16384 images
    (3 GB of inputs per chunk: far beyond L2).  The events are queued behind the chunk's generation kernels without a
    host sync in between, so host launch latency never sits inside an event interval; the host-side exactness check
    runs after ALL timing (its 15 busy worker processes would otherwise delay this process's launches)."""
    import numpy as np
    import torch
    from owl_vit_object_detection_b200 import ops
    from owl_vit_object_detection_b200.accounting import matcher_cost_bytes_per_image
    P, C, CH = 576, 80, 16384
    out = {"config": "576 predictions x T targets, 80 classes, cost matrix + assignment (reference src/matcher.py:103-137)",
           "images_per_T": 0, "chunk": CH, "per_T": {}}
    saved = {}
    for T in (10, 50, 100):
        n_chunks = max(1, (args.matcher_images + CH - 1) // CH)
        costT = torch.empty((CH, T, P), device=dev)
        match = torch.empty((CH, T), dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        nt = torch.full((CH,), T, dtype=torch.int32, device=dev)
        t_cost = t_lsap = 0.0
        for c in range(n_chunks):
            sims, pred, lab, tgt = matcher_inputs(CH, T, seed=1000 * T + c, dev=dev)
            if c == 0:   # warm-up
                ops.matcher_cost(sims, pred, lab, tgt, nt, costT, status)
                ops.lsap(costT, nt, match, status)
                torch.cuda.synchronize()
                sims, pred, lab, tgt = matcher_inputs(CH, T, seed=1000 * T + c, dev=dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            ops.matcher_cost(sims, pred, lab, tgt, nt, costT, status)
            ev[1].record()
            ops.lsap(costT, nt, match, status)
            ev[2].record()
            torch.cuda.synchronize()
            t_cost += ev[0].elapsed_time(ev[1])
            t_lsap += ev[1].elapsed_time(ev[2])
            if c == 0 and pool is not None and args.matcher_check > 0:
                k = min(CH, args.matcher_check)
                saved[T] = (sims[:k].cpu(), pred[:k].cpu(), lab[:k].cpu(), tgt[:k].cpu(), match[:k].cpu().numpy().copy(),
                            costT[:min(k, 256)].cpu())
            del sims, pred, lab, tgt
        n_done = n_chunks * CH
        assert status.item() == 0, "matcher status"
        bytes_img = matcher_cost_bytes_per_image(P, C, T)
        cost_gbs = bytes_img * n_done / (t_cost * 1e-3) / 1e9
        out["per_T"][str(T)] = {
            "images": n_done, "cost_us_per_image": t_cost * 1e3 / n_done, "lsap_us_per_image": t_lsap * 1e3 / n_done,
            "us_per_image": (t_cost + t_lsap) * 1e3 / n_done, "cost_bytes_per_image": bytes_img,
            "cost_achieved_gbs": cost_gbs, "cost_frac_of_hbm": cost_gbs / pk["hbm_gbs"]}
        out["images_per_T"] = n_done
        del costT, match
        torch.cuda.empty_cache()
    # ---- index-exactness against the reference's arithmetic (host processes; after all device timing)
    if saved:
        from oracle import matcher_oracle as mo
        jobs, spans = [], {}
        for T, (sc, pc, lc, tc, _, _) in saved.items():
            k = sc.shape[0]
            per = max(1, (k + 4 * n_workers - 1) // (4 * n_workers))
            first = len(jobs)
            # numpy slices: pickled by value through the pool's pipes (no dependence on the size of /dev/shm)
            jobs += [(sc[i:i + per].numpy(), pc[i:i + per].numpy(), lc[i:i + per].numpy(), tc[i:i + per].numpy())
                     for i in range(0, k, per)]
            spans[T] = (first, len(jobs))
        parts = pool.map(_matcher_ref_worker, jobs, chunksize=1)
        for T, (sc, pc, lc, tc, ours, cost_dev) in saved.items():
            ref = np.concatenate([p[0] for p in parts[spans[T][0]:spans[T][1]]])
            bad = np.nonzero((ref != ours).any(axis=1))[0]
            worst_gap = 0.0
            for b in bad:      # same optimum reached through a tie, or a real difference: compare totals on the oracle cost
                c = mo.cost_matrix(sc[b], pc[b], lc[b], tc[b]).numpy().astype(np.float64)
                worst_gap = max(worst_gap, abs(c[ours[b], np.arange(T)].sum() - c[ref[b], np.arange(T)].sum()))
            k = cost_dev.shape[0]      # cost-matrix agreement on the first 256 checked images
            refc = torch.stack([mo.cost_matrix(sc[b], pc[b], lc[b], tc[b]) for b in range(k)])   # [k,P,T]
            d = (cost_dev.transpose(1, 2) - refc).abs()
            out["per_T"][str(T)]["exactness"] = {
                "images_checked": int(ref.shape[0]), "mismatching_images": int(len(bad)),
                "max_total_cost_gap_of_mismatches": worst_gap,
                "cost_matrix_vs_oracle": {"entries": int(d.numel()), "bit_equal_frac": float((d == 0).float().mean()),
                                          "max_abs_diff": float(d.max())},
                "against": "oracle cost ops (= the reference's fp32 torch ops, pinned by tests/golden/matcher_T*.npz) "
                           "+ scipy.optimize.linear_sum_assignment, %d host processes" % n_workers}
    # ---- the reference's HungarianMatcher.forward on the host (single process, per image, as the reference runs it)
    out["reference_cpu"] = matcher_reference_cpu(args.matcher_ref_images)
    for T in (10, 50, 100):
        r = out["per_T"][str(T)]
        r["speedup_vs_reference_cpu"] = out["reference_cpu"]["us_per_image"][str(T)] / r["us_per_image"]
    return out


def matcher_reference_cpu(n_images: int):
    """reference src/matcher.py:85-159 `HungarianMatcher.forward`, per image, single process, host cores."""
    import torch
    from oracle import ref_arm
    from owl_vit_object_detection_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    real = ref_arm.available()
    res = {}
    for T in (10, 50, 100):
        n = n_images
        sims, pred, lab, tgt = synth.make_matcher_inputs(min(n, 256), T, seed=4)
        m = sims.shape[0]
        if real:
            matcher = ref_arm.matcher(80)

            def one(b):
                matcher({"pred_logits": sims[b:b + 1], "pred_boxes": pred[b:b + 1]}, [{"labels": lab[b], "boxes": tgt[b]}])
        else:
            from oracle import matcher_oracle as mo

            def one(b):
                mo.hungarian(sims[b:b + 1], pred[b:b + 1], [lab[b]], [tgt[b]], 80)
        for b in range(5):
            one(b % m)
        t0 = time.perf_counter()
        for b in range(n):
            one(b % m)
        res[str(T)] = (time.perf_counter() - t0) / n * 1e6
    return {"us_per_image": res, "images_per_T": n_images, "kind": "reference" if real else "port",
            "threads": torch.get_num_threads(), "cpu": cpu_model(),
            "what": "HungarianMatcher.forward per image (cost ops + .cpu() + scipy LSAP + index build), single process"}


# ====================================================================================== stock torch-CUDA baseline
def torch_cuda_baseline(B: int, dev):
    """The reference's stock torch-CUDA path on this GPU: the REAL reference classes (`OwlViT(...).to("cuda")`,
    `PushPullLoss`, `torch.optim.AdamW(model.parameters())`) driven as reference main.py:74-91 drives them - fp32, torch
    defaults (matmul TF32 off) - plus the same model under bf16 autocast + SDPA (BASELINE.md §3 "strong baseline").  The
    reference's loss is batch-1 only: the batch-16 columns run the model on the batch and loop the loss per
    image; `batch1_literal` is main.py's own batch-1 loop.  Falls back to the oracle port when the real classes are not
    available (kind says which).  Reported beside our number; never part of the product path."""
    import torch
    from oracle import ref_arm
    from owl_vit_object_detection_b200 import synth
    cfg = synth.B32
    sd = synth.make_weights(cfg, seed=0)
    imgs = synth.make_images(cfg, B, seed=2).to(dev)
    labels, tboxes, nt = synth.make_targets(cfg, B, seed=3)
    labels_d, tboxes_d = labels.to(dev), tboxes.to(dev)
    scales = synth.make_class_scales(cfg)
    out = {"batch": B}
    if not ref_arm.available():
        out["kind"] = "port"
        out.update(_torch_cuda_port(B, cfg, sd, imgs, labels, tboxes, nt, scales))
        return out
    out["kind"] = "reference"
    model = ref_arm.build_model(cfg, sd, attn_implementation="sdpa").to(dev)      # load_model(...).to(device)
    crit = ref_arm.criterion(cfg.n_classes, scales.to(dev))
    opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD)
    model.train()
    hf_cfgs = [m.config for m in model.modules() if hasattr(m, "config") and hasattr(m.config, "_attn_implementation")]

    def set_attn(impl):
        for c in hf_cfgs:
            c._attn_implementation = impl

    def step(batch_imgs, idx, autocast, with_loss):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            boxes, _, sims, _ = model(batch_imgs)
        boxes, sims = boxes.float(), sims.float()
        if with_loss:
            total = 0.0
            for j, b in enumerate(idx):       # reference loss: one image per call (src/losses.py:23-24 squeeze_(0))
                t = int(nt[b])
                l = crit(sims[j:j + 1].clone(), labels_d[b:b + 1, :t], boxes[j:j + 1].clone(), tboxes_d[b:b + 1, :t])
                total = total + (l["loss_ce"] + l["loss_bg"] + l["loss_bbox"] + l["loss_giou"])
            (total / len(idx)).backward()
        else:
            (sims.sum() + boxes.sum()).backward()
        opt.step()

    def timed(fn, n_img, iters=3):
        ts = []
        for it in range(iters + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return n_img / min(ts[1:])

    all_idx = list(range(B))
    for name, attn, autocast in (("fp32_sdpa", "sdpa", False), ("fp32_eager", "eager", False), ("bf16_autocast_sdpa", "sdpa", True)):
        set_attn(attn)
        out[name] = {
            "full_step_images_per_s": timed(lambda: step(imgs, all_idx, autocast, True), B),
            "fwd_bwd_adamw_only_images_per_s": timed(lambda: step(imgs, all_idx, autocast, False), B),
        }
    set_attn("eager")

    def literal():
        for b in range(4):
            step(imgs[b:b + 1], [b], False, True)
    out["fp32_eager"]["batch1_literal_images_per_s"] = timed(literal, 4, iters=2)
    out["stock"] = "fp32_eager"
    out["note"] = ("the REAL reference classes on cuda, fp32 (matmul TF32 off = torch default); 'fp32_eager' is the attention the "
                   "reference's pinned transformers 4.30.2 runs (bmm/softmax/bmm) = the stock path; 'fp32_sdpa' is what the installed "
                   "transformers 5.5.0 picks by default; 'bf16_autocast_sdpa' is the strong-baseline column (forward under autocast, "
                   "reference loss in fp32 per image); fwd_bwd_adamw_only replaces the loss by a sum (upper bound for any loss)")
    return out


def _torch_cuda_port(B, cfg, sd, imgs, labels, tboxes, nt, scales):
    import torch
    from oracle import matcher_oracle as mo
    from oracle import owlvit_oracle as oo
    from owl_vit_object_detection_b200 import synth
    dev = imgs.device
    sd = {k: v.to(dev) for k, v in sd.items()}
    train = synth.trainable_names(cfg)
    for n in train:
        sd[n].requires_grad_(True)
    opt = torch.optim.AdamW([sd[n] for n in train], lr=LR, weight_decay=WD)
    orig_box_bias = oo.box_bias
    oo.box_bias = lambda c: orig_box_bias(c).to(dev)
    out = {}
    try:
        for name, with_loss in (("fwd_bwd_adamw_only_images_per_s", False), ("full_step_images_per_s", True)):
            ts = []
            for it in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                opt.zero_grad()
                boxes, sims = oo.forward(sd, cfg, imgs)
                if with_loss:
                    lab_l = [labels[b, :nt[b]] for b in range(B)]
                    box_l = [tboxes[b, :nt[b]] for b in range(B)]
                    l, _, _ = mo.push_pull_loss(sims.cpu(), boxes.cpu(), lab_l, box_l, cfg.n_classes, scales)
                    sum(l.values()).backward()
                else:
                    (sims.sum() + boxes.sum()).backward()
                opt.step()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            out[name] = B / min(ts[1:])
    finally:
        oo.box_bias = orig_box_bias
    return {"fp32_eager": out, "stock": "fp32_eager", "note": "oracle port on cuda, fp32 (the real reference classes were not available)"}


# ====================================================================================== our arm
def time_layer_kernels(eng, ws, cfg, B, reps=4):
    """Average device time (ms) of each kernel of an encoder layer (HF:490-511) under the conditions it meets inside the
    step: the twelve layers are swept in order with their own weights, exactly as Engine.forward launches them, so
    every launch finds activations as L2-warm as its producer left them and weights as cold as in the step; CUDA
    events on the launching stream around every single launch (the queue stays full, so an interval is kernel
    duration + the ~1 us hand-over, never host latency).  The first sweep is a warm-up."""
    import torch
    from owl_vit_object_detection_b200 import ops
    S, D, F, H, dh, eps = cfg.tokens, cfg.hidden, cfg.ff, cfg.heads, cfg.head_dim, cfg.ln_eps
    M = B * S
    L = eng.layout
    names = ("ln1", "qkv", "attn", "out_proj", "ln2", "fc1", "fc2")
    acc = {n: [] for n in names}
    x0 = ws.x.clone()          # a real residual-stream state (input of the last layer of the last step)
    for rep in range(reps + 1):
        ws.x.copy_(x0)
        torch.cuda._sleep(8_000_000)      # ~4 ms spin kernel: the host enqueues the whole sweep behind it
        evs = []
        for i in range(cfg.layers):
            p = f"backbone.encoder.layers.{i}."
            lo, hi = L.span(p + "self_attn.q_proj.weight", p + "self_attn.v_proj.weight")
            wqkv = eng.flat16[lo:hi].view(3 * D, D)
            lo, hi = L.span(p + "self_attn.q_proj.bias", p + "self_attn.v_proj.bias")
            bqkv = eng.flat32[lo:hi]
            seq = (
                ("ln1", lambda: ops.layernorm(ws.x, eng.p32(p + "layer_norm1.weight"), eng.p32(p + "layer_norm1.bias"), ws.h1, rows=M, D=D, eps=eps)),
                ("qkv", lambda: ops.gemm(ws.h1, wqkv, ws.qkv, M=M, N=3 * D, K=D, bias=bqkv)),
                ("attn", lambda: ops.flash_attn_fwd(ws.qkv, ws.ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)),
                ("out_proj", lambda: ops.gemm(ws.ctx, eng.p16(p + "self_attn.out_proj.weight"), ws.x, M=M, N=D, K=D,
                                              bias=eng.p32(p + "self_attn.out_proj.bias"), resid=ws.x)),
                ("ln2", lambda: ops.layernorm(ws.x, eng.p32(p + "layer_norm2.weight"), eng.p32(p + "layer_norm2.bias"), ws.h2, rows=M, D=D, eps=eps)),
                ("fc1", lambda: ops.gemm(ws.h2, eng.p16(p + "mlp.fc1.weight"), ws.m, M=M, N=F, K=D, bias=eng.p32(p + "mlp.fc1.bias"), act="quick_gelu")),
                ("fc2", lambda: ops.gemm(ws.m, eng.p16(p + "mlp.fc2.weight"), ws.x, M=M, N=D, K=F, bias=eng.p32(p + "mlp.fc2.bias"), resid=ws.x)),
            )
            for name, fn in seq:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                evs.append((name, e0, e1))
        torch.cuda.synchronize()
        if rep > 0:
            for name, e0, e1 in evs:
                acc[name].append(e0.elapsed_time(e1))
    return {n: sum(v) / len(v) for n, v in acc.items()}


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    do_matcher = (not args.no_matcher) and world == 1 and args.workload == "b32"
    pool, n_workers = None, 0
    if do_matcher and args.matcher_check > 0:
        # host workers for the exactness check: forked BEFORE this process touches CUDA
        import multiprocessing as mp
        n_workers = max(1, (os.cpu_count() or 2) - 1)
        pool = mp.get_context("fork").Pool(n_workers)

    import torch
    import torch.distributed as dist
    from owl_vit_object_detection_b200 import _lib, ops, synth
    from owl_vit_object_detection_b200.accounting import flops_per_image
    from owl_vit_object_detection_b200.loss import PushPullLoss
    from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT
    from owl_vit_object_detection_b200.train import TrainStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at communicator creation: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    cfg = synth.B32 if args.workload == "b32" else synth.L14
    B = args.batch or (BATCH_PER_GPU if args.workload == "b32" else 4)
    raw_u8 = args.input == "u8" and cfg.patch_size % 8 == 0
    workload = WORKLOAD if args.workload == "b32" else WORKLOAD.replace("B/32", "L/14").replace("768x768", "840x840")
    metric = METRIC if args.workload == "b32" else METRIC.replace("B/32 768px", "L/14 840px")
    sd = synth.make_weights(cfg, seed=0)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device=dev)
    del sd
    crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).to(dev))
    opt = FusedAdamW(model, lr=LR, weight_decay=WD)
    # rotating input batches: together with > 1 GB of activations per step they exceed the 126 MB L2
    n_slots = 6 if raw_u8 else 3
    step = TrainStep(model, crit, opt, batch=B, n_input_slots=n_slots, raw_u8=raw_u8)

    # synthetic COCO-shaped data, different per rank and per slot
    host = []
    for s in range(n_slots):
        if raw_u8:
            img = synth.make_images_u8(cfg, B, seed=100 + rank * 16 + s).pin_memory()
        else:
            img = synth.make_images(cfg, B, seed=100 + rank * 16 + s).pin_memory()
        lab, box, nt = synth.make_targets(cfg, B, seed=200 + rank * 16 + s)
        host.append((img, lab.pin_memory(), box.pin_memory(), nt.pin_memory()))
    for s in range(n_slots):
        step.load(*host[s], slot=s)
    torch.cuda.synchronize()
    step.warmup()                                               # graph capture
    torch.cuda.synchronize()
    l0 = _lib.KERNEL_LAUNCHES
    step.use_graph = False
    step.run(slot=0)                                            # one eager step just to count kernels per step
    step.use_graph = True
    torch.cuda.synchronize()
    launches_per_step = _lib.KERNEL_LAUNCHES - l0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, graph replay, CUDA events, max over ranks
    warm = max(3, args.warmup)
    for _ in range(warm):
        step.run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        losses = step.run()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)
    final_losses = losses.tolist()

    # ---------------- e2e: the reference-facing call sequence with HOST buffers: every step copies its batch
    # from pinned host memory (prefetched one step ahead on a copy stream) and reads the 4 losses back.
    h2d = sum(x.numel() * x.element_size() for x in host[0])
    d2h = 16
    for i in range(2):
        step.run(slot=step.load(*host[i % n_slots])).tolist()
    barrier()
    e0.record()
    nxt = step.load(*host[0])
    pending = None
    for i in range(args.steps):
        cur = nxt
        if i + 1 < args.steps:
            nxt = step.load(*host[(i + 1) % n_slots])           # overlaps with this step's compute
        step.run(slot=cur, readback=True)                       # queues the 16-byte D2H copy of the step's losses
        if pending is not None:
            _ = step.result(pending)                            # host reads step i-1's losses while step i runs
        pending = cur
    _ = step.result(pending)                                    # every step's result is read inside the timed region
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (t.item() * 1e-3)
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions (value and e2e)
    # what the e2e number is bounded by: the host->device rate of one batch of images from pinned memory
    e0.record()
    for _ in range(3):
        step.slots[0]["image"].copy_(host[0][0], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * host[0][0].numel() * host[0][0].element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # ---------------- rooflines of the dominant kernels, timed live (CUDA events per launch, L2 flushed between launches)
    pk, pk_kind = peaks()
    ncu = ncu_facts()
    use_ncu = B == BATCH_PER_GPU and args.workload == "b32"
    M = B * cfg.tokens
    D, F = cfg.hidden, cfg.ff
    eng = model.engine
    ws = eng.workspace(B)
    p = f"backbone.encoder.layers.{cfg.layers - 1}."
    kt = time_layer_kernels(eng, ws, cfg, B)
    layer_ms = sum(kt.values())

    def roof(name, key_t, flops, key_ncu):
        k_ms = kt[key_t]
        ach = flops / (k_ms * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops"], "traffic": ncu.get(key_ncu + "_dram_bytes") if use_ncu else None,
                "tensor_pipe_active_pct_ncu": ncu.get(key_ncu + "_tensor_pipe_pct") if use_ncu else None,
                "traffic_source": ncu.get("source"), "peak_source": pk_kind + " (burst figure, of measured)",
                "launch_us": k_ms * 1e3, "flops_per_launch": flops, "share_of_layer": k_ms / layer_ms,
                "timing": "CUDA events around every launch of a 12-layer sweep with each layer's own weights (step-like cache "
                          "state; mean over %d launches)" % (4 * cfg.layers)}

    # fc2 is the launch with the largest share of the step (profiles/*_step_launches.txt)
    roofline = roof("gemm_tc_kernel (MLP fc2 + bias + residual, fp32 out, M=%d N=%d K=%d)" % (M, D, F), "fc2", 2.0 * M * D * F, "fc2_gemm")
    roofline_fc1 = roof("gemm_tc_kernel (MLP fc1 + bias + quick_gelu, fp16 out, M=%d N=%d K=%d)" % (M, F, D), "fc1", 2.0 * M * F * D, "fc1_gemm")
    roofline_qkv = roof("gemm_tc_kernel (fused q|k|v projection + bias, fp16 out, M=%d N=%d K=%d)" % (M, 3 * D, D), "qkv", 2.0 * M * 3 * D * D, "qkv_gemm")
    roofline_out = roof("gemm_tc_kernel (attention out-proj + bias + residual, fp32 out, M=%d N=%d K=%d)" % (M, D, D), "out_proj", 2.0 * M * D * D, "out_proj_gemm")
    roofline_attn = roof("flash_attn_fwd kernel (S=%d, H=%d, dh=%d)" % (cfg.tokens, cfg.heads, cfg.head_dim), "attn",
                         4.0 * B * cfg.heads * cfg.tokens * cfg.tokens * cfg.head_dim, "flash_attn")
    roofline_attn["note"] = "head_dim 64: exp2 on the MUFU needs 2x the MMA cycles per score tile, see DESIGN.md"
    ln_bytes = M * D * (4 + 2)
    roofline_ln = {"bound": "hbm", "kernel": "layernorm_kernel (fp32 in, fp16 out, %d x %d)" % (M, D),
                   "achieved": ln_bytes / (kt["ln1"] * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                   "frac": ln_bytes / (kt["ln1"] * 1e-3) / 1e9 / pk["hbm_gbs"], "launch_us": kt["ln1"] * 1e3,
                   "bytes_per_launch": ln_bytes, "share_of_layer": (kt["ln1"] + kt["ln2"]) / layer_ms,
                   "note": "input just written by the previous GEMM: partly served from L2, so the fraction can exceed 1"}
    fl = flops_per_image(cfg)
    step_tflops = fl["fwd_bwd_ref_policy"] * B / (ms_per_step * 1e-3) / 1e12
    step_roof = {"flops_per_image": fl["fwd_bwd_ref_policy"], "achieved_tflops_per_gpu": step_tflops,
                 "peak": pk["bf16_tflops_sustained"], "frac": step_tflops / pk["bf16_tflops_sustained"],
                 "attention_gemm_flops_per_image": fl["attn_core"] * cfg.layers, "peak_source": pk_kind + " (sustained)"}

    line = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}" + (" (all-reduce of one flat fp32 grad buffer)" if world > 1 else ""),
                   "precision": "fp16 operands, fp32 accumulate / residual / master weights",
                   "input": ("raw RGB uint8 [B,768,768,3]; rescale + CLIP normalise (reference src/dataset.py:64-71) fused into "
                             "the patch gather on the device" if raw_u8 else "fp32 pixel_values [B,3,H,W] as the reference's DataLoader yields"),
                   "l2": f"{n_slots} rotating input batches ({n_slots * h2d >> 20} MB) + >1 GB of activations per step exceed the 126 MB L2",
                   "launch": step.launch_description(),
                   "e2e": "per step: H2D of the batch from pinned host memory (prefetched one step ahead on a copy stream) + "
                          "D2H of the 4 losses, read on the host while the next step runs"},
        "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_gbs_measured": h2d_gbs, "h2d_ms_per_step_at_that_rate": h2d / h2d_gbs * 1e-6},
        "roofline": roofline, "roofline_fc1": roofline_fc1, "roofline_qkv": roofline_qkv, "roofline_out_proj": roofline_out,
        "roofline_attention": roofline_attn, "roofline_layernorm": roofline_ln, "roofline_step": step_roof,
        "encoder_layer_us": {k: v * 1e3 for k, v in kt.items()}, "final_losses": final_losses,
        "host": {"cpu": cpu_model(), "cores": os.cpu_count()},
    }

    if rank == 0 and world == 1:
        del step, model, crit, opt, ws, eng
        torch.cuda.empty_cache()
        if do_matcher:
            line["matcher"] = matcher_block(args, pool, n_workers, dev, pk)
        if pool is not None:
            pool.close()
            pool.join()
            pool = None
        if not args.no_torch_cuda_baseline and args.workload == "b32":
            try:
                line["torch_cuda_baseline"] = torch_cuda_baseline(B, dev)
                stock = line["torch_cuda_baseline"][line["torch_cuda_baseline"]["stock"]]["full_step_images_per_s"]
                line["torch_cuda_baseline"]["ours_over_stock"] = value / stock
            except Exception as e:      # a baseline leg must never take the measurement down
                line["torch_cuda_baseline"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline and args.workload == "b32":
            sec, cores, kind = reference_cpu_step_time(2, iters=3, warmup=1)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": kind, "cpu": cpu_model(),
                                    "sample": "3 timed passes over 2 images (batch-1 loop as reference main.py:70-93: fwd + "
                                              "PushPullLoss + bwd + AdamW), fp32 torch CPU, "
                                              + ("the REAL reference classes" if kind == "reference" else "oracle port")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # the step graphs hold captured NCCL kernels: tearing the communicator down underneath them can block at
        # interpreter exit, so every rank synchronises and leaves without running destructors
        sys.stdout.flush()
        sys.stderr.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
