#!/usr/bin/env python
"""bench.py — images/sec of the OWL-ViT-B/32 768 px fine-tuning step (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU implementation of the path
                                                           (oracle port, fp32 torch on the host cores)

One "step" = reference main.py:74-91 on one batch: zero_grad, forward, PushPullLoss (matcher + losses),
backward under the reference freeze policy, gradient all-reduce (N > 1), AdamW.  Workload = BASELINE.json
configs[1]: OWL-ViT-B/32, synthetic 768x768 images, batch 16 per GPU (weak scaling: global batch 16 N).

Prints ONE JSON line (rank 0).  Keys are described in DESIGN.md §Measurement.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: 16 for b32, 4 for l14)")
    ap.add_argument("--workload", default="b32", choices=["b32", "l14"],
                    help="b32 = BASELINE.json configs[1] (the metric's configuration); l14 = configs[3], OWL-ViT-L/14 "
                         "840x840, batch 4 per GPU (an extension over the reference, SURVEY D5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-cuda-baseline", action="store_true",
                    help="also time the fp32 torch-CUDA restatement of the reference step (stock-path denominator)")
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
    tools/ncu_facts.py from an `ncu --set full` capture of the same kernels at batch 16)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_facts.json")) as f:
            return json.load(f)
    except Exception:
        return {}


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
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ====================================================================================== reference arm (CPU)
def oracle_cpu_step_time(n_images: int, iters: int, warmup: int = 1):
    """The reference's own implementation of the path, restated (oracle/): fp32 torch on the host cores, batch-1
    loop exactly like reference main.py:70-93 (forward, PushPullLoss, backward, AdamW).  Returns s / image."""
    import torch
    from oracle import matcher_oracle as mo
    from oracle import owlvit_oracle as oo
    from owl_vit_object_detection_b200 import synth
    cfg = synth.B32
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_weights(cfg, seed=0)
    train = synth.trainable_names(cfg)
    for n in train:
        sd[n].requires_grad_(True)
    opt = torch.optim.AdamW([sd[n] for n in train], lr=3e-6, weight_decay=0.1)
    imgs = synth.make_images(cfg, n_images, seed=2)
    labels, tboxes, nt = synth.make_targets(cfg, n_images, seed=3)
    scales = synth.make_class_scales(cfg)
    times = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        for b in range(n_images):
            opt.zero_grad()
            boxes, sims = oo.forward(sd, cfg, imgs[b:b + 1])
            t = int(nt[b])
            l, _, _ = mo.push_pull_loss(sims, boxes, [labels[b, :t]], [tboxes[b, :t]], cfg.n_classes, scales)
            sum(l.values()).backward()
            opt.step()
        if it >= warmup:
            times.append((time.perf_counter() - t0) / n_images)
    return statistics.mean(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = 2
    t0 = time.time()
    sec_per_img, cores = oracle_cpu_step_time(n_img, iters=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    val = 1.0 / sec_per_img
    sample = (f"{args.steps} timed steps of {n_img} images each (batch-1 loop as reference main.py:70-93: fwd + "
              f"PushPullLoss + bwd + AdamW), fp32 torch CPU, {cores} threads; wall {time.time() - t0:.0f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_img * 1e3 * n_img, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": n_img, "note": "CPU arm does not use the GPUs; n_gpus echoes the launch"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ====================================================================================== our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from owl_vit_object_detection_b200 import _lib, ops, synth
    from owl_vit_object_detection_b200.loss import PushPullLoss
    from owl_vit_object_detection_b200.model import FusedAdamW, OwlViT
    from owl_vit_object_detection_b200.train import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    workload = WORKLOAD if args.workload == "b32" else WORKLOAD.replace("B/32", "L/14").replace("768x768", "840x840")
    metric = METRIC if args.workload == "b32" else METRIC.replace("B/32 768px", "L/14 840px")
    sd = synth.make_weights(cfg, seed=0)
    model = OwlViT({k: v for k, v in sd.items() if k != "queries"}, sd["queries"], cfg=cfg, device=dev)
    del sd
    crit = PushPullLoss(cfg.n_classes, synth.make_class_scales(cfg).to(dev))
    opt = FusedAdamW(model, lr=3e-6, weight_decay=0.1)          # reference config.yaml: lr 3e-6, wd 0.1
    n_slots = 3                                                 # 3 x 113 MB of fp32 images > 126 MB of L2
    step = TrainStep(model, crit, opt, batch=B, n_input_slots=n_slots)

    # synthetic COCO-shaped data, different per rank and per slot
    host = []
    for s in range(n_slots):
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
    for _ in range(max(3, args.warmup)):
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
    # what the e2e number is bounded by: the host->device rate of one batch of fp32 images from pinned memory
    e0.record()
    for _ in range(3):
        step.slots[0]["image"].copy_(host[0][0], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * host[0][0].numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # ---------------- roofline of the dominant kernel (the MLP fc1 tcgen05 GEMM, bias + quick_gelu epilogue),
    # timed live with CUDA events on the launching stream
    pk, pk_kind = peaks()
    M = B * cfg.tokens
    ws = model.engine.workspace(B)
    p = f"backbone.encoder.layers.{cfg.layers - 1}."
    w1, b1 = model.engine.p16(p + "mlp.fc1.weight"), model.engine.p32(p + "mlp.fc1.bias")

    def fc1():
        ops.gemm(ws.h2, w1, ws.m, M=M, N=cfg.ff, K=cfg.hidden, bias=b1, act="quick_gelu")
    for _ in range(3):
        fc1()
    torch.cuda.synchronize()
    reps = 20
    e0.record()
    for _ in range(reps):
        fc1()
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    k_flops = 2.0 * M * cfg.ff * cfg.hidden
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    ncu = ncu_facts()
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel<256,K,K,EpiF16<quick_gelu>> (MLP fc1, M=%d N=%d K=%d)" % (M, cfg.ff, cfg.hidden),
                "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"],
                "traffic": ncu.get("fc1_gemm_dram_bytes") if (B == BATCH_PER_GPU and args.workload == "b32") else None,
                "traffic_source": ncu.get("source"), "peak_source": pk_kind + " (burst: kernel timed alone)",
                "launch_us": k_ms * 1e3, "flops_per_launch": k_flops}
    # the fused attention kernel (north star: fraction of the attention-GEMM roofline), timed the same way
    def fa():
        ops.flash_attn_fwd(ws.qkv, ws.ctx, B=B, S=cfg.tokens, H=cfg.heads, head_dim=cfg.head_dim, scale=cfg.head_dim ** -0.5)
    for _ in range(3):
        fa()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fa()
    e1.record()
    torch.cuda.synchronize()
    a_ms = e0.elapsed_time(e1) / reps
    a_flops = 4.0 * B * cfg.heads * cfg.tokens * cfg.tokens * cfg.head_dim
    roofline_attn = {"bound": "tensor", "kernel": "flash_attn_fwd2_kernel (S=%d, H=%d, dh=%d)" % (cfg.tokens, cfg.heads, cfg.head_dim),
                     "achieved": a_flops / (a_ms * 1e-3) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": a_flops / (a_ms * 1e-3) / 1e12 / pk["bf16_tflops"], "launch_us": a_ms * 1e3,
                     "flops_per_launch": a_flops, "traffic": ncu.get("flash_attn_dram_bytes") if (B == BATCH_PER_GPU and args.workload == "b32") else None,
                     "tensor_pipe_active_pct_ncu": ncu.get("flash_attn_tensor_pipe_pct"),
                     "note": "head_dim 64: the MUFU (exp2) and the TMEM read of S each need 2x the MMA cycles, see DESIGN.md"}
    from oracle.owlvit_oracle import flops_per_image  # FLOP accounting only
    fl = flops_per_image(cfg)
    step_tflops = fl["fwd_bwd_ref_policy"] * B / (ms_per_step * 1e-3) / 1e12
    step_roof = {"flops_per_image": fl["fwd_bwd_ref_policy"], "achieved_tflops_per_gpu": step_tflops,
                 "peak": pk["bf16_tflops_sustained"], "frac": step_tflops / pk["bf16_tflops_sustained"],
                 "attention_gemm_flops_per_image": fl["attn_core"] * cfg.layers}

    line = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}" + (" (NCCL all-reduce of one flat fp32 grad buffer)" if world > 1 else ""),
                   "precision": "fp16 operands, fp32 accumulate / residual / master weights",
                   "l2": f"{n_slots} rotating input batches ({n_slots * h2d >> 20} MB) + >1 GB of activations per step exceed the 126 MB L2",
                   "launch": ("one CUDA-graph replay per step (fwd + loss + bwd + AdamW)" if world == 1 else
                              "two CUDA-graph replays per step (fwd+loss+bwd, AdamW) around the NCCL all-reduce"),
                   "e2e": "per step: H2D of the batch from pinned memory (prefetched one step ahead on a copy stream) + "
                          "D2H of the 4 losses, read on the host while the next step runs"},
        "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_gbs_measured": h2d_gbs, "h2d_ms_per_step_at_that_rate": h2d / h2d_gbs * 1e-6},
        "roofline": roofline, "roofline_attention": roofline_attn, "roofline_step": step_roof, "final_losses": final_losses,
    }

    if rank == 0 and world == 1 and args.torch_cuda_baseline:
        line["torch_cuda_baseline"] = torch_cuda_baseline(B)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "b32":
        del step, model
        torch.cuda.empty_cache()
        sec, cores = oracle_cpu_step_time(2, iters=3, warmup=1)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "3 timed passes over 2 images (batch-1 loop as reference main.py:70-93: fwd + "
                                          "PushPullLoss + bwd + AdamW), oracle fp32 torch CPU"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def torch_cuda_baseline(B: int):
    """The reference's stock torch-CUDA path, restated: fp32 torch ops on the GPU (cuBLAS/ATen), loss looped per image
    with its host syncs, torch AdamW.  Reported beside our number; not part of the product path."""
    import torch
    from oracle import matcher_oracle as mo
    from oracle import owlvit_oracle as oo
    from owl_vit_object_detection_b200 import synth
    cfg = synth.B32
    dev = "cuda"
    sd = {k: v.to(dev) for k, v in synth.make_weights(cfg, seed=0).items()}
    train = synth.trainable_names(cfg)
    for n in train:
        sd[n].requires_grad_(True)
    opt = torch.optim.AdamW([sd[n] for n in train], lr=3e-6, weight_decay=0.1)
    imgs = synth.make_images(cfg, B, seed=2).to(dev)
    labels, tboxes, nt = synth.make_targets(cfg, B, seed=3)
    scales = synth.make_class_scales(cfg)
    orig_box_bias = oo.box_bias
    oo.box_bias = lambda c: orig_box_bias(c).to(dev)
    out = {}
    try:
        for name, with_loss in (("fwd_bwd_only", False), ("full_step", True)):
            ts = []
            for it in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                opt.zero_grad()
                boxes, sims = oo.forward(sd, cfg, imgs)
                if with_loss:
                    # the reference's loss runs on the host per image (src/matcher.py:132 `.cpu()`, SciPy)
                    lab_l = [labels[b, :nt[b]] for b in range(B)]
                    box_l = [tboxes[b, :nt[b]] for b in range(B)]
                    sc, bc = sims.cpu(), boxes.cpu()
                    l, _, _ = mo.push_pull_loss(sc, bc, lab_l, box_l, cfg.n_classes, scales)
                    sum(l.values()).backward()
                else:
                    (sims.sum() + boxes.sum()).backward()
                opt.step()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            out[name + "_images_per_s"] = B / min(ts[1:])
    finally:
        oo.box_bias = orig_box_bias
    out["note"] = "fp32 torch ops (TF32 off for matmul, torch default), batch %d" % B
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
