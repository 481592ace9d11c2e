"""`TrainStep` — one fine-tuning step (reference main.py:74-91: zero_grad, forward, PushPullLoss, backward,
optimizer step) as a replayable CUDA graph, with the data-parallel gradient all-reduce between the backward
graph and the optimizer graph.

  step.load(images_host, labels, boxes, num_targets)   async H2D into static device buffers (side stream)
  losses4 = step.run()                                  device tensor [loss_ce, loss_bg, loss_bbox, loss_giou]

CUDA streams and graphs replace a tracing compiler: the ~190 kernels of a step are launched by ONE graph replay
on a single GPU (forward + loss + backward + AdamW), and by two replays around the NCCL all-reduce call (one flat
fp32 buffer) under data parallelism.  `run(readback=True)` also queues the 16-byte copy of the four
losses into pinned host memory; `result(slot)` waits for that copy only, so the host can read step i while step
i + 1 is already running.
"""
from __future__ import annotations

from typing import Optional

import torch

from .loss import PushPullLoss
from .model import FusedAdamW, OwlViT


class TrainStep:
    def __init__(self, model: OwlViT, criterion: PushPullLoss, optimizer: FusedAdamW, batch: int,
                 max_targets: int = 100, use_graph: bool = True, n_input_slots: int = 2, group=None,
                 raw_u8: bool = False, world: Optional[int] = None):
        """raw_u8: the input slots hold raw RGB bytes [B,IS,IS,3] (what a decoder produces; the reference's CPU
        rescale + normalise, src/dataset.py:64-71, then happens inside the patch gather on the device) instead of
        the reference's fp32 `pixel_values` [B,3,IS,IS]: a quarter of the host-to-device bytes per step."""
        cfg = model.cfg
        dev = model.flat_params.device
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.batch, self.group, self.use_graph = batch, group, use_graph
        self.slots = []
        for _ in range(n_input_slots):
            self.slots.append(dict(
                image=(torch.zeros((batch, cfg.image_size, cfg.image_size, 3), dtype=torch.uint8, device=dev) if raw_u8
                       else torch.zeros((batch, 3, cfg.image_size, cfg.image_size), dtype=torch.float32, device=dev)),
                labels=torch.full((batch, max_targets), -1, dtype=torch.int64, device=dev),
                boxes=torch.zeros((batch, max_targets, 4), dtype=torch.float32, device=dev),
                nt=torch.zeros((batch,), dtype=torch.int32, device=dev)))   # no targets until load()
        self.losses = [torch.zeros(4, dtype=torch.float32, device=dev) for _ in range(n_input_slots)]
        self._ones4 = torch.ones(4, dtype=torch.float32, device=dev)
        self._dsims = torch.zeros((batch, cfg.patches, cfg.n_classes), dtype=torch.float32, device=dev)
        self._dboxes = torch.zeros((batch, cfg.patches, 4), dtype=torch.float32, device=dev)
        self.host_losses = [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in range(n_input_slots)]
        self.read_done = [torch.cuda.Event() for _ in range(n_input_slots)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.loaded = [torch.cuda.Event() for _ in range(n_input_slots)]
        self.consumed = [torch.cuda.Event() for _ in range(n_input_slots)]
        self._fwdbwd = [None] * n_input_slots
        self._opt_graph = None
        self._world = 1
        import torch.distributed as dist
        if world is not None:
            self._world = world         # tests: a single-process reference inside a distributed run
        elif dist.is_available() and dist.is_initialized():
            self._world = dist.get_world_size(group)
        optimizer.grad_mul = 1.0 / self._world
        self.comm_stream = torch.cuda.Stream(device=dev)
        import os as _os
        # N > 1: reduce the gradients inside the NVSwitch when the platform has multicast support (OWL_DP_MULTIMEM=0
        # keeps NCCL); must happen before the first backward so that every .grad view points into the symmetric buffer
        # (measured, 35.2 MB: 2 GPUs multimem 121 us vs NCCL 85 us - every byte, the rank's own included, crosses the
        # switch; 8 GPUs: see DESIGN.md §6 - so the default is multimem from 4 ranks up, OWL_DP_MULTIMEM=1 forces it)
        self.multimem = False
        want = _os.environ.get("OWL_DP_MULTIMEM", "auto")
        if self._world > 1 and world is None and (want == "1" or (want == "auto" and self._world >= 4)):
            self.multimem = model.use_symmetric_grads(group)
        self._buckets = model.engine.grad_buckets()
        import os
        # Default: ONE all-reduce of the whole flat buffer after the backward pass, inside the step graph.  OWL_DP_BUCKETS=3
        # reduces three buckets on a forked stream as the backward pass completes them; measured SLOWER on 2 x B200
        # (3.641 vs 3.623 ms per step; single GPU 3.50): the NCCL kernels take SMs away from the persistent GEMM grids
        # (one CTA per SM, static tile schedule), whose tail then waits for the evicted CTAs.
        self._bucketed = os.environ.get("OWL_DP_BUCKETS", "1") == "3"
        self._one_graph = True          # N > 1: collectives captured inside the step graph (falls back if capture fails)
        self._next_load = 0
        self._next_run = 0
        self._results_read = 0
        self.status_every = 64          # result() checks the matcher status word every this many steps (one 4-byte D2H)

    # ------------------------------------------------------------------ data
    def load(self, image, labels, boxes, num_targets, slot: Optional[int] = None) -> int:
        """Asynchronous copy of one batch (host pinned or device tensors) into an input slot."""
        if slot is None:
            slot = self._next_load
            self._next_load = (self._next_load + 1) % len(self.slots)
        s = self.slots[slot]
        self.copy_stream.wait_event(self.consumed[slot])
        with torch.cuda.stream(self.copy_stream):
            s["image"].copy_(image, non_blocking=True)
            s["labels"].copy_(labels, non_blocking=True)
            s["boxes"].copy_(boxes, non_blocking=True)
            s["nt"].copy_(num_targets, non_blocking=True)
            self.loaded[slot].record(self.copy_stream)
        return slot

    # ------------------------------------------------------------------ the step
    def _fwd_bwd(self, slot: int) -> None:
        """reference main.py:74-90 (zero_grad, forward, criterion, backward) as a straight kernel sequence: the same
        Engine / matcher / loss entry points `OwlViT.forward`, `PushPullLoss.forward` and autograd's backward reach,
        without the autograd glue between them (no loss adds, no `stack`, no ones / zeros tensors for the graph)."""
        from . import ops
        s = self.slots[slot]
        model, crit = self.model, self.criterion
        eng = model.engine
        model._check_policy()
        model.zero_grad(set_to_none=False)                      # the backward kernels accumulate into flat_grad
        boxes, sims = eng.forward(s["image"], save_for_backward=True)
        lab, box, nt = s["labels"], s["boxes"], s["nt"]
        if crit.scales is not None and crit.scales.device != sims.device:
            crit.scales = crit.scales.to(sims.device)
        buf = crit.matcher.assign(sims, boxes, lab, box, nt)
        ops.match_loss(sims, boxes, lab, box, nt, buf.match, crit.scales, crit.background_label,
                       tc_matched=buf.tc_matched, tc_final=buf.tc_final, pred_sorted=buf.pred_sorted,
                       tgt_sorted=buf.tgt_sorted, losses_per_image=buf.losses_per_image, losses_mean4=self.losses[slot],
                       dsims_unit=buf.dsims_unit, dl1=buf.dl1, dgiou=buf.dgiou)
        # d(loss_ce + loss_bg + loss_bbox + loss_giou) / d(each loss) = 1   (reference main.py:84-90)
        ops.loss_backward(buf.dsims_unit, buf.tc_final, buf.match, buf.dl1, buf.dgiou, self._ones4,
                          crit.background_label, self._dsims, self._dboxes)
        inside = self._world > 1 and (self._one_graph or not self.use_graph)
        eng.backward(self._dsims, self._dboxes, model.flat_grad,
                     on_ready=self._reduce_bucket if (inside and self._bucketed) else None)
        if inside and self._bucketed:
            torch.cuda.current_stream().wait_stream(self.comm_stream)     # join: every bucket is reduced
        elif inside:
            model.allreduce_grads(self.group)

    def _reduce_bucket(self, i: int) -> None:
        """Data parallelism: all-reduce (sum; the 1/world is folded into the AdamW kernel) of gradient
        bucket i on the communication stream, as soon as the backward kernels that complete it are enqueued - the
        reduction of the heads' and the MLP's gradients runs under the kernels of the rest of the backward pass, only
        the last bucket (attention projections, 27 % of the bytes) is exposed.  Inside a graph capture the side stream
        forks from and joins the capturing stream, so the collectives become nodes of the step's ONE graph."""
        import torch.distributed as dist
        lo, hi = self._buckets[i]
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.model.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    def _capture(self, fn):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()                                     # warm-up: allocates workspaces, sets kernel attributes
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # thread_local: other threads of the process (the NCCL watchdog) may call CUDA APIs during the capture
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                fn()
        cur.wait_stream(side)
        return g

    def warmup(self) -> None:
        """Captures the graphs.  NOTE: the capture warm-ups run real steps on whatever the slots hold."""
        if not self.use_graph:
            return
        state = (self.model.flat_params.clone(), self.optimizer.exp_avg.clone(), self.optimizer.exp_avg_sq.clone(),
                 self.optimizer.state.clone())
        def whole(i):
            self._fwd_bwd(i)
            self.optimizer.step()
        if self._world == 1 or self._one_graph:
            # the whole step is ONE graph: on a single GPU nothing sits between the backward and the optimizer; under
            # data parallelism the bucketed all-reduces are captured on a forked stream inside the same graph
            try:
                for i in range(len(self.slots)):
                    if self._fwdbwd[i] is None:
                        self._fwdbwd[i] = self._capture(lambda i=i: whole(i))
            except Exception as e:
                if self._world == 1:
                    raise
                import warnings
                warnings.warn(f"capturing the collectives into the step graph failed ({e!r}); using two graphs around "
                              "an eager all-reduce")
                torch.cuda.synchronize()
                self._one_graph = False
                self._fwdbwd = [None] * len(self.slots)
        if self._world > 1 and not self._one_graph:
            for i in range(len(self.slots)):
                if self._fwdbwd[i] is None:
                    self._fwdbwd[i] = self._capture(lambda i=i: self._fwd_bwd(i))
            if self._opt_graph is None:
                self._opt_graph = self._capture(self.optimizer.step)
        # undo the optimizer steps the capture warm-ups made
        self.model.flat_params.copy_(state[0])
        self.optimizer.exp_avg.copy_(state[1])
        self.optimizer.exp_avg_sq.copy_(state[2])
        self.optimizer.state.copy_(state[3])
        self.model.engine.refresh_shadow()

    def launch_description(self) -> str:
        if not self.use_graph:
            return "eager launches"
        if self._world == 1:
            return "one CUDA-graph replay per step (fwd + loss + bwd + AdamW)"
        if self._one_graph:
            if self.multimem:
                return ("one CUDA-graph replay per step (fwd + loss + bwd + in-switch multimem all-reduce of the flat grad "
                        "buffer [own kernel, NVLS] + AdamW)")
            if not self._bucketed:
                return "one CUDA-graph replay per step (fwd + loss + bwd + one NCCL all-reduce of the flat grad buffer + AdamW)"
            return ("one CUDA-graph replay per step (fwd + loss + bwd + 3 bucketed NCCL all-reduces on a forked stream, "
                    "overlapped with the rest of the backward + AdamW)")
        return "two CUDA-graph replays per step (fwd+loss+bwd, AdamW) around the NCCL all-reduce"

    def result(self, slot: int):
        """The four losses of the last `run(slot, readback=True)` as Python floats (waits for that copy only)."""
        self.read_done[slot].synchronize()
        self._results_read += 1
        if self._results_read % self.status_every == 0:
            self.criterion.check_status()       # what the reference asserts inline (degenerate boxes, bad labels)
        return self.host_losses[slot].tolist()

    def run(self, slot: Optional[int] = None, readback: bool = False) -> torch.Tensor:
        if slot is None:
            slot = self._next_run
            self._next_run = (self._next_run + 1) % len(self.slots)
        cur = torch.cuda.current_stream()
        cur.wait_event(self.loaded[slot])
        if self.use_graph:
            if self._fwdbwd[slot] is None:
                self.warmup()
            self._fwdbwd[slot].replay()
        else:
            self._fwd_bwd(slot)
        self.consumed[slot].record(cur)
        if not self.use_graph:
            self.optimizer.step()                      # (_fwd_bwd already reduced the buckets)
        elif self._world > 1 and not self._one_graph:
            self.model.allreduce_grads(self.group)     # fallback: eager all-reduce between two graph replays
            self._opt_graph.replay()
        if readback:
            self.host_losses[slot].copy_(self.losses[slot], non_blocking=True)
            self.read_done[slot].record(cur)
        return self.losses[slot]
