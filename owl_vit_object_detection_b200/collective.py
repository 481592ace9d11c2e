"""Data-parallel gradient exchange over NVSwitch multicast.

`SymmetricGrad` puts the flat fp32 gradient buffer of an `OwlViT` into symmetric memory (torch.distributed provides the
allocation, the rendezvous and the signal-pad barrier: plumbing) and reduces it with ONE hand-written kernel per rank
(`owl_allreduce_multimem`: multimem.ld_reduce / multimem.st, csrc/collective.cu) between two cross-rank barriers.  All
of it is stream work, so `TrainStep` captures it inside the step graph.  When the platform has no multicast support
(no NVSwitch / fabric manager) `create()` returns None and the caller keeps the NCCL all-reduce.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class SymmetricGrad:
    def __init__(self, buf: torch.Tensor, handle, rank: int, world: int):
        self.buf, self.handle, self.rank, self.world = buf, handle, rank, world

    @staticmethod
    def create(n: int, device, group=None) -> Optional["SymmetricGrad"]:
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = group if group is not None else dist.group.WORLD
            buf = symm_mem.empty(n, dtype=torch.float32, device=device)
            handle = symm_mem.rendezvous(buf, group)
            if not getattr(handle, "multicast_ptr", 0):
                return None
            buf.zero_()
            return SymmetricGrad(buf, handle, dist.get_rank(group), dist.get_world_size(group))
        except Exception as e:      # no symmetric-memory support on this platform / build
            import warnings
            warnings.warn(f"symmetric gradient buffer not available ({e!r}); using the NCCL all-reduce")
            return None

    def all_reduce(self) -> None:
        """sum over ranks, in place, on the current stream."""
        with torch.cuda.device(self.buf.device):
            self.handle.barrier(channel=0)          # every rank's backward kernels have written their gradients
            ops.allreduce_multimem(int(self.handle.multicast_ptr), self.buf.numel(), self.rank, self.world)
            self.handle.barrier(channel=1)          # every rank's slice has been written into every copy
