"""NCCL all-reduce time of the flat gradient buffer (35.2 MB fp32) and of its three buckets (dev tool; under torchrun)."""
import os, sys
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (8_792_064, 4_725_000, 2_487_000, 1_580_000):
    x = torch.zeros(n, device="cuda")
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"all_reduce {n * 4 / 1e6:.1f} MB x{dist.get_world_size()}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
