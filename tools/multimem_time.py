"""Times the in-switch all-reduce of the flat gradient buffer (barrier + multimem kernel + barrier) against NCCL
(dev tool; under torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from owl_vit_object_detection_b200.collective import SymmetricGrad
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 8_792_064
sg = SymmetricGrad.create(n, dev)
x = torch.zeros(n, device=dev)
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
if sg is None:
    if rank == 0: print("no multicast support")
else:
    sg.buf.fill_(rank + 1.0)
    sg.all_reduce(); torch.cuda.synchronize()
    w = dist.get_world_size()
    ok = bool((sg.buf == w * (w + 1) / 2).all())
    a = t(sg.all_reduce)
    b = t(lambda: sg.handle.barrier(channel=0))
    c = t(lambda: dist.all_reduce(x))
    if rank == 0: print(f"x{w}: multimem all-reduce 35.2 MB (2 barriers + kernel) {a:.1f} us; one barrier {b:.1f} us; NCCL {c:.1f} us; correct={ok}", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
