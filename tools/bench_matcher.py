"""Matcher microbench (BASELINE.json configs[4]): 576 predictions x T targets, cost matrix + assignment, over N
synthetic images processed in chunks.  Prints us/image for the cost kernel, the LSAP kernel and both, plus an
exactness check of a sample of images against the oracle (index-exact) (dev tool; numbers go to profiles/).

usage: bench_matcher.py [n_images=1000000] [chunk=16384] [check=64]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
CHECK = int(sys.argv[3]) if len(sys.argv) > 3 else 64
P, C = 576, 80
dev = "cuda"


def gen(n, T, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    sims = torch.rand((n, P, C), generator=g, device=dev) * 0.4 - 0.1

    def boxes(k):
        cxy = 0.1 + 0.8 * torch.rand((n, k, 2), generator=g, device=dev)
        wh = 0.02 + 0.48 * torch.rand((n, k, 2), generator=g, device=dev)
        lo = (cxy - wh / 2).clamp(0.0, 1.0)
        hi = torch.maximum((cxy + wh / 2).clamp(0.0, 1.0), lo + 1e-3)
        return torch.cat([lo, hi], dim=-1).contiguous()
    return sims, boxes(P), torch.randint(0, C, (n, T), generator=g, device=dev), boxes(T)


for T in (10, 50, 100):
    n_chunks = (N + CH - 1) // CH
    costT = torch.empty((CH, T, P), device=dev)
    match = torch.empty((CH, T), dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nt = torch.full((CH,), T, dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_cost = t_lsap = 0.0
    mism = 0
    for c in range(n_chunks):
        sims, pred, lab, tgt = gen(CH, T, seed=1000 * T + c)
        if c == 0:   # warm-up
            ops.matcher_cost(sims, pred, lab, tgt, nt, costT, status)
            ops.lsap(costT, nt, match, status)
        torch.cuda.synchronize()
        ev[0].record()
        ops.matcher_cost(sims, pred, lab, tgt, nt, costT, status)
        ev[1].record()
        ops.lsap(costT, nt, match, status)
        ev[2].record()
        torch.cuda.synchronize()
        t_cost += ev[0].elapsed_time(ev[1])
        t_lsap += ev[1].elapsed_time(ev[2])
        if c == 0 and CHECK:
            from oracle import matcher_oracle as mo     # checker only
            sc, pc, lc, tc = sims[:CHECK].cpu(), pred[:CHECK].cpu(), lab[:CHECK].cpu(), tgt[:CHECK].cpu()
            m = match[:CHECK].cpu()
            for b in range(CHECK):
                rows, cols = mo.lsap(mo.cost_matrix(sc[b], pc[b], lc[b], tc[b]).numpy())
                exp = torch.full((T,), -1, dtype=torch.int32)
                exp[torch.from_numpy(cols)] = torch.from_numpy(rows).int()
                mism += int(not torch.equal(exp, m[b]))
    n_done = n_chunks * CH
    assert status.item() == 0
    bytes_img = P * C * 4 + P * 4 * 4 + T * 24 + P * T * 4
    print(f"T={T:3d}: {n_done} images  cost {t_cost * 1e3 / n_done:.4f} us/img ({bytes_img * n_done / t_cost / 1e6:.0f} GB/s algorithmic)"
          f"  lsap {t_lsap * 1e3 / n_done:.4f} us/img  total {(t_cost + t_lsap) * 1e3 / n_done:.4f} us/img"
          f"  | oracle check: {mism}/{CHECK} images differ", flush=True)
