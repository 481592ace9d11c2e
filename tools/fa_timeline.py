"""Phase timeline of the fused attention kernel (dev tool).  Needs the instrumented build:
    make -C owl_vit_object_detection_b200/csrc clean && make -C owl_vit_object_detection_b200/csrc -j8 FA_TIMELINE=1"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops, _lib
B, S, H, dh = 16, 577, 12, 64
D = H * dh
qkv = torch.randn((B * S, 3 * D), device="cuda").half()
ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
dbg = torch.zeros(64 + 3 * 5 * H * B, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
torch.cuda.synchronize()
_lib.lib().owl_flash_attn_debug(ctypes.c_void_p(dbg.data_ptr()))
ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
torch.cuda.synchronize()
_lib.lib().owl_flash_attn_debug(ctypes.c_void_p(0))
t = dbg.cpu().tolist()
t0 = t[0]
names = {0: "setup done", 1: "after pdl wait", 30: "PV last done", 31: "epilogue stored"}
for j in range(5):
    names[2 + 4 * j] = f"S{j} ready"; names[3 + 4 * j] = f"max{j} known"; names[4 + 4 * j] = f"P{j} written"; names[5 + 4 * j] = f"p_full{j} arrived"
for j in range(5):
    names[40 + 2 * j] = f"  ctl: p_full{j} seen"; names[41 + 2 * j] = f"  ctl: PV{j}+S{j+1} issued"
for k in sorted(names, key=lambda k: t[k]):
    if t[k]: print(f"{names[k]:18s} +{(t[k] - t0) / 1000:.2f} us")

import numpy as np
c = np.array(t[64:]).reshape(-1, 3)
st, en, sm = c[:, 0] - c[:, 0].min(), c[:, 1] - c[:, 0].min(), c[:, 2]
dur = (en - st) / 1000.0
print(f"CTAs {len(c)}  kernel span {en.max() / 1000:.1f} us  CTA duration mean {dur.mean():.2f} min {dur.min():.2f} max {dur.max():.2f} us")
qt = np.arange(len(c)) % 5
for q in range(5):
    print(f"  q-tile {q}: mean {dur[qt == q].mean():.2f} us")
per_sm = np.bincount(sm, minlength=148)
print("CTAs per SM: min", per_sm.min(), "max", per_sm.max())
# busy time per SM (union of intervals) and concurrency
busy = []
for s_ in range(148):
    iv = sorted(zip(st[sm == s_], en[sm == s_]))
    tot, cur_s, cur_e = 0, None, None
    for a, b in iv:
        if cur_e is None or a > cur_e:
            if cur_e is not None: tot += cur_e - cur_s
            cur_s, cur_e = a, b
        else:
            cur_e = max(cur_e, b)
    if cur_e is not None: tot += cur_e - cur_s
    busy.append(tot / 1000.0)
print(f"SM busy (union) mean {np.mean(busy):.1f} us, sum of CTA durations per SM mean {dur.sum() / 148:.1f} us")
order = np.argsort(st)
print("first 12 CTA starts (us):", np.round(st[order][:12] / 1000, 2), " last 6 ends:", np.round(np.sort(en)[-6:] / 1000, 2))
