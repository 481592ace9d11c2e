"""Per-phase clock64 sums of the second-generation fused attention kernel (dev tool; OWL_FA_GEN=94 selects the
instrumented instantiation).  Prints the mean cycles per sub-block of each loop phase of the MMA warp and of one
softmax thread, for the full and the partial query tiles and for the first / last launch wave."""
import os, sys, ctypes
os.environ.setdefault("OWL_FA_GEN", "94")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from owl_vit_object_detection_b200 import ops, _lib
B, S, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 577, 12)
dh = 64
D = H * dh
qkv = torch.randn((B * S, 3 * D), device="cuda").half()
ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
tiles = (S + 127) // 128
n_cta = tiles * H * B
dbg = torch.zeros(16 * n_cta + 8 * 128, dtype=torch.int64, device="cuda")
_lib.lib().owl_flash_attn_debug(ctypes.c_void_p(dbg.data_ptr()))
for _ in range(3):
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
torch.cuda.synchronize()
_lib.lib().owl_flash_attn_debug(ctypes.c_void_p(0))
raw = dbg.cpu().numpy()
t = raw[:16 * n_cta].reshape(n_cta, 16).astype(np.float64)
n_sub = (S + 31) // 32
n_full = (S // 128) * H * B
names = ["mma: wait P", "mma: PV issue", "mma: S issue(+kv wait)", "mma: load", "mma warp total",
         "smx: wait S", "smx: tmem ld", "smx: max", "smx: exp+st", "smx: rescale", "smx: st wait+arrive", "smx: prologue", "smx total"]
def show(label, rows):
    if len(rows) == 0:
        return
    m = t[rows].mean(0)
    print(f"--- {label}: {len(rows)} CTAs, {n_sub} sub-blocks each")
    for i, nm in enumerate(names):
        per = "" if nm.endswith("total") or nm.endswith("prologue") else f"   {m[i] / n_sub:8.0f} clk / sub-block"
        print(f"  {nm:24s} {m[i]:10.0f} clk{per}")
idx = np.arange(n_cta)
show("full tiles, first wave", idx[(idx < n_full) & (idx < 592)])
show("full tiles, later", idx[(idx < n_full) & (idx >= 592)])
show("partial tiles", idx[idx >= n_full])

# event timeline of CTA 0 (clock64, same SM): M = MMA warp, T = first softmax thread
n_sub_ = n_sub
ev = raw[16 * n_cta:16 * n_cta + 8 * n_sub_].reshape(n_sub_, 8).astype(np.int64)
base = ev[ev > 0].min()
lab = ["M P seen", "M PV issued", "M S(t+2) issued", "M load done", "T S seen", "T ld done", "T exp done", "T arrived"]
print("--- CTA 0 timeline (clk since first event)")
for tt in range(min(n_sub_, 12)):
    print(f"t={tt:2d} " + "  ".join(f"{lab[i]} {ev[tt, i] - base:6d}" for i in range(8)))
