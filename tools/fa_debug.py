"""Where does the fused attention differ from fp64?  (dev tool)  usage: python tools/fa_debug.py B S H"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from owl_vit_object_detection_b200 import ops
B, S, H = (int(x) for x in sys.argv[1:4])
dh = 64; D = H * dh
g = torch.Generator().manual_seed(S * 31 + H)
qkv = (torch.randn((B * S, 3 * D), generator=g) * 1.5).half().cuda()
ctx = torch.full((B * S, D), float("nan"), dtype=torch.float16, device="cuda")
for rep in range(3):
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
    torch.cuda.synchronize()
    q = qkv[:, :D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    k = qkv[:, D:2 * D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    v = qkv[:, 2 * D:].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    sc = q @ k.transpose(-1, -2) * dh ** -0.5
    ref = (torch.softmax(sc, -1) @ v)
    out = ctx.view(B, S, H, dh).permute(0, 2, 1, 3).double()
    err = (out - ref).abs().amax(-1)          # [B, H, S]
    bad = (err > 5e-3).nonzero()
    print(f"rep {rep}: S={S} max err {err.max().item():.3e}, bad rows {len(bad)}")
    if len(bad):
        rows = bad[:, 2]
        print("  bad (b,h) pairs:", sorted(set((int(a), int(b)) for a, b, _ in bad.tolist()))[:8])
        print("  bad rows min/max:", int(rows.min()), int(rows.max()), " tiles:", sorted(set((rows // 128).tolist())), " row%32 set size:", len(set((rows % 128).tolist())))
        # is the bad output consistent with dropping / duplicating a key range?
        b_, h_, r_ = bad[0].tolist()
        p = torch.softmax(sc[b_, h_, r_], -1)
        print("  first bad row", (b_, h_, r_), "row max prob", float(p.max()), "argmax key", int(p.argmax()))
        for lo in range(0, S, 32):
            pm = p.clone(); pm[lo:lo + 32] = 0
            alt = (pm / pm.sum()) @ v[b_, h_]
            if (alt - out[b_, h_, r_]).abs().max() < 2e-3:
                print("  -> matches attention WITHOUT keys", lo, "..", lo + 31)
