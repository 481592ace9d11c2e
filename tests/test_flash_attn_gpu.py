"""GPU parity of the fused tcgen05 attention forward (owl_flash_attn_fwd) against an fp64 torch reference of
HF:379-404 (softmax(q k^T / sqrt(dh)) v).  Tolerance (stated): fp16 q/k/v/P, fp32 accumulation -> |err| <= 2e-3 of
the output magnitude."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, B, S, H, dh):
    D = H * dh
    q = qkv[:, :D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    k = qkv[:, D:2 * D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    v = qkv[:, 2 * D:].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    p = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, D)


@pytest.mark.parametrize("B,S,H", [(2, 577, 12), (1, 17, 2), (3, 208, 1), (1, 209, 2), (2, 416, 3), (1, 625, 2),
                                    (1, 1300, 2)])
def test_flash_attn_fwd(B, S, H):
    from owl_vit_object_detection_b200 import ops
    dh = 64
    D = H * dh
    g = torch.Generator().manual_seed(S * 31 + H)
    qkv = (torch.randn((B * S, 3 * D), generator=g) * 1.5).half().cuda()
    ctx = torch.full((B * S, D), float("nan"), dtype=torch.float16, device="cuda")
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
    torch.cuda.synchronize()
    ref = _ref(qkv, B, S, H, dh)
    err = (ctx.double() - ref).abs().max().item()
    mag = ref.abs().max().item()
    print(f"\nflash attn B={B} S={S} H={H}: max err {err:.2e} (|ref| max {mag:.2f})")
    assert err <= 2e-3 * max(mag, 1.0)


def test_flash_attn_peaky_rows():
    """Rows whose maximum grows from block to block exercise the O rescaling path."""
    from owl_vit_object_detection_b200 import ops
    B, S, H, dh = 1, 577, 2, 64
    D = H * dh
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn((B * S, 3 * D), generator=g)
    qkv[:, D:2 * D] *= torch.linspace(0.5, 6.0, S)[:, None]      # later keys score higher
    qkv = qkv.half().cuda()
    ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
    torch.cuda.synchronize()
    ref = _ref(qkv, B, S, H, dh)
    assert (ctx.double() - ref).abs().max().item() <= 4e-3 * max(ref.abs().max().item(), 1.0)


def test_flash_attn_timing_b16():
    from owl_vit_object_detection_b200 import ops
    B, S, H, dh = 16, 577, 12, 64
    D = H * dh
    qkv = torch.randn((B * S, 3 * D), device="cuda").half()
    ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
    for _ in range(3):
        ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=dh ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 4.0 * B * H * S * S * dh
    print(f"\nflash attn fwd B=16 S=577 H=12: {ms * 1e3:.1f} us, {fl / ms / 1e9:.1f} TFLOP/s")
