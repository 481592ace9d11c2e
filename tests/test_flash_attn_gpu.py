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
                                    (1, 1300, 2), (2, 65, 2), (1, 129, 3), (1, 64, 1), (1, 1, 1), (1, 3601, 1)])
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


@pytest.mark.parametrize("B,S,H", [(2, 577, 12), (1, 209, 2)])
def test_flash_attn_lse_and_backward_pieces(B, S, H):
    """The attention backward of the last layer (autograd of HF:393-404) is built from: the log-sum-exp saved by the
    forward, P recomputed in a GEMM epilogue (act exp_row), the row term delta (owl_attn_delta) and the softmax
    backward fused into the dP GEMM epilogue (act softmax_grad).  Each piece against fp64 torch; tolerances: lse
    1e-3 abs (fp32 exp2/log2 approximations), P 2e-3 abs (fp16 storage), dS 1e-2 of its magnitude (fp16 P, dctx)."""
    from owl_vit_object_detection_b200 import ops
    dh = 64
    D = H * dh
    Sp = (S + 7) // 8 * 8
    scale = dh ** -0.5
    g = torch.Generator().manual_seed(S + H)
    qkv = torch.randn((B * S, 3 * D), generator=g).half().cuda()
    dctx = torch.randn((B * S, D), generator=g).half().cuda()
    ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
    lse = torch.zeros((B * H, S), dtype=torch.float32, device="cuda")
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=scale, lse=lse)
    q = qkv[:, :D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    k = qkv[:, D:2 * D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    v = qkv[:, 2 * D:].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    s = q @ k.transpose(-1, -2) * scale
    ref_lse = torch.logsumexp(s, -1).reshape(B * H, S)
    assert (lse.double() - ref_lse).abs().max().item() <= 1e-3
    # P = exp(scale q k^T - lse)
    probs = torch.zeros((B * H, S, Sp), dtype=torch.float16, device="cuda")
    ops.gemm(qkv, qkv[:, D:], probs, M=S, N=S, K=dh, a_ld=3 * D, b_ld=3 * D, ldo=Sp, batches_outer=B, heads=H,
             a_outer_stride=S * 3 * D, b_outer_stride=S * 3 * D, a_head_col=dh, b_head_col=dh,
             o_outer_stride=H * S * Sp, o_head_stride=S * Sp, alpha=scale, act="exp_row", rowvec=lse, rowvec_stride=S)
    ref_p = torch.softmax(s, -1).reshape(B * H, S, S)
    assert (probs[:, :, :S].double() - ref_p).abs().max().item() <= 2e-3
    assert probs[:, :, S:].abs().max().item() == 0 if Sp > S else True
    # delta = scale * sum_d dctx * ctx
    delta = torch.zeros((B * H, S), dtype=torch.float32, device="cuda")
    ops.attn_delta(ctx, dctx, delta, B=B, S=S, H=H, head_dim=dh, alpha=scale)
    ref_delta = scale * (ctx.double() * dctx.double()).view(B, S, H, dh).sum(-1).permute(0, 2, 1).reshape(B * H, S)
    assert (delta.double() - ref_delta).abs().max().item() <= 1e-4 * max(1.0, ref_delta.abs().max().item())
    # dS = P * (scale dctx v^T - delta)
    ds = torch.zeros((B * H, S, Sp), dtype=torch.float16, device="cuda")
    ops.gemm(dctx, qkv[:, 2 * D:], ds, M=S, N=S, K=dh, a_ld=D, b_ld=3 * D, ldo=Sp, batches_outer=B, heads=H,
             a_outer_stride=S * D, b_outer_stride=S * 3 * D, a_head_col=dh, b_head_col=dh,
             o_outer_stride=H * S * Sp, o_head_stride=S * Sp, alpha=scale, act="softmax_grad", act_src=probs,
             act_src_outer_stride=H * S * Sp, act_src_head_stride=S * Sp, rowvec=delta, rowvec_stride=S)
    torch.cuda.synchronize()
    do = dctx.view(B, S, H, dh).permute(0, 2, 1, 3).double()
    dp = do @ v.transpose(-1, -2)
    p64 = torch.softmax(s, -1)
    ref_ds = (p64 * (dp - (p64 * dp).sum(-1, keepdim=True)) * scale).reshape(B * H, S, S)
    err = (ds[:, :, :S].double() - ref_ds).abs().max().item()
    assert err <= 1e-2 * ref_ds.abs().max().item(), (err, ref_ds.abs().max().item())


@pytest.mark.parametrize("B,S,H", [(2, 577, 12), (1, 209, 2), (1, 128, 1), (2, 300, 3), (1, 65, 1)])
def test_fused_attention_backward(B, S, H):
    """owl_attn_bwd (dq, dk, dv in one kernel, scores in TMEM) against fp64 autograd of softmax(q k^T / sqrt(dh)) v.
    Tolerance: fp16 q/k/v/dctx/P/dS operands, fp32 accumulation -> 1e-2 of each gradient's magnitude."""
    from owl_vit_object_detection_b200 import ops
    dh = 64
    D = H * dh
    scale = dh ** -0.5
    g = torch.Generator().manual_seed(7 * S + H)
    qkv = torch.randn((B * S, 3 * D), generator=g).half().cuda()
    dctx = torch.randn((B * S, D), generator=g).half().cuda()
    ctx = torch.zeros((B * S, D), dtype=torch.float16, device="cuda")
    lse = torch.zeros((B * H, S), dtype=torch.float32, device="cuda")
    delta = torch.zeros((B * H, S), dtype=torch.float32, device="cuda")
    ops.flash_attn_fwd(qkv, ctx, B=B, S=S, H=H, head_dim=dh, scale=scale, lse=lse)
    ops.attn_delta(ctx, dctx, delta, B=B, S=S, H=H, head_dim=dh, alpha=scale)
    dqkv = torch.full((B * S, 3 * D), float("nan"), dtype=torch.float16, device="cuda")
    dq32 = torch.empty((B * S, D), dtype=torch.float32, device="cuda")
    ops.attn_bwd(qkv, dctx, lse, delta, dqkv, dq32, B=B, S=S, H=H, head_dim=dh, scale=scale)
    torch.cuda.synchronize()
    x = qkv.double().view(B, S, 3, H, dh).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)   # [3, B, H, S, dh]
    p = torch.softmax(x[0] @ x[1].transpose(-1, -2) * scale, -1)
    o = (p @ x[2]).permute(0, 2, 1, 3).reshape(B * S, D)
    o.backward(dctx.double())
    ref = x.grad.permute(1, 3, 0, 2, 4).reshape(B * S, 3 * D)       # back to the packed layout
    assert torch.isfinite(dqkv).all()
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        err = (dqkv[:, sl].double() - ref[:, sl]).abs().max().item()
        mag = ref[:, sl].abs().max().item()
        assert err <= 1e-2 * mag, (name, err, mag)
