"""Numerics of the tcgen05 GEMM (owl_gemm) against a plain fp32/fp64 torch reference of the same op.
Tolerance: inputs are fp16-exact in both paths, accumulation is fp32 in TMEM, so the only difference to
an fp64 reference is fp32 accumulation order: |err| <= 2e-3 * sqrt(K)-scaled magnitude (stated per test)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from owl_vit_object_detection_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.float16).cuda()


def _check(got, ref, tol):
    err = (got.double() - ref.double()).abs().max().item()
    mag = ref.double().abs().max().item()
    assert err <= tol * max(mag, 1.0), f"max err {err} vs magnitude {mag}"


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("shape", [(300, 200, 136), (128, 256, 64), (577, 768, 768)])
def test_kk_f32_plain(bn, shape):
    ops = _ops()
    M, N, K = shape
    a, b = _rand((M, K), 1), _rand((N, K), 2)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(a, b, out, M=M, N=N, K=K, bn=bn)
    torch.cuda.synchronize()
    _check(out, a.double() @ b.double().T, 1e-4)


def test_kk_f16_bias_qgelu_prestore():
    ops = _ops()
    M, N, K = 1000, 3072, 768
    a, b = _rand((M, K), 3), _rand((N, K), 4, K ** -0.5)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    pre = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias, act="quick_gelu", pre_out=pre)
    torch.cuda.synchronize()
    z = a.double() @ b.double().T + bias.double()
    _check(pre, z, 2e-3)
    _check(out, z * torch.sigmoid(1.702 * z), 2e-3)


def test_kk_f32_resid_pos_remap():
    ops = _ops()
    P, D, K, B = 576, 768, 3072, 2
    M = B * P
    a, b = _rand((M, K), 5, 0.5), _rand((D, K), 6, K ** -0.5)
    pos = torch.randn(P + 1, D, device="cuda")
    out = torch.zeros((B * (P + 1), D), device="cuda")
    ops.gemm(a, b, out, M=M, N=D, K=K, pos=pos, rows_per_img=P)
    torch.cuda.synchronize()
    ref = (a.double() @ b.double().T).view(B, P, D) + pos[1:].double()
    _check(out.view(B, P + 1, D)[:, 1:], ref, 1e-4)
    assert out.view(B, P + 1, D)[:, 0].abs().max().item() == 0.0
    # residual + bias, accumulate mode
    M2 = 700
    a2 = _rand((M2, 768), 7)
    w2 = _rand((768, 768), 8, 768 ** -0.5)
    bias = torch.randn(768, device="cuda")
    resid = torch.randn(M2, 768, device="cuda")
    out2 = torch.ones((M2, 768), device="cuda")
    ops.gemm(a2, w2, out2, M=M2, N=768, K=768, bias=bias, resid=resid, out_mode=1)
    torch.cuda.synchronize()
    _check(out2, a2.double() @ w2.double().T + bias.double() + resid.double() + 1.0, 1e-4)


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_dgrad_b_mn_major(bn):
    # dX[M,K] = dY[M,N] @ W[N,K]  ->  GEMM with reduction over N, B = W read MN-major
    ops = _ops()
    M, N, K = 333, 200, 264
    dy, w = _rand((M, N), 9), _rand((N, K), 10)
    out = torch.zeros((M, K), device="cuda")
    ops.gemm(dy, w, out, M=M, N=K, K=N, b_mn=True, bn=bn)
    torch.cuda.synchronize()
    _check(out, dy.double() @ w.double(), 1e-4)


def test_dgrad_f16_with_act_grad():
    ops = _ops()
    M, N, K = 500, 768, 3072   # dH[M,3072] = dY[M,768] @ W2[768,3072] * qgelu'(pre)
    dy, w = _rand((M, N), 11), _rand((N, K), 12, N ** -0.5)
    pre = _rand((M, K), 13)
    out = torch.zeros((M, K), device="cuda", dtype=torch.float16)
    ops.gemm(dy, w, out, M=M, N=K, K=N, b_mn=True, act="quick_gelu_grad", act_src=pre)
    torch.cuda.synchronize()
    x = pre.double()
    s = torch.sigmoid(1.702 * x)
    ref = (dy.double() @ w.double()) * (s + 1.702 * x * s * (1 - s))
    _check(out, ref, 2e-3)


@pytest.mark.parametrize("split_k", [1, 4, 7])
@pytest.mark.parametrize("bn", [64, 256])
def test_wgrad_mn_mn_split_k(split_k, bn):
    # dW[N,K] = dY[M,N]^T @ X[M,K]: reduction over M, both operands read MN-major
    ops = _ops()
    M, N, K = 1154, 200, 328
    dy, x = _rand((M, N), 14), _rand((M, K), 15)
    out = torch.zeros((N, K), device="cuda")
    ops.gemm(dy, x, out, M=N, N=K, K=M, a_mn=True, b_mn=True, split_k=split_k, bn=bn, out_mode=2)
    torch.cuda.synchronize()
    _check(out, dy.double().T @ x.double(), 1e-4)


def test_batched_attention_shapes():
    # S = Q K^T and O = P V for (image, head) batches addressed inside a packed [tokens, 3*hidden] buffer
    ops = _ops()
    B, H, S, dh = 2, 3, 577, 64
    D = H * dh
    qkv = _rand((B * S, 3 * D), 16)
    Sp = 584
    scores = torch.zeros((B * H, S, Sp), device="cuda", dtype=torch.float16)
    ops.gemm(qkv, qkv[:, D:], scores, M=S, N=S, K=dh, a_ld=3 * D, b_ld=3 * D, ldo=Sp,
             batches_outer=B, heads=H, a_outer_stride=S * 3 * D, b_outer_stride=S * 3 * D,
             a_head_col=dh, b_head_col=dh, o_outer_stride=H * S * Sp, o_head_stride=S * Sp, alpha=0.125)
    torch.cuda.synchronize()
    q = qkv[:, :D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    k = qkv[:, D:2 * D].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    v = qkv[:, 2 * D:].view(B, S, H, dh).permute(0, 2, 1, 3).double()
    ref = 0.125 * q @ k.transpose(-1, -2)
    _check(scores.view(B, H, S, Sp)[..., :S], ref, 2e-3)
    assert scores.view(B, H, S, Sp)[..., S:].abs().max().item() == 0.0
    p = torch.softmax(ref, -1).to(torch.float16)
    pbuf = torch.zeros((B * H, S, Sp), device="cuda", dtype=torch.float16)
    pbuf[..., :S] = p.view(B * H, S, S)
    pbuf[..., S:] = 7.0  # padding columns must be ignored (tensor-map bound = S)
    o = torch.zeros((B * S, D), device="cuda", dtype=torch.float16)
    ops.gemm(pbuf, qkv[:, 2 * D:], o, M=S, N=dh, K=S, b_mn=True, a_ld=Sp, b_ld=3 * D, ldo=D,
             batches_outer=B, heads=H, a_outer_stride=H * S * Sp, a_head_stride=S * Sp,
             b_outer_stride=S * 3 * D, b_head_col=dh, o_outer_stride=S * D, o_head_stride=dh)
    torch.cuda.synchronize()
    ref_o = (p.double() @ v).permute(0, 2, 1, 3).reshape(B * S, D)
    _check(o, ref_o, 2e-3)


def test_pool3_epilogue():
    ops = _ops()
    M, E, C = 1152, 512, 80
    a = _rand((M, E), 17, E ** -0.5)
    q = _rand((3 * C, E), 18)
    sims = torch.zeros((M, C), device="cuda")
    arg = torch.zeros((M, C), device="cuda", dtype=torch.uint8)
    ops.gemm(a, q, sims, M=M, N=3 * C, K=E, pool3=True, argmax=arg)
    torch.cuda.synchronize()
    full = (a.double() @ q.double().T).view(M, C, 3)
    _check(sims, full.max(-1).values, 1e-4)
    picked = torch.gather(full, 2, arg.long()[..., None])[..., 0]
    _check(picked, full.max(-1).values, 1e-4)


def test_full_size_qkv_timing():
    ops = _ops()
    M, N, K = 16 * 577, 2304, 768
    a, b = _rand((M, K), 19), _rand((N, K), 20, K ** -0.5)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    for _ in range(3):
        ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"\nqkv gemm {M}x{N}x{K}: {ms * 1e3:.1f} us, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    _check(out, a.double() @ b.double().T + bias.double(), 2e-3)


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("shape", [(300, 512, 136), (1154, 768, 768), (256, 256, 64), (9232, 768, 192)])
def test_cluster_multicast_kk(bn, shape):
    """2-CTA clusters (B tile TMA-multicast): odd numbers of M tiles, M/N/K tails, both epilogue families."""
    ops = _ops()
    M, N, K = shape
    a, b = _rand((M, K), 21), _rand((N, K), 22, K ** -0.5)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(a, b, out, M=M, N=N, K=K, bn=bn, bias=bias, resid=resid, cluster_m=2)
    out16 = torch.zeros((M, N), device="cuda", dtype=torch.float16)
    ops.gemm(a, b, out16, M=M, N=N, K=K, bn=bn, bias=bias, act="quick_gelu", cluster_m=2)
    torch.cuda.synchronize()
    z = a.double() @ b.double().T + bias.double()
    _check(out, z + resid.double(), 1e-4)
    _check(out16, z * torch.sigmoid(1.702 * z), 2e-3)


@pytest.mark.parametrize("bn", [128, 256])
def test_cluster_multicast_dgrad_b_mn(bn):
    ops = _ops()
    M, N, K = 1200, 520, 264          # dX[M,K] = dY[M,N] @ W[N,K], B = W read MN-major
    dy, w = _rand((M, N), 23), _rand((N, K), 24)
    out = torch.zeros((M, K), device="cuda")
    ops.gemm(dy, w, out, M=M, N=K, K=N, b_mn=True, bn=bn, cluster_m=2)
    torch.cuda.synchronize()
    _check(out, dy.double() @ w.double(), 1e-4)
