"""GPU parity of the training-path backward kernels (through the C-ABI) against torch autograd in fp32 on the same
bf16-rounded operands (the reference differentiates exactly these torch ops: nn.Conv3d, nn.GroupNorm, nn.SiLU,
nn.Linear, the matmul DWT/IDWT -- guided_diffusion/nn.py:17-39, DWT_IDWT_Functions.py:115-208).

Tolerances: weight gradients are accumulated in fp32 from bf16 operands: |err| <= 2e-3 * max|ref| + 1e-4;
activation gradients are stored in bf16: |err| <= 1e-2 * max|ref| + 1e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


WGRAD_CASES = [
    # N, D, H, W, cin, cout, k
    (1, 4, 16, 8, 64, 64, 3),       # M=64 interleaved accumulators, one full tile per plane
    (1, 5, 18, 10, 64, 64, 3),      # ragged tiles in H and W
    (2, 3, 7, 5, 32, 64, 3),        # padded input channels, tiny dims, batch 2
    (1, 4, 16, 8, 64, 8, 3),        # output conv 64 -> 8 (dY zero padded to 64)
    (1, 6, 16, 16, 128, 128, 3),    # M=128, N=128
    (1, 4, 20, 12, 256, 64, 3),     # M=64, N=128 (WaveletDownsample 256 -> 64)
    (1, 4, 12, 8, 64, 128, 3),      # M=128, N=64
    (1, 3, 14, 10, 128, 256, 3),    # two M blocks
    (1, 6, 12, 24, 64, 128, 1),     # 1x1x1 skip conv
    (1, 2, 14, 10, 256, 128, 1),
    (1, 2, 8, 8, 128, 64, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3d_wgrad(case):
    from fcwdm import ops
    from gpu_util import bf16_round, to_cl
    _no_tf32()
    N, D, H, W, cin, cout, k = case
    g = torch.Generator().manual_seed(1)
    x = bf16_round(torch.randn(N, cin, D, H, W, generator=g)).cuda()
    dy = bf16_round(torch.randn(N, cout, D, H, W, generator=g)).cuda()
    xc, dyc = to_cl(x), to_cl(dy)
    dw = torch.full((cout, cin, k, k, k), 0.5, dtype=torch.float32, device="cuda")
    ops.conv3d_wgrad(xc, dyc, dw, (N, D, H, W), cin, cout, k, accumulate=False)
    ops.conv3d_wgrad(xc, dyc, dw, (N, D, H, W), cin, cout, k, accumulate=True)      # tied weights: second use adds
    torch.cuda.synchronize()
    w = torch.zeros(cout, cin, k, k, k, device="cuda", requires_grad=True)
    F.conv3d(x, w, None, padding=k // 2).backward(dy)
    ref = 2.0 * w.grad
    err = float((dw - ref).abs().max())
    tol = 2e-3 * float(ref.abs().max()) + 1e-4
    assert err <= tol, (err, tol)


def test_conv3d_wgrad_tap_selectivity():
    """Every tap lands in its own slot: x = one-hot voxel, dy = one-hot voxel -> exactly one non-zero tap."""
    from fcwdm import ops
    from gpu_util import to_cl
    _no_tf32()
    N, D, H, W, C = 1, 4, 16, 8, 64
    for (dz, dh, dw_) in [(-1, -1, -1), (0, 0, 0), (1, 1, 1), (-1, 0, 1), (1, -1, 0), (0, 1, -1)]:
        x = torch.zeros(N, C, D, H, W, device="cuda")
        dy = torch.zeros(N, C, D, H, W, device="cuda")
        dy[0, 3, 2, 7, 4] = 1.0
        x[0, 5, 2 + dz, 7 + dh, 4 + dw_] = 1.0
        dw = torch.zeros(C, C, 3, 3, 3, device="cuda")
        ops.conv3d_wgrad(to_cl(x), to_cl(dy), dw, (N, D, H, W), C, C, 3, accumulate=False)
        torch.cuda.synchronize()
        assert float(dw[3, 5, 1 + dz, 1 + dh, 1 + dw_]) == 1.0
        assert float(dw.abs().sum()) == 1.0


DGRAD_CASES = [(1, 4, 16, 8, 64, 64, 3), (1, 5, 18, 10, 64, 128, 3), (1, 4, 12, 8, 128, 64, 3), (1, 4, 16, 8, 64, 8, 3),
               (1, 6, 12, 24, 64, 128, 1), (1, 3, 14, 10, 256, 256, 3)]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("pair", [False, True])
def test_conv3d_dgrad(case, pair):
    """dX = forward conv kernel on dY with transposed, tap-reversed weights (+ fused gradient fan-in add)."""
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    _no_tf32()
    N, D, H, W, cin, cout, k = case
    if pair and not ops.conv3d_pair_supported(cout, cin, k):
        pytest.skip("pair kernel needs C <= 64, 3x3x3")
    g = torch.Generator().manual_seed(2)
    w = bf16_round(torch.randn(cout, cin, k, k, k, generator=g) / np.sqrt(cout * k ** 3)).cuda()
    dy = bf16_round(torch.randn(N, cout, D, H, W, generator=g)).cuda()
    acc = bf16_round(torch.randn(N, cin, D, H, W, generator=g)).cuda()
    wt = ops.conv3d_transpose_flip_weights(w)
    dxc = torch.zeros((N * D * H * W, (cin + 63) // 64 * 64), dtype=torch.bfloat16, device="cuda")
    if pair:
        ops.conv3d_pair_cl(to_cl(dy), ops.conv3d_pair_pack_weights(wt), None, dxc, (N, D, H, W), cout, cin,
                           residual=to_cl(acc))
    else:
        ops.conv3d_cl(to_cl(dy), ops.conv3d_pack_weights(wt), None, dxc, (N, D, H, W), cout, cin, k, residual=to_cl(acc))
    torch.cuda.synchronize()
    got = from_cl(dxc, (N, cin, D, H, W))
    x = torch.zeros(N, cin, D, H, W, device="cuda", requires_grad=True)
    F.conv3d(x, w, None, padding=k // 2).backward(dy)
    ref = x.grad + acc
    err = float((got - ref).abs().max())
    tol = 1e-2 * float(ref.abs().max()) + 1e-3
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("shape", [(2, 64, 6, 10, 8, 32), (1, 128, 4, 6, 10, 32), (1, 256, 3, 7, 5, 32), (1, 64, 4, 4, 4, 8)])
@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("with_acc", [False, True])
def test_groupnorm_silu_bwd(shape, silu, with_acc):
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    N, C, D, H, W, G = shape
    S = D * H * W
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(N, C, D, H, W, generator=g) * 1.5 + 0.3).cuda()
    dy = bf16_round(torch.randn(N, C, D, H, W, generator=g)).cuda()
    acc = bf16_round(torch.randn(N, C, D, H, W, generator=g)).cuda() if with_acc else None
    gamma = (torch.randn(C, generator=g) * 0.5 + 1.0).cuda()
    beta = (torch.randn(C, generator=g) * 0.2).cuda()
    xc, dyc = to_cl(x), to_cl(dy)
    stats = torch.empty((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
    ops.groupnorm_stats(xc, stats, N, S, C, G)
    dxc = torch.empty_like(xc)
    dgamma = torch.full((C,), 0.25, device="cuda")
    dbeta = torch.full((C,), -0.5, device="cuda")
    cs = torch.zeros((N, C + 8), device="cuda")              # fused column sums of the stored dx (bias / embedding gradients)
    ops.groupnorm_bwd(xc, dyc, stats, gamma, beta, dxc, dgamma, dbeta, N, S, C, G, 1e-5, silu,
                      acc=to_cl(acc) if with_acc else None, colsum=cs)
    torch.cuda.synchronize()
    want_cs = dxc.float().view(N, S, -1)[:, :, :C].sum(1)
    assert float((cs[:, :C] - want_cs).abs().max()) <= 1e-3 * float(want_cs.abs().max()) + 1e-3
    assert float(cs[:, C:].abs().max()) == 0.0
    per, tot = torch.ones((N, C + 4), device="cuda"), torch.full((C,), 2.0, device="cuda")
    ops.colsum_scatter(cs, N, C, out_sample=per[:, 4:], out_total=tot)
    assert float((per[:, 4:] - 1.0 - cs[:, :C]).abs().max()) <= 1e-5 * float(cs.abs().max()) + 1e-6
    assert float((tot - 2.0 - cs[:, :C].sum(0)).abs().max()) <= 1e-5 * float(cs.abs().max()) + 1e-5
    assert float((per[:, :4] - 1.0).abs().max()) == 0.0
    dx_plain = torch.empty_like(xc)                          # the variant without column sums stores the same dx
    ops.groupnorm_bwd(xc, dyc, stats, gamma, beta, dx_plain, torch.zeros_like(dgamma), torch.zeros_like(dbeta), N, S, C, G,
                      1e-5, silu, acc=to_cl(acc) if with_acc else None)
    assert torch.equal(dx_plain, dxc)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr, G, gr, br, 1e-5)
    if silu:
        y = F.silu(y)
    y.backward(dy)
    ref = xr.grad + (acc if with_acc else 0.0)
    got = from_cl(dxc, (N, C, D, H, W))
    err = float((got - ref).abs().max())
    assert err <= 1e-2 * float(ref.abs().max()) + 1e-3, err
    for got_p, ref_p, base in ((dgamma, gr.grad, 0.25), (dbeta, br.grad, -0.5)):
        e = float((got_p - base - ref_p).abs().max())
        assert e <= 2e-3 * float(ref_p.abs().max()) + 1e-3, e


def test_colsum_cl():
    from fcwdm import ops
    from gpu_util import bf16_round, to_cl
    N, C, D, H, W = 2, 64, 5, 6, 8
    x = bf16_round(torch.randn(N, C, D, H, W, generator=torch.Generator().manual_seed(4))).cuda()
    per = torch.zeros((N, 72), device="cuda")
    tot = torch.ones((C,), device="cuda")
    ops.colsum_cl(to_cl(x), N, D * H * W, C, out_sample=per[:, 4:68], out_total=tot)
    torch.cuda.synchronize()
    ref = x.sum(dim=(2, 3, 4))
    assert float((per[:, 4:68] - ref).abs().max()) <= 1e-3
    assert float((tot - 1.0 - ref.sum(0)).abs().max()) <= 2e-3
    assert float(per[:, :4].abs().max()) == 0.0 and float(per[:, 68:].abs().max()) == 0.0


@pytest.mark.parametrize("C", [64, 128])
def test_haar_cl_adjoints(C):
    """<DWT(x), g> == <x, DWT^T(g)> for the channels-last kernels, with the scales, the LLL-only variant, the fan-in
    adds and the high-band accumulation."""
    from fcwdm import ops
    from gpu_util import bf16_round
    N, D, H, W = 1, 4, 6, 8
    S, s = D * H * W, D * H * W // 8
    g = torch.Generator().manual_seed(5)
    rnd = lambda *sh: bf16_round(torch.randn(*sh, generator=g)).cuda().to(torch.bfloat16)
    x = rnd(N * S, C)
    # forward DWT: lll = LLL/3, hi = 7 bands * 0.5
    lll = torch.empty((N * s, C), dtype=torch.bfloat16, device="cuda")
    hi = torch.empty((7, N * s, C), dtype=torch.bfloat16, device="cuda")
    ops.dwt3d_cl(x, (N, D, H, W), C, lll, hi, lll_scale=1.0 / 3.0, hi_scale=0.5)
    g_lll, g_hi, acc = rnd(N * s, C), rnd(7, N * s, C), rnd(N * S, C)
    dx = torch.empty_like(x)
    ops.dwt3d_cl_bwd(g_lll, g_hi, (N, D, H, W), C, dx, acc=acc, lll_scale=1.0 / 3.0, hi_scale=0.5)
    lhs = float((lll.double() * g_lll.double()).sum() + (hi.double() * g_hi.double()).sum())
    rhs = float((x.double() * (dx.double() - acc.double())).sum())
    scale = float(x.double().norm() * (dx.double() - acc.double()).norm())
    assert abs(lhs - rhs) <= 1e-2 * scale, (lhs, rhs, scale)
    dx0 = torch.empty_like(x)
    ops.dwt3d_cl_bwd(g_lll, None, (N, D, H, W), C, dx0, lll_scale=1.0 / 3.0)          # LLL-only (x_upd branch)
    rhs0 = float((x.double() * dx0.double()).sum())
    lhs0 = float((lll.double() * g_lll.double()).sum())
    assert abs(lhs0 - rhs0) <= 1e-2 * float(x.double().norm() * dx0.double().norm()), (lhs0, rhs0)
    # forward IDWT: y = IDWT(3*lll, hi)
    y = torch.empty((N * S, C), dtype=torch.bfloat16, device="cuda")
    ops.idwt3d_cl(lll, hi, (N, D, H, W), C, y, lll_scale=3.0)
    gy = rnd(N * S, C)
    d_lll = torch.empty_like(lll)
    d_hi = g_hi.clone()
    lacc = rnd(N * s, C)
    ops.idwt3d_cl_bwd(gy, (N, D, H, W), C, d_lll, d_hi, lll_acc=lacc, hi_accumulate=True, lll_scale=3.0)
    lhs = float((y.double() * gy.double()).sum())
    rhs = float((lll.double() * (d_lll.double() - lacc.double())).sum() +
                (hi.double() * (d_hi.double() - g_hi.double())).sum())
    scale = float(y.double().norm() * gy.double().norm())
    assert abs(lhs - rhs) <= 1e-2 * scale, (lhs, rhs, scale)


def test_linear_bwd_and_adamw():
    from fcwdm import ops
    g = torch.Generator().manual_seed(6)
    N, K, M = 2, 256, 192
    x = torch.randn(N, K, generator=g).cuda()
    Wt = (torch.randn(M, K, generator=g) / 16).cuda()
    dy_full = torch.randn(N, M + 64, generator=g).cuda()
    dy = dy_full[:, 32:32 + M]                                   # strided slice, as the engine passes it
    dx = torch.full((N, K), 0.5, device="cuda")
    dW = torch.zeros(M, K, device="cuda")
    db = torch.zeros(M, device="cuda")
    ops.linear_bwd(x, Wt, dy, dx=dx, dW=dW, db=db, act_in=1, accumulate_dx=True)
    torch.cuda.synchronize()
    xr, wr = x.clone().requires_grad_(True), Wt.clone().requires_grad_(True)
    br = torch.zeros(M, device="cuda", requires_grad=True)
    F.linear(F.silu(xr), wr, br).backward(dy)
    assert float((dx - 0.5 - xr.grad).abs().max()) <= 1e-4 * float(xr.grad.abs().max()) + 1e-5
    assert float((dW - wr.grad).abs().max()) <= 1e-4 * float(wr.grad.abs().max()) + 1e-5
    assert float((db - br.grad).abs().max()) <= 1e-5
    # AdamW: three steps against torch.optim.AdamW on the same flat parameter
    n = 1003
    p0 = torch.randn(n, generator=g).cuda()
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g).cuda()
        ops.adamw(p, gr * 2.0, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, grad_scale=0.5)
        pr.grad = gr.clone()
        opt.step()
    torch.cuda.synchronize()
    assert float((p - pr.detach()).abs().max()) <= 2e-6
