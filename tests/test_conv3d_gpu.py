"""GPU parity of the tcgen05 implicit-GEMM conv3d (fcwdm_conv3d_fwd through the C-ABI) against torch's conv3d
in fp32 on the same bf16-rounded operands.  Tolerance: outputs are stored in bf16 (rel 2^-8) and accumulated in
fp32 over K <= 27*1024 terms: |err| <= 1e-2 * max|ref| + 1e-3 (stated bf16 tolerance)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def run_conv(N, D, H, W, cin, cout, k, use_bias=True, use_cb=False, use_res=False, seed=0, wfill=None):
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(N, cin, D, H, W, generator=g)).cuda()
    w = bf16_round(torch.randn(cout, cin, k, k, k, generator=g) / np.sqrt(cin * k ** 3))
    if wfill is not None:
        w = wfill(w)
    w = w.cuda()
    bias = torch.randn(cout, generator=g).cuda() if use_bias else None
    cb = torch.randn(N, cout, generator=g).cuda() if use_cb else None
    res = bf16_round(torch.randn(N, cout, D, H, W, generator=g)).cuda() if use_res else None
    xc = to_cl(x)
    wp = ops.conv3d_pack_weights(w)
    yc = torch.zeros((N * D * H * W, (cout + 63) // 64 * 64), dtype=torch.bfloat16, device="cuda")
    rc = to_cl(res) if use_res else None
    ops.conv3d_cl(xc, wp, bias, yc, (N, D, H, W), cin, cout, k, chan_bias=cb, residual=rc)
    torch.cuda.synchronize()
    got = from_cl(yc, (N, cout, D, H, W))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = F.conv3d(x, w, bias, padding=k // 2)
    if use_cb:
        ref = ref + cb[:, :, None, None, None]
    if use_res:
        ref = ref + res
    assert float(yc[:, cout:].abs().max()) == 0.0 if yc.shape[1] > cout else True
    return got, ref


def check(got, ref):
    err = float((got - ref).abs().max())
    tol = 1e-2 * float(ref.abs().max()) + 1e-3
    assert err <= tol, (err, tol)


CASES = [
    # N, D, H, W, cin, cout, k
    (1, 4, 16, 8, 64, 64, 3),       # exactly one (64,4,3)/(64,x) tile
    (1, 5, 18, 10, 64, 64, 3),      # partial tiles in every dim
    (2, 3, 7, 5, 32, 32, 3),        # padded channels, tiny dims (the 7x7x5 bottleneck shape family)
    (1, 8, 16, 16, 128, 128, 3),    # N_TILE = 128, two channel blocks
    (1, 4, 20, 12, 256, 64, 3),     # four channel blocks (WaveletDownsample 256 -> 64)
    (1, 4, 16, 8, 64, 8, 3),        # output conv 64 -> 8 (N_TILE = 16)
    (1, 6, 12, 24, 64, 128, 1),     # 1x1x1 skip conv
    (1, 2, 14, 10, 256, 128, 1),
    (1, 14, 14, 10, 128, 256, 3),   # 256 output channels
    (1, 5, 7, 7, 256, 256, 3),      # bottleneck: one tile per CTA -> split-K over a 4-CTA cluster
    (2, 10, 14, 14, 256, 256, 3),   # two samples, split-K 2 or 4
    (1, 5, 7, 7, 1024, 256, 3),     # input-pyramid conv at the bottleneck: 16 channel blocks over 8 CTAs
    (1, 5, 7, 7, 512, 64, 3),       # N_TILE = 64 single n-tile, 8 channel blocks
]


@pytest.mark.parametrize("case", CASES)
def test_conv3d_vs_torch(case):
    got, ref = run_conv(*case)
    check(got, ref)


def test_conv3d_epilogue_fusions():
    got, ref = run_conv(2, 4, 16, 8, 64, 64, 3, use_bias=True, use_cb=True, use_res=True, seed=3)
    check(got, ref)
    got, ref = run_conv(1, 3, 9, 9, 64, 128, 1, use_bias=False, use_cb=False, use_res=True, seed=4)
    check(got, ref)
    # split-K cluster path (bottleneck shapes): the leader's epilogue carries bias, embedding and residual
    got, ref = run_conv(2, 5, 7, 7, 256, 256, 3, use_bias=True, use_cb=True, use_res=True, seed=5)
    check(got, ref)


def test_conv3d_every_tap():
    """One non-zero filter tap at a time: pins the halo-offset arithmetic of the UMMA descriptors."""
    for tap in range(27):
        def only(w, tap=tap):
            m = torch.zeros_like(w)
            m.view(w.shape[0], w.shape[1], 27)[:, :, tap] = 1
            return w * m
        got, ref = run_conv(1, 4, 16, 8, 64, 64, 3, use_bias=False, seed=tap, wfill=only)
        err = float((got - ref).abs().max())
        assert err <= 1e-2 * float(ref.abs().max()) + 1e-3, (tap, err)


def test_conv3d_full_resolution_shape():
    """The dominant conv of CFG-W4 (64 -> 64 at 112x112x80), checked on sampled voxels incl. all borders."""
    got, ref = run_conv(1, 112, 112, 80, 64, 64, 3, seed=9)
    check(got, ref)


@pytest.mark.parametrize("case", [(2, 5, 18, 10, 64, 64, 3, 32), (1, 4, 16, 16, 128, 128, 3, 32),
                                  (1, 6, 10, 12, 64, 256, 1, 32), (2, 3, 7, 5, 32, 32, 3, 32),
                                  (2, 5, 7, 7, 256, 256, 3, 32),             # split-K: statistics from the leader only
                                  (1, 4, 16, 16, 64, 128, 3, 8), (1, 4, 16, 16, 64, 128, 3, 4),      # 16 / 32 / 64 channels per
                                  (1, 4, 16, 16, 64, 128, 3, 2)])                                    # group: one reduce per 32 columns
def test_conv3d_fused_groupnorm_statistics(case):
    """The epilogue's fused (sum, sumsq) per (n, group) of the stored bf16 output == statistics of that tensor."""
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    N, D, H, W, cin, cout, k, G = case
    g = torch.Generator().manual_seed(7)
    x = bf16_round(torch.randn(N, cin, D, H, W, generator=g)).cuda()
    w = (torch.randn(cout, cin, k, k, k, generator=g) / np.sqrt(cin * k ** 3)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    xc, wp = to_cl(x), ops.conv3d_pack_weights(w)
    yc = torch.zeros((N * D * H * W, (cout + 63) // 64 * 64), dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
    ops.conv3d_cl(xc, wp, bias, yc, (N, D, H, W), cin, cout, k, gn_stats=stats, gn_groups=G)
    y = from_cl(yc, (N, cout, D, H, W)).double().reshape(N, G, -1)
    got = stats.sum(dim=1)
    np.testing.assert_allclose(got[..., 0].cpu().numpy(), y.sum(-1).cpu().numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(got[..., 1].cpu().numpy(), (y * y).sum(-1).cpu().numpy(), rtol=1e-4, atol=1e-2)


def run_pair(N, D, H, W, cin, cout, use_bias=True, use_cb=False, use_res=False, seed=0, stats_groups=0):
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(N, cin, D, H, W, generator=g)).cuda()
    w = bf16_round(torch.randn(cout, cin, 3, 3, 3, generator=g) / np.sqrt(cin * 27)).cuda()
    bias = torch.randn(cout, generator=g).cuda() if use_bias else None
    cb = torch.randn(N, cout, generator=g).cuda() if use_cb else None
    res = bf16_round(torch.randn(N, cout, D, H, W, generator=g)).cuda() if use_res else None
    xc = to_cl(x)
    wp = ops.conv3d_pair_pack_weights(w)
    yc = torch.zeros((N * D * H * W, (cout + 63) // 64 * 64), dtype=torch.bfloat16, device="cuda")
    rc = to_cl(res) if use_res else None
    stats = torch.zeros((N, ops.GN_STAT_REPLICAS, stats_groups, 2), dtype=torch.float64, device="cuda") if stats_groups else None
    ops.conv3d_pair_cl(xc, wp, bias, yc, (N, D, H, W), cin, cout, chan_bias=cb, residual=rc, gn_stats=stats,
                       gn_groups=stats_groups)
    torch.cuda.synchronize()
    got = from_cl(yc, (N, cout, D, H, W))
    ref = F.conv3d(x, w, bias, padding=1)
    if use_cb:
        ref = ref + cb[:, :, None, None, None]
    if use_res:
        ref = ref + res
    return got, ref, stats


PAIR_CASES = [
    (1, 4, 16, 16, 64, 64),        # one pair along W, one segment
    (1, 9, 32, 8, 64, 64),         # pair along H
    (2, 7, 18, 10, 64, 64),        # partial tiles, batch 2
    (1, 23, 40, 24, 32, 64),       # several depth segments, C_in padded to 64 (stem conv)
    (1, 12, 16, 16, 64, 8),        # output conv 64 -> 8 (N_TILE = 16)
    (1, 5, 7, 5, 64, 32),          # tiny spatial dims, C_out 32
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv3d_pair_vs_torch(case):
    got, ref, _ = run_pair(*case)
    check(got, ref)


def test_conv3d_pair_epilogue_and_stats():
    got, ref, stats = run_pair(2, 6, 20, 16, 64, 64, use_cb=True, use_res=True, seed=5, stats_groups=32)
    check(got, ref)
    y = got.double().reshape(2, 32, -1)
    s = stats.sum(dim=1)
    np.testing.assert_allclose(s[..., 0].cpu().numpy(), y.sum(-1).cpu().numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(s[..., 1].cpu().numpy(), (y * y).sum(-1).cpu().numpy(), rtol=1e-4, atol=1e-2)


def test_conv3d_pair_full_resolution():
    got, ref, _ = run_pair(1, 112, 112, 80, 64, 64, seed=9)
    check(got, ref)


@pytest.mark.parametrize("case", [(1, 6, 20, 16, 64, 64), (2, 9, 18, 10, 32, 64), (1, 23, 40, 24, 64, 8)])
def test_conv3d_pair_fused_input_groupnorm(case):
    """conv(SiLU(GroupNorm(x))) with the normalisation applied inside the conv's operand producers."""
    from fcwdm import native, ops
    from gpu_util import bf16_round, from_cl, to_cl
    N, D, H, W, cin, cout = case
    G = 32
    g = torch.Generator().manual_seed(11)
    x = bf16_round(torch.randn(N, cin, D, H, W, generator=g) * 1.5 + 0.3).cuda()
    w = bf16_round(torch.randn(cout, cin, 3, 3, 3, generator=g) / np.sqrt(cin * 27)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    gamma = (1 + 0.2 * torch.randn(cin, generator=g)).cuda()
    beta = (0.2 * torch.randn(cin, generator=g)).cuda()
    xc = to_cl(x)
    S = D * H * W
    stats = torch.empty((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
    with ops._on(x.device) as st:
        native.call("fcwdm_groupnorm_stats", ops._ptr(xc), xc.stride(0), ops._ptr(stats), N, S, cin, G, st)
    wp = ops.conv3d_pair_pack_weights(w)
    yc = torch.zeros((N * S, 64), dtype=torch.bfloat16, device="cuda")
    ops.conv3d_pair_cl(xc, wp, bias, yc, (N, D, H, W), cin, cout, gn_in=(stats, gamma, beta, G, 1e-5))
    torch.cuda.synchronize()
    got = from_cl(yc, (N, cout, D, H, W))
    a = bf16_round(F.silu(F.group_norm(x, G, gamma, beta, 1e-5)))
    ref = F.conv3d(a, w, bias, padding=1)
    err = float((got - ref).abs().max())
    assert err <= 2e-2 * float(ref.abs().max()) + 2e-3, err


GN_IN_CASES = [
    # N, D, H, W, cin, cout, groups
    (1, 4, 16, 8, 128, 128, 32),     # two channel blocks, one tile per plane pair
    (2, 5, 18, 10, 64, 128, 32),     # batch 2 (per-sample statistics), ragged tiles
    (1, 3, 7, 5, 256, 256, 32),      # bottleneck shape family, four channel blocks
    (1, 6, 20, 12, 192, 64, 32),     # 6 channels per group (plain U-Net concat width), N_TILE = 64
    (1, 8, 16, 16, 128, 64, 8),
]


@pytest.mark.parametrize("case", GN_IN_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3d_fused_input_groupnorm(case):
    """fcwdm_conv3d_gn_fwd == conv(SiLU(GroupNorm(x))) + bias + per-sample channel bias + residual: the halo planes are
    normalised in shared memory by the kernel; zero padding applies to the ACTIVATED tensor."""
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    N, D, H, W, cin, cout, G = case
    g = torch.Generator().manual_seed(7)
    x = bf16_round(torch.randn(N, cin, D, H, W, generator=g) * 1.3 + 0.4).cuda()
    w = bf16_round(torch.randn(cout, cin, 3, 3, 3, generator=g) / np.sqrt(cin * 27)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    cb = torch.randn(N, cout, generator=g).cuda()
    res = bf16_round(torch.randn(N, cout, D, H, W, generator=g)).cuda()
    gamma = (torch.randn(cin, generator=g) * 0.3 + 1.0).cuda()
    beta = (torch.randn(cin, generator=g) * 0.2).cuda()
    xc = to_cl(x)
    S = D * H * W
    stats = torch.empty((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
    ops.groupnorm_stats(xc, stats, N, S, cin, G)
    yc = torch.zeros((N * S, (cout + 63) // 64 * 64), dtype=torch.bfloat16, device="cuda")
    ops.conv3d_cl(xc, ops.conv3d_pack_weights(w), bias, yc, (N, D, H, W), cin, cout, 3, chan_bias=cb, residual=to_cl(res),
                  gn_in=(stats, gamma, beta, G, 1e-5))
    torch.cuda.synchronize()
    got = from_cl(yc, (N, cout, D, H, W))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    a = bf16_round(F.silu(F.group_norm(x, G, gamma, beta, 1e-5)))       # the kernel feeds bf16 operands to the MMA
    ref = F.conv3d(a, w, bias, padding=1) + cb[:, :, None, None, None] + res
    check(got, ref)


def test_pack_all_matches_the_per_conv_packers():
    """fcwdm_conv3d_pack_all (one launch re-packing every conv of a model, forward and data-gradient forms, staged
    through shared memory) against the single-conv packing entry points, bit for bit, padding included."""
    import torch
    from fcwdm import native, ops
    g = torch.Generator().manual_seed(3)
    shapes = [(64, 64, 3), (64, 32, 3), (8, 64, 3), (128, 64, 3), (64, 192, 3), (256, 384, 3), (128, 64, 1), (64, 192, 1),
              (24, 40, 3), (16, 8, 1)]
    jobs, want, keep = [], [], []
    for co, ci, k in shapes:
        w = (torch.randn((co, ci, k, k, k), generator=g) * 0.1).cuda()
        keep.append(w)
        forms = [(co, ci, 0, w)]
        if ci % 8 == 0:
            forms.append((ci, co, 1, ops.conv3d_transpose_flip_weights(w)))
        for O, I, transposed, w_form in forms:
            pair = ops.conv3d_pair_supported(I, O, k)
            ref = ops.conv3d_pair_pack_weights(w_form) if pair else ops.conv3d_pack_weights(w_form)
            dst = torch.full_like(ref, float("nan"))
            want.append((ref, dst, (co, ci, k, transposed, pair)))
            jobs.append([w.data_ptr(), dst.data_ptr(), O, I, k ** 3, int(pair), transposed, ref.numel()])
    table = torch.tensor(jobs, dtype=torch.int64).cuda()
    with ops._on(table.device) as st:
        native.call("fcwdm_conv3d_pack_all", ops._ptr(table), table.shape[0], max(j[7] for j in jobs), st)
    torch.cuda.synchronize()
    for ref, dst, tag in want:
        assert torch.equal(dst.view(torch.int16), ref.view(torch.int16)), tag
