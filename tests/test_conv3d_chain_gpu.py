"""GPU parity of the persistent low-resolution conv chain (fcwdm_conv3d_chain, csrc/conv3d_chain.cu): a RUN of ResBlock
convolutions in one launch -- grid barriers between layers, split-K with a reduce-scatter over distributed shared memory,
fused input GroupNorm + SiLU from statistics accumulated by the previous layer inside the same launch -- against torch
fp32 conv3d / group_norm on the same bf16-rounded operands, layer by layer (each layer's reference is fed the CHAIN's own
stored input, so tolerances do not compound), and against the per-layer kernels (fcwdm_conv3d_fwd / _gn_fwd).

Tolerance per layer: |err| <= 1e-2 * max|ref| + 1e-3 (bf16 output, fp32 accumulation); fused statistics rtol 1e-4."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

G = 32
EPS = 1e-5


def _mk(gen, *shape, scale=1.0):
    from gpu_util import bf16_round
    return bf16_round(torch.randn(*shape, generator=gen) * scale).cuda()


def _ref_layer(x, w, bias, cb, res, gn):
    """fp32 reference of one chain layer.  x: planar (N,C,D,H,W) = what the chain read; gn = (gamma, beta) or None."""
    from gpu_util import bf16_round
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    a = x
    if gn is not None:
        a = bf16_round(F.silu(F.group_norm(x, G, gn[0], gn[1], EPS)))       # the kernel stores the activated operand in bf16
    y = F.conv3d(a, w, bias, padding=1)
    if cb is not None:
        y = y + cb[:, :, None, None, None]
    if res is not None:
        y = y + res
    return y


def run_chain(N, dims, widths, seed, with_res=True):
    """A ResBlock-like run: layer 0 plain (Cin = widths[0]), then alternating [GN_IN + emb + stats] / [GN_IN + residual +
    stats] layers through widths[1:].  Returns per-layer (got planar, ref planar, fused statistics, y buffer)."""
    from fcwdm import ops
    from gpu_util import from_cl, to_cl
    D, H, W = dims
    S = D * H * W
    gen = torch.Generator().manual_seed(seed)
    x0 = _mk(gen, N, widths[0], D, H, W)
    cur_cl, cur_c = to_cl(x0), widths[0]
    layers, keep, meta = [], [], []
    prev_stats = None
    for li, cout in enumerate(widths[1:]):
        cin = cur_c
        w = _mk(gen, cout, cin, 3, 3, 3, scale=1.0 / np.sqrt(cin * 27))
        bias = torch.randn(cout, generator=gen).cuda()
        wp = ops.conv3d_pack_weights(w)
        y = torch.zeros((N * S, cout), dtype=torch.bfloat16, device="cuda")
        stats = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
        cb = res = gn = gn_in = None
        if li > 0 and cin <= 256:
            gamma, beta = (torch.rand(cin, generator=gen) + 0.5).cuda(), (torch.randn(cin, generator=gen) * 0.2).cuda()
            gn, gn_in = (gamma, beta), (prev_stats, gamma, beta, G, EPS)
        if li % 2 == 1:
            cb = torch.randn(N, cout, generator=gen).cuda()
        elif li > 0 and with_res:
            res = _mk(gen, N, cout, D, H, W)
        res_cl = to_cl(res) if res is not None else None
        L, ka = ops.conv3d_chain_layer(cur_cl, wp, bias, y, (N, D, H, W), cin, cout, chan_bias=cb, residual=res_cl,
                                       gn_stats=stats, gn_groups=G, gn_in=gn_in)
        layers.append(L)
        keep.append(ka)
        meta.append((cur_cl, cin, w, bias, cb, res, gn, y, cout, stats))
        cur_cl, cur_c, prev_stats = y, cout, stats
    counter = torch.zeros(2, dtype=torch.int64, device="cuda")
    ops.conv3d_chain(layers, counter)
    torch.cuda.synchronize()
    assert int(counter[0]) > 0 or len(layers) == 1
    out = []
    for (x_cl, cin, w, bias, cb, res, gn, y, cout, stats) in meta:
        x = from_cl(x_cl, (N, cin, D, H, W))
        ref = _ref_layer(x, w, bias, cb, res, gn)
        got = from_cl(y, (N, cout, D, H, W))
        out.append((got, ref, stats, y))
    return out


def check_layers(out, N):
    for li, (got, ref, stats, y) in enumerate(out):
        err = float((got - ref).abs().max())
        tol = 1e-2 * float(ref.abs().max()) + 1e-3
        assert err <= tol, (li, err, tol)
        # fused statistics of the STORED bf16 output
        yy = got.double().reshape(N, G, -1)
        s = stats.sum(dim=1)
        np.testing.assert_allclose(s[..., 0].cpu().numpy(), yy.sum(-1).cpu().numpy(), rtol=1e-4, atol=1e-2)
        np.testing.assert_allclose(s[..., 1].cpu().numpy(), (yy * yy).sum(-1).cpu().numpy(), rtol=1e-4, atol=1e-2)


CHAINS = [
    # N, (D, H, W), channel widths along the run
    (1, (5, 7, 7), (256, 256, 256, 256)),               # bottleneck: 10 tiles, split 4 on every layer
    (1, (10, 14, 14), (128, 256, 256, 128, 128)),        # level 3: 20/40 tiles, split 2 and 4, both C_out
    (1, (20, 28, 28), (128, 128, 128)),                  # level 2: 160 tiles > grid: one full wave + a shared remainder
    (1, (5, 7, 7), (1024, 256, 256)),                    # input-pyramid conv at the bottleneck (16 channel blocks, no GroupNorm in)
    (2, (4, 9, 11), (64, 128, 128, 256)),                # two samples, ragged tiles, C_in = 64 (one channel block: split 1)
    (1, (10, 14, 14), (1024, 128)),                      # single-layer chain (no barrier)
    (1, (5, 7, 7), (256, 512, 1024)),                    # 16 and 32 channels per group in the fused statistics
]


@pytest.mark.parametrize("case", CHAINS)
def test_chain_against_torch(case):
    N, dims, widths = case
    out = run_chain(N, dims, widths, seed=len(widths) + dims[0])
    check_layers(out, N)


def test_chain_equals_per_layer_kernels():
    """The same run through fcwdm_conv3d_fwd / fcwdm_conv3d_gn_fwd, one launch per layer: same operands, same fused
    statistics; the only difference is the order of the fp32 partial sums (split-K grouping)."""
    from fcwdm import ops
    from gpu_util import to_cl
    N, (D, H, W), widths = 1, (10, 14, 14), (128, 256, 256, 256)
    S = D * H * W
    gen = torch.Generator().manual_seed(3)
    x0 = _mk(gen, N, widths[0], D, H, W)
    specs = []
    for li, cout in enumerate(widths[1:]):
        cin = widths[li]
        w = _mk(gen, cout, cin, 3, 3, 3, scale=1.0 / np.sqrt(cin * 27))
        specs.append((ops.conv3d_pack_weights(w), torch.randn(cout, generator=gen).cuda(),
                      (torch.rand(cin, generator=gen) + 0.5).cuda(), (torch.randn(cin, generator=gen) * 0.2).cuda(),
                      torch.randn(N, cout, generator=gen).cuda(), cin, cout))

    def run(chain):
        cur = to_cl(x0)
        prev = None
        layers, ys = [], []
        for li, (wp, bias, gamma, beta, cb, cin, cout) in enumerate(specs):
            y = torch.zeros((N * S, cout), dtype=torch.bfloat16, device="cuda")
            st = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
            gn_in = (prev, gamma, beta, G, EPS) if li > 0 else None
            res = ys[-1][0] if (li == 2) else None                   # a residual produced INSIDE the run (layer 1's output)
            if chain:
                layers.append(ops.conv3d_chain_layer(cur, wp, bias, y, (N, D, H, W), cin, cout, chan_bias=cb, residual=res,
                                                     gn_stats=st, gn_groups=G, gn_in=gn_in))
            else:
                ops.conv3d_cl(cur, wp, bias, y, (N, D, H, W), cin, cout, 3, chan_bias=cb, residual=res, gn_stats=st,
                              gn_groups=G, gn_in=gn_in)
            ys.append((y, st))
            cur, prev = y, st
        if chain:
            ops.conv3d_chain([l for l, _ in layers], torch.zeros(2, dtype=torch.int64, device="cuda"))
        torch.cuda.synchronize()
        return ys

    a, b = run(True), run(False)
    for li, ((ya, sa), (yb, sb)) in enumerate(zip(a, b)):
        fa, fb = ya.float(), yb.float()
        scale = float(fb.abs().max())
        # layer 0 sees identical inputs: differences are fp32 summation order only (<= 1 bf16 ulp of the output)
        assert float((fa - fb).abs().max()) <= (2.0 ** -7 if li == 0 else 3e-2) * scale, li
        np.testing.assert_allclose(sa.sum(1).cpu().numpy(), sb.sum(1).cpu().numpy(), rtol=(1e-3 if li == 0 else 3e-2), atol=1.0)


def test_engine_uses_the_chain_and_matches_the_per_layer_plan(monkeypatch):
    """WavUNetModel.forward with the chain (default) and with FCWDM_NO_CHAIN=1: same network output within bf16 noise, and
    the chained plan launches fewer kernels."""
    from fcwdm import native
    from guided_diffusion.wunet import WavUNetModel
    from oracle import wunet as ow
    cfg = dict(image_size=64, in_channels=32, model_channels=64, out_channels=8, num_res_blocks=2,
               attention_resolutions=(), channel_mult=(1, 2, 2, 4), dims=3, num_groups=32,
               bottleneck_attention=False, resblock_updown=True, use_freq=True)
    outs, launches = {}, {}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 32, 32, 32, 16, generator=g)
    t = torch.tensor([500])
    for mode in ("1", "0"):
        monkeypatch.setenv("FCWDM_NO_CHAIN", mode)
        m = WavUNetModel(**cfg)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = ow.tie_output_blocks(ow.seeded_state_dict(shapes, seed=0, std=0.02), 4)
        m.load_state_dict(sd, strict=True)
        m.to("cuda").eval()
        with torch.no_grad():
            m(x.cuda(), t.cuda())                                  # packing, lazy init
            n0 = native.launch_count
            outs[mode] = m(x.cuda(), t.cuda()).cpu()
            launches[mode] = native.launch_count - n0
        if mode == "0":
            assert m.engine().chain_launches > 0
    ref = ow.wunet_forward(sd, x, t, model_channels=64, channel_mult=(1, 2, 2, 4))
    r_chain = float((outs["0"] - ref).norm() / ref.norm())
    r_plain = float((outs["1"] - ref).norm() / ref.norm())
    print(f"chain {r_chain:.3e} vs per-layer {r_plain:.3e} rel-L2 against the oracle; launches {launches['0']} vs {launches['1']}")
    assert r_chain <= 3e-2 and r_plain <= 3e-2
    assert launches["0"] < launches["1"]


@pytest.mark.parametrize("C,dims,N", [(128, (20, 28, 28), 1), (256, (10, 14, 14), 1), (64, (4, 6, 10), 2), (128, (2, 14, 14), 3)])
def test_chain_wavelet_ops_equal_the_standalone_kernels(C, dims, N):
    """In-chain Haar DWT / IDWT (the ResBlock re-sampling between the convs of a run) == fcwdm_dwt3d_cl / fcwdm_idwt3d_cl
    bit for bit (same butterfly, same scaling), and their fused GroupNorm statistics == the statistics of what they stored."""
    from fcwdm import ops
    D, H, W = dims
    S, s2 = D * H * W, (D // 2) * (H // 2) * (W // 2)
    gen = torch.Generator().manual_seed(C + D)
    x = torch.randn(N * S, C, generator=gen).cuda().to(torch.bfloat16)
    emb = torch.randn(N, C, generator=gen).cuda()
    # --- reference: the standalone kernels
    lll_ref = torch.zeros((N * s2, C), dtype=torch.bfloat16, device="cuda")
    hi_ref = torch.zeros((7, N * s2, C), dtype=torch.bfloat16, device="cuda")
    ops.dwt3d_cl(x, (N, D, H, W), C, lll_ref, hi_ref, lll_bias=emb, lll_scale=1.0 / 3.0)
    only_ref = torch.zeros((N * s2, C), dtype=torch.bfloat16, device="cuda")
    ops.dwt3d_cl(x, (N, D, H, W), C, only_ref, None, lll_scale=1.0 / 3.0)
    cat_ref = torch.zeros((N * s2, 8 * C), dtype=torch.bfloat16, device="cuda")
    ops.dwt3d_cl(x, (N, D, H, W), C, cat_ref[:, :C], cat_ref[:, C:], lll_scale=1.0 / 3.0, hi_scale=1.0 / 3.0, hi_sb=C)
    y_ref = torch.zeros((N * S, C), dtype=torch.bfloat16, device="cuda")
    ops.idwt3d_cl(lll_ref, hi_ref, (N, D, H, W), C, y_ref, bias=emb, lll_scale=3.0)
    # --- the same four ops as ONE chain launch (the IDWT consumes what the first DWT of the same launch produced)
    lll = torch.zeros_like(lll_ref)
    hi = torch.zeros_like(hi_ref)
    only = torch.zeros_like(only_ref)
    cat = torch.zeros_like(cat_ref)
    y = torch.zeros_like(y_ref)
    st_l = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
    st_y = torch.zeros_like(st_l)
    layers = [ops.chain_dwt_op(x, (N, D, H, W), C, lll, hi, lll_bias=emb, lll_scale=1.0 / 3.0, gn_stats=st_l, gn_groups=G),
              ops.chain_dwt_op(x, (N, D, H, W), C, only, None, lll_scale=1.0 / 3.0),
              ops.chain_dwt_op(x, (N, D, H, W), C, cat[:, :C], cat[:, C:], lll_scale=1.0 / 3.0, hi_scale=1.0 / 3.0, hi_sb=C),
              ops.chain_idwt_op(lll, hi, (N, D, H, W), C, y, bias=emb, lll_scale=3.0, gn_stats=st_y, gn_groups=G)]
    ops.conv3d_chain([l for l, _ in layers], torch.zeros(2, dtype=torch.int64, device="cuda"))
    torch.cuda.synchronize()
    for name, a, b in (("lll", lll, lll_ref), ("hi", hi, hi_ref), ("lll only", only, only_ref), ("concat", cat, cat_ref), ("idwt", y, y_ref)):
        assert torch.equal(a, b), name
    for st, t, rows in ((st_l, lll, s2), (st_y, y, S)):
        v = t.double().reshape(N, rows, G, C // G)
        s = st.sum(dim=1)
        np.testing.assert_allclose(s[..., 0].cpu().numpy(), v.sum(dim=(1, 3)).cpu().numpy(), rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(s[..., 1].cpu().numpy(), (v * v).sum(dim=(1, 3)).cpu().numpy(), rtol=1e-5, atol=1e-3)


def test_chain_conv_dwt_conv():
    """conv -> in-chain DWT (LLL / 3 + embedding, fused statistics) -> conv with fused input GroupNorm from those
    statistics, in one launch, against torch fp32 step by step."""
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    N, (D, H, W), C = 1, (10, 14, 14), 128
    S, s2 = D * H * W, (D // 2) * (H // 2) * (W // 2)
    gen = torch.Generator().manual_seed(11)
    x0 = _mk(gen, N, C, D, H, W)
    w1, w2 = (_mk(gen, C, C, 3, 3, 3, scale=1.0 / np.sqrt(C * 27)) for _ in range(2))
    b1, b2 = torch.randn(C, generator=gen).cuda(), torch.randn(C, generator=gen).cuda()
    emb = torch.randn(N, C, generator=gen).cuda()
    gamma, beta = (torch.rand(C, generator=gen) + 0.5).cuda(), (torch.randn(C, generator=gen) * 0.2).cuda()
    xc = to_cl(x0)
    y1 = torch.zeros((N * S, C), dtype=torch.bfloat16, device="cuda")
    lll = torch.zeros((N * s2, C), dtype=torch.bfloat16, device="cuda")
    hi = torch.zeros((7, N * s2, C), dtype=torch.bfloat16, device="cuda")
    y2 = torch.zeros((N * s2, C), dtype=torch.bfloat16, device="cuda")
    st = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda")
    layers = [ops.conv3d_chain_layer(xc, ops.conv3d_pack_weights(w1), b1, y1, (N, D, H, W), C, C),
              ops.chain_dwt_op(y1, (N, D, H, W), C, lll, hi, lll_bias=emb, lll_scale=1.0 / 3.0, gn_stats=st, gn_groups=G),
              ops.conv3d_chain_layer(lll, ops.conv3d_pack_weights(w2), b2, y2, (N, D // 2, H // 2, W // 2), C, C,
                                     gn_in=(st, gamma, beta, G, EPS))]
    ops.conv3d_chain([l for l, _ in layers], torch.zeros(2, dtype=torch.int64, device="cuda"))
    torch.cuda.synchronize()
    got1 = from_cl(y1, (N, C, D, H, W))
    ref1 = _ref_layer(x0, w1, b1, None, None, None)
    assert float((got1 - ref1).abs().max()) <= 1e-2 * float(ref1.abs().max()) + 1e-3
    lll_ref = torch.zeros_like(lll)
    ops.dwt3d_cl(y1, (N, D, H, W), C, lll_ref, torch.zeros_like(hi), lll_bias=emb, lll_scale=1.0 / 3.0)
    assert torch.equal(lll, lll_ref)
    xin = from_cl(lll, (N, C, D // 2, H // 2, W // 2))
    ref2 = _ref_layer(xin, w2, b2, None, None, (gamma, beta))
    got2 = from_cl(y2, (N, C, D // 2, H // 2, W // 2))
    assert float((got2 - ref2).abs().max()) <= 1e-2 * float(ref2.abs().max()) + 1e-3
