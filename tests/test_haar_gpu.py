"""GPU parity of the Haar DWT/IDWT kernels (through the drop-in modules and the C-ABI) against the oracle and
the reference-generated golden fixtures.  fp32 tolerance: 1e-6 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import haar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from DWT_IDWT.DWT_IDWT_layer import DWT_3D, IDWT_3D
    return DWT_3D("haar"), IDWT_3D("haar")


def test_golden_bands_roundtrip_backward(golden, mods):
    dwt, idwt = mods
    g = golden("haar")
    x = torch.from_numpy(g["x"]).cuda()
    kat = dwt(torch.arange(8.0, device="cuda").reshape(1, 1, 2, 2, 2))
    np.testing.assert_allclose(np.array([float(b) for b in kat]), g["kat_bands"], atol=1e-6)
    bands = dwt(x)
    assert len(bands) == 8
    for i, b in enumerate(bands):
        assert b.shape == (2, 3, 2, 6, 4)
        np.testing.assert_allclose(b.cpu().numpy(), g["bands"][i], rtol=0, atol=4e-7)
    rt = idwt(*bands)
    np.testing.assert_allclose(rt.cpu().numpy(), g["roundtrip"], rtol=0, atol=5e-7)
    xg = x.clone().requires_grad_(True)
    torch.autograd.backward(dwt(xg), [torch.from_numpy(t).cuda() for t in g["grad_bands"]])
    np.testing.assert_allclose(xg.grad.cpu().numpy(), g["grad_x"], rtol=0, atol=2e-6)
    # IDWT backward == DWT of the upstream gradient
    bl = [b.detach().clone().requires_grad_(True) for b in bands]
    go = torch.from_numpy(g["x"]).cuda() * 0.5 + 0.25
    idwt(*bl).backward(go)
    ref = haar.dwt3d(go.cpu().numpy())
    for i in range(8):
        np.testing.assert_allclose(bl[i].grad.cpu().numpy(), ref[i], rtol=0, atol=5e-7)


@pytest.mark.parametrize("shape", [(1, 1, 2, 2, 2), (2, 3, 6, 10, 14), (1, 2, 8, 8, 16), (3, 1, 4, 2, 24),
                                   (1, 5, 10, 12, 8), (1, 1, 32, 48, 40)])
def test_vs_oracle_shapes(shape, mods):
    dwt, idwt = mods
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(shape, generator=g)
    ref = haar.dwt3d(x.numpy())
    got = dwt(x.cuda())
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=0, atol=4e-7)
    np.testing.assert_allclose(idwt(*got).cpu().numpy(), haar.idwt3d(*ref), rtol=0, atol=5e-7)


def test_noncontiguous_and_channel_sliced_inputs(mods):
    dwt, idwt = mods
    x = torch.rand(2, 6, 8, 8, 16, device="cuda")
    xs = x[:, 1:5]                      # N/C strides differ from a packed tensor, spatial dims contiguous
    ref = haar.dwt3d(xs.cpu().numpy())
    for a, b in zip(dwt(xs), ref):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=0, atol=4e-7)
    xt = x.permute(0, 1, 3, 2, 4)       # spatial dims not contiguous -> wrapper makes a packed copy
    ref = haar.dwt3d(xt.cpu().numpy())
    for a, b in zip(dwt(xt), ref):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=0, atol=4e-7)


def test_edge_cases(mods):
    from fcwdm import FcwdmError
    dwt, idwt = mods
    out = dwt(torch.zeros(0, 1, 4, 4, 8, device="cuda"))          # empty batch
    assert all(o.shape == (0, 1, 2, 2, 4) for o in out)
    with pytest.raises(FcwdmError):
        dwt(torch.zeros(1, 1, 3, 4, 4, device="cuda"))            # odd depth
    with pytest.raises(AssertionError):
        dwt(torch.zeros(1, 4, 4, 4, device="cuda"))               # reference asserts 5-D
    with pytest.raises(FcwdmError):
        dwt(torch.zeros(1, 1, 2, 2, 2))                           # CPU tensor: no fallback
    with pytest.raises(TypeError):
        dwt(torch.zeros(1, 1, 2, 2, 2, dtype=torch.float64, device="cuda"))
    # depth larger than max(H, W): the reference fails (layer.py:465), the kernel does not
    x = torch.rand(1, 1, 16, 4, 8, device="cuda")
    np.testing.assert_allclose(idwt(*dwt(x)).cpu().numpy(), x.cpu().numpy(), atol=5e-7)


def test_bf16(mods):
    dwt, idwt = mods
    x = torch.rand(1, 2, 8, 12, 16).to(torch.bfloat16)
    ref = haar.dwt3d(x.float().numpy())
    got = dwt(x.cuda())
    for a, b in zip(got, ref):
        assert a.dtype == torch.bfloat16
        np.testing.assert_allclose(a.float().cpu().numpy(), b, rtol=0, atol=2 ** -7)     # bf16 output rounding
    rt = idwt(*got).float().cpu()
    assert (rt - x.float()).abs().max() < 3e-2
    xo = torch.rand(1, 1, 4, 6, 10).to(torch.bfloat16)      # W % 8 != 0 -> generic kernel
    ref = haar.dwt3d(xo.float().numpy())
    for a, b in zip(dwt(xo.cuda()), ref):
        np.testing.assert_allclose(a.float().cpu().numpy(), b, rtol=0, atol=2 ** -7)


def test_full_size_properties(mods):
    """BASELINE config 1/2 size: 1x4x224x224x160 fp32.  Size-independent properties: perfect reconstruction
    (rel-L2 <= 1e-6), energy conservation (orthonormal transform), linearity, and the closed-form butterfly on a
    sub-block."""
    dwt, idwt = mods
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 4, 224, 224, 160, generator=g).cuda()
    bands = dwt(x)
    assert all(b.shape == (1, 4, 112, 112, 80) for b in bands)
    rt = idwt(*bands)
    rel = float((rt - x).double().norm() / x.double().norm())
    assert rel <= 1e-6, rel
    assert float((rt - x).abs().max()) <= 1e-6
    e_in = float(x.double().pow(2).sum())
    e_out = float(sum(b.double().pow(2).sum() for b in bands))
    assert abs(e_in - e_out) / e_in < 1e-6
    y = torch.rand(1, 4, 224, 224, 160, generator=g).cuda()
    by = dwt(y)
    bxy = dwt(2.0 * x - 0.5 * y)
    for a, b, c in zip(bands, by, bxy):
        assert float((2.0 * a - 0.5 * b - c).abs().max()) < 3e-6
    sub = x[:, :1, :16, :16, :32].cpu().numpy()
    ref = haar.dwt3d_butterfly_f64(sub)
    for a, b in zip(bands, ref):
        np.testing.assert_allclose(a[:, :1, :8, :8, :16].cpu().numpy(), b, atol=5e-7)


def test_concat_and_scale_layouts():
    from fcwdm import ops
    x = torch.rand(2, 1, 8, 8, 16, device="cuda")
    ref = haar.dwt3d(x.cpu().numpy())
    cat = ops.dwt3d_planar(x, lll_scale=1.0 / 3.0, concat=True)          # sample.py:92-93
    expect = np.concatenate([ref[0] / np.float32(3.0)] + list(ref[1:]), axis=1)
    np.testing.assert_allclose(cat.cpu().numpy(), expect, atol=4e-7)
    back = ops.idwt3d_planar(cat, lll_scale=3.0, concat=True)            # sample.py:113-121
    np.testing.assert_allclose(back.cpu().numpy(), x.cpu().numpy(), atol=1e-6)


def test_channels_last_variants():
    from fcwdm import ops
    from gpu_util import from_cl, to_cl, bf16_round
    N, C, D, H, W = 2, 64, 4, 6, 8
    x = bf16_round(torch.randn(N, C, D, H, W, device="cuda"))
    xc = to_cl(x)
    ref = haar.dwt3d(x.cpu().numpy())
    s = (D // 2) * (H // 2) * (W // 2)
    lll = torch.zeros((N * s, 64), dtype=torch.bfloat16, device="cuda")
    hi = torch.zeros((7, N * s, 64), dtype=torch.bfloat16, device="cuda")
    bias = torch.randn(N, C, device="cuda")
    ops.dwt3d_cl(xc, (N, D, H, W), C, lll, hi, lll_bias=bias, lll_scale=1.0 / 3.0)
    shp = (N, C, D // 2, H // 2, W // 2)
    exp0 = ref[0] / 3.0 + bias.cpu().numpy()[:, :, None, None, None]
    np.testing.assert_allclose(from_cl(lll, shp).cpu().numpy(), exp0, atol=3e-2, rtol=1e-2)
    for b in range(1, 8):
        np.testing.assert_allclose(from_cl(hi[b - 1], shp).cpu().numpy(), ref[b], atol=3e-2, rtol=1e-2)
    # LLL-only variant (wunet.py:241)
    lll2 = torch.zeros_like(lll)
    ops.dwt3d_cl(xc, (N, D, H, W), C, lll2, None, lll_scale=1.0 / 3.0)
    np.testing.assert_allclose(from_cl(lll2, shp).cpu().numpy(), ref[0] / 3.0, atol=3e-2, rtol=1e-2)
    # concatenated / 3 variant (WaveletDownsample, wunet.py:143-144)
    cat = torch.zeros((N * s, 8 * C), dtype=torch.bfloat16, device="cuda")
    ops.dwt3d_cl(xc, (N, D, H, W), C, cat[:, :C], cat[:, C:], lll_scale=1.0 / 3.0, hi_scale=1.0 / 3.0, hi_sb=C)
    expc = np.concatenate(ref, axis=1) / 3.0
    np.testing.assert_allclose(from_cl(cat, (N, 8 * C, D // 2, H // 2, W // 2)).cpu().numpy(), expc, atol=3e-2, rtol=1e-2)
    # inverse with bias (Upsample + timestep embedding add, wunet.py:76,262)
    y = torch.zeros((N * D * H * W, 64), dtype=torch.bfloat16, device="cuda")
    lll3 = torch.zeros_like(lll)
    ops.dwt3d_cl(xc, (N, D, H, W), C, lll3, hi, lll_scale=1.0 / 3.0)
    ops.idwt3d_cl(lll3, hi, (N, D, H, W), C, y, bias=bias, lll_scale=3.0)
    expy = x.cpu().numpy() + bias.cpu().numpy()[:, :, None, None, None]
    np.testing.assert_allclose(from_cl(y, (N, C, D, H, W)).cpu().numpy(), expy, atol=6e-2, rtol=2e-2)
