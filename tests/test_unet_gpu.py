"""GPU parity of the drop-in plain UNetModel (run.sh's use_freq=False model) against the
reference-generated fixture and the oracle, plus the resampling kernels against torch.

Stated bf16 tolerance, as for WavUNetModel: relative L2 <= 3e-2, max-abs <= 6e-2 * max|ref|, PSNR >= 40 dB."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import unet as ou
from oracle import wunet as ow
from oracle.make_golden_unet import UNET_SMALL_CFG

pytestmark = pytest.mark.gpu


def seeded_unet(cfg, seed=0):
    from guided_diffusion.unet import UNetModel
    m = UNetModel(**cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = ow.seeded_state_dict(shapes, seed=seed)
    m.load_state_dict(sd, strict=True)
    m.to("cuda")
    m.eval()
    return m, sd


def check(y, ref, label):
    from test_wunet_gpu import psnr
    rel = float((y - ref).norm() / ref.norm())
    mx = float((y - ref).abs().max())
    print(f"{label}: rel-L2 {rel:.3e} max-abs {mx:.3e} ref-max {float(ref.abs().max()):.3f} PSNR {psnr(y, ref):.1f} dB")
    assert rel <= 3e-2 and mx <= 6e-2 * float(ref.abs().max()) and psnr(y, ref) >= 40.0


@pytest.mark.parametrize("depth", [True, False])
def test_resample_kernels(depth):
    from fcwdm import ops
    from gpu_util import bf16_round, from_cl, to_cl
    N, C, D, H, W = 2, 64, 4, 6, 10
    x = bf16_round(torch.randn(N, C, D, H, W, generator=torch.Generator().manual_seed(1))).cuda()
    xc = to_cl(x)
    fd = 2 if depth else 1
    y = torch.zeros((N * (D // fd) * (H // 2) * (W // 2), 64), dtype=torch.bfloat16, device="cuda")
    ops.avgpool2_cl(xc, (N, D, H, W), C, y, pool_depth=depth)
    k = 2 if depth else (1, 2, 2)
    ref = F.avg_pool3d(x, kernel_size=k, stride=k)
    got = from_cl(y, tuple(ref.shape))
    assert float((got - ref).abs().max()) <= 2.0 ** -8 * float(ref.abs().max()) + 1e-6
    u = torch.zeros((N * D * fd * H * 2 * W * 2, 64), dtype=torch.bfloat16, device="cuda")
    ops.upsample2_cl(xc, (N, D, H, W), C, u, up_depth=depth)
    ref = F.interpolate(x, scale_factor=2, mode="nearest") if depth else F.interpolate(x, (D, 2 * H, 2 * W), mode="nearest")
    assert torch.equal(from_cl(u, tuple(ref.shape)), ref)


def test_unet_small_matches_reference_fixture(golden):
    g = golden("unet_small")
    m, _ = seeded_unet(UNET_SMALL_CFG)
    with torch.no_grad():
        y = m(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda())
    ref = torch.from_numpy(g["y"])
    assert y.shape == ref.shape and y.dtype == torch.float32
    check(y.cpu(), ref, "small plain unet vs reference fixture")


@pytest.mark.parametrize("resample_2d", [False, True])
def test_unet_wide_matches_oracle(resample_2d):
    """64 base channels (CTA-pair kernel, concat widths 128/192/256), ragged spatial tiles, both resampling modes."""
    cfg = dict(UNET_SMALL_CFG, model_channels=64, resample_2d=resample_2d)
    m, sd = seeded_unet(cfg, seed=2)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 32, 8, 20, 12, generator=g)
    t = torch.tensor([812])
    with torch.no_grad():
        y = m(x.cuda(), t.cuda()).cpu()
    ref = ou.unet_forward(sd, x, t, model_channels=64, channel_mult=(1, 2, 2), resample_2d=resample_2d)
    check(y, ref, f"64-channel plain unet vs oracle (resample_2d={resample_2d})")


def test_unet_sampling_loop_runs_fused(golden):
    """p_sample_loop through the plain U-Net takes the fused CUDA-graph sampler; eager and graph steps agree."""
    import os
    from guided_diffusion.script_util import create_gaussian_diffusion
    m, _ = seeded_unet(UNET_SMALL_CFG)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 8, 8, 8, 8, generator=g).cuda()
    cond = torch.rand(1, 24, 8, 8, 8, generator=g).cuda()
    outs = {}
    for mode in ("graph", "eager"):
        os.environ["FCWDM_NO_GRAPH"] = "1" if mode == "eager" else "0"
        d = create_gaussian_diffusion(steps=1000, predict_xstart=True, timestep_respacing="4", mode="i2i")
        torch.manual_seed(21)
        outs[mode] = d.p_sample_loop(m, x.shape, noise=x.clone(), cond=cond, progress=False)
        assert len(d._samplers) == 1
    os.environ.pop("FCWDM_NO_GRAPH", None)
    assert torch.isfinite(outs["graph"]).all()
    assert float((outs["graph"] - outs["eager"]).norm() / outs["eager"].norm()) <= 5e-3
