"""GPU parity of the fused diffusion step, q_sample, final image step, GroupNorm+SiLU, the timestep path and
the layout converters, each called through the C-ABI (fcwdm.ops) and checked against the oracle / golden
fixtures produced by the reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import diffusion as od
from oracle import wunet as ow
from oracle.make_golden import toy_model

pytestmark = pytest.mark.gpu


def _tab10():
    b10, m10 = od.respaced_betas(od.named_beta_schedule("linear", 10, "sampled"), od.space_timesteps(10, [10]))
    return od.Tables(b10), m10


def _coef(tab, device="cuda"):
    t = np.arange(tab.num_timesteps)
    sig = np.exp(0.5 * tab.fixed_large_log_variance) * (t != 0)
    c = np.stack([tab.posterior_mean_coef1, tab.posterior_mean_coef2, sig, tab.sqrt_recip_alphas_cumprod,
                  tab.sqrt_recipm1_alphas_cumprod], axis=1).astype(np.float32)
    return torch.from_numpy(c).to(device)


@pytest.mark.parametrize("tval", [9, 4, 0])
def test_p_sample_step_golden(golden, tval):
    from fcwdm import ops
    g = golden("diffusion")
    tab, tmap = _tab10()
    x, cond = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["cond"]).cuda()
    t = torch.tensor([tval, tval], device="cuda")
    t_model = torch.tensor(tmap, device="cuda")[t]
    mo = toy_model(torch.cat([x, cond], 1), t_model)
    torch.manual_seed(100 + tval)
    noise = torch.randn(x.shape)          # the fixture was drawn on the CPU generator
    xp, pred = ops.p_sample_step(mo, x, noise.cuda(), _coef(tab), t)
    np.testing.assert_allclose(pred.cpu().numpy(), g[f"p_sample_t{tval}_pred_xstart"], atol=3e-6)
    np.testing.assert_allclose(xp.cpu().numpy(), g[f"p_sample_t{tval}_sample"], atol=4e-6)


def test_p_sample_step_variants():
    from fcwdm import ops
    from gpu_util import to_cl, bf16_round
    tab, _ = _tab10()
    g = torch.Generator().manual_seed(5)
    for shape in [(2, 8, 3, 5, 7), (1, 8, 4, 4, 8)]:          # scalar path (S % 4 != 0) and vector path
        x = torch.randn(shape, generator=g)
        mo = bf16_round(torch.randn(shape, generator=g) * 0.3 + 0.2)
        nz = torch.randn(shape, generator=g)
        t = torch.tensor([7] * shape[0])
        for clip in (True, False):
            for pxs in (True, False):
                model = lambda xin, tt: mo
                ref = od.p_sample(tab, model, x, t, clip_denoised=clip, predict_xstart=pxs, noise=nz)
                xp, pred = ops.p_sample_step(mo.cuda(), x.cuda(), nz.cuda(), _coef(tab), t.cuda(), clip_denoised=clip,
                                             predict_xstart=pxs)
                np.testing.assert_allclose(pred.cpu().numpy(), ref["pred_xstart"].numpy(), atol=1e-5)
                np.testing.assert_allclose(xp.cpu().numpy(), ref["sample"].numpy(), atol=1e-5)
        # channels-last bf16 model output + bf16 cl copy of x_prev
        mo_cl = to_cl(mo.cuda(), ld=16)
        S = shape[2] * shape[3] * shape[4]
        xp_cl = torch.zeros((shape[0] * S, 64), dtype=torch.bfloat16, device="cuda")
        ref = od.p_sample(tab, lambda a, b: mo, x, t, noise=nz)
        xp, _ = ops.p_sample_step(mo_cl, x.cuda(), nz.cuda(), _coef(tab), t.cuda(), want_pred=False, model_out_cl_ld=16,
                                  x_prev_cl=xp_cl)
        np.testing.assert_allclose(xp.cpu().numpy(), ref["sample"].numpy(), atol=1e-5)
        got_cl = xp_cl[:, :8].float().reshape(shape[0], S, 8).permute(0, 2, 1).reshape(shape)
        np.testing.assert_allclose(got_cl.cpu().numpy(), ref["sample"].numpy(), atol=2e-2, rtol=1e-2)
        assert float(xp_cl[:, 8:].abs().max()) == 0.0


def test_q_sample_and_postprocess_golden(golden):
    from fcwdm import ops
    g = golden("diffusion")
    tab, _ = _tab10()
    coef = torch.from_numpy(np.stack([tab.sqrt_alphas_cumprod, tab.sqrt_one_minus_alphas_cumprod], 1).astype(np.float32)).cuda()
    q = ops.q_sample(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["q_noise"]).cuda(), coef,
                     torch.from_numpy(g["q_t"]).cuda())
    np.testing.assert_allclose(q.cpu().numpy(), g["q_sample"], atol=1e-6)
    img = ops.sample_to_image(torch.from_numpy(g["post_in"]).cuda(), torch.from_numpy(g["post_cond1"]).cuda())
    out = img.squeeze(1)[:, :, :, :155]
    np.testing.assert_allclose(out.cpu().numpy(), g["post_out"], atol=2e-6)


@pytest.mark.parametrize("C,G,S", [(64, 32, 4 * 6 * 8), (128, 32, 333), (256, 32, 50), (32, 32, 64), (64, 32, 70000)])
def test_groupnorm_silu(C, G, S):
    from fcwdm import ops
    from gpu_util import bf16_round
    N = 2
    g = torch.Generator().manual_seed(C + S)
    x = bf16_round(torch.randn(N, S, C, generator=g) * 1.7 + 0.4).cuda()
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    beta = (0.1 * torch.randn(C, generator=g)).cuda()
    ld = (C + 63) // 64 * 64
    xb = torch.zeros((N * S, ld), dtype=torch.bfloat16, device="cuda")
    xb[:, :C] = x.reshape(N * S, C).to(torch.bfloat16)
    yb = torch.zeros_like(xb)
    stats = torch.empty((N, 16, G, 2), dtype=torch.float64, device="cuda")
    ops.groupnorm_silu(xb, yb, stats, gamma, beta, N, S, C, G)
    ref = F.silu(F.group_norm(x.permute(0, 2, 1).float(), G, gamma, beta, 1e-5)).permute(0, 2, 1)
    got = yb[:, :C].float().reshape(N, S, C)
    err = (got - ref).abs()
    assert float(err.max()) < 3e-2 and float(err.mean()) < 3e-3, (float(err.max()), float(err.mean()))
    ops.groupnorm_silu(xb, yb, stats, gamma, beta, N, S, C, G, silu=False)
    ref = F.group_norm(x.permute(0, 2, 1).float(), G, gamma, beta, 1e-5).permute(0, 2, 1)
    assert float((yb[:, :C].float().reshape(N, S, C) - ref).abs().max()) < 4e-2


def test_timestep_path():
    from fcwdm import ops
    t = torch.tensor([0, 1, 111, 555, 999], device="cuda")
    out = torch.empty((5, 64), device="cuda")
    ops.timestep_embedding(t, out, 64)
    ref = ow.timestep_embedding(t.cpu(), 64)
    # |d cos(t*f)| <= t * ulp(f): the fp32 exp of the frequency may differ by 1 ulp between libms (t <= 999)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-4)
    np.testing.assert_allclose(out[:3].cpu().numpy(), ref[:3].numpy(), atol=2e-5)
    # fractional timesteps (rescale_timesteps=True: the model sees t * 1000 / T as float32)
    tf = torch.tensor([0.0, 0.5, 111.0, 333.3, 999.0], device="cuda")
    ops.timestep_embedding(tf, out, 64)
    reff = ow.timestep_embedding(tf.cpu(), 64)
    np.testing.assert_allclose(out.cpu().numpy(), reff.numpy(), atol=1e-4)
    with pytest.raises(TypeError):
        ops.timestep_embedding(tf.double(), out, 64)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5, 64, generator=g)
    W = torch.randn(256, 64, generator=g) * 0.1
    b = torch.randn(256, generator=g)
    y = torch.empty((5, 256), device="cuda")
    ops.linear(x.cuda(), W.cuda(), b.cuda(), y, act_in=0, act_out=1)
    np.testing.assert_allclose(y.cpu().numpy(), F.silu(F.linear(x, W, b)).numpy(), atol=2e-5, rtol=1e-5)
    ops.linear(x.cuda(), W.cuda(), b.cuda(), y, act_in=1, act_out=0)
    np.testing.assert_allclose(y.cpu().numpy(), F.linear(F.silu(x), W, b).numpy(), atol=2e-5, rtol=1e-5)


def test_layout_roundtrip():
    from gpu_util import from_cl, to_cl, bf16_round
    x = bf16_round(torch.randn(2, 24, 3, 5, 7)).cuda()
    buf = to_cl(x, ld=32)
    assert buf.shape == (2 * 105, 32)
    np.testing.assert_array_equal(from_cl(buf, x.shape).cpu().numpy(), x.cpu().numpy())
    np.testing.assert_array_equal(buf[:, :24].float().reshape(2, 105, 24).permute(0, 2, 1).reshape(x.shape).cpu().numpy(),
                                  x.cpu().numpy())
