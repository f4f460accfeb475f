"""Pins the oracle (oracle/*.py) against fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import diffusion as od
from oracle import haar
from oracle import wunet as ow
from oracle.make_golden import SMALL_CFG, toy_model


def test_haar_kat(golden):
    g = golden("haar")
    bands = haar.dwt3d(np.arange(8, dtype=np.float32).reshape(1, 1, 2, 2, 2))
    got = np.array([float(b.ravel()[0]) for b in bands], dtype=np.float32)
    np.testing.assert_allclose(got, g["kat_bands"], atol=1e-6)
    # SURVEY.md section 4 KAT values
    np.testing.assert_allclose(got[[0, 1, 2, 4]], [9.899494, -1.414213, -2.828427, -5.656854], atol=2e-6)


def test_haar_bands_and_roundtrip(golden):
    g = golden("haar")
    bands = haar.dwt3d(g["x"])
    for i, b in enumerate(bands):
        np.testing.assert_allclose(b, g["bands"][i], rtol=0, atol=4e-7)
    np.testing.assert_allclose(haar.idwt3d(*bands), g["roundtrip"], rtol=0, atol=5e-7)
    np.testing.assert_allclose(haar.idwt3d(*bands), g["x"], rtol=0, atol=5e-7)


def test_haar_backward_is_idwt(golden):
    g = golden("haar")
    np.testing.assert_allclose(haar.idwt3d(*g["grad_bands"]), g["grad_x"], rtol=0, atol=2e-6)


def test_haar_closed_form(golden):
    g = golden("haar")
    ref = haar.dwt3d_butterfly_f64(g["x"])
    for a, b in zip(haar.dwt3d(g["x"]), ref):
        np.testing.assert_allclose(a, b, rtol=0, atol=5e-7)


def test_schedules(golden):
    g = golden("schedules")
    np.testing.assert_array_equal(od.named_beta_schedule("linear", 10, "sampled"), g["sampled10"])
    np.testing.assert_array_equal(od.named_beta_schedule("linear", 20, "direct"), g["direct20"])
    np.testing.assert_array_equal(od.named_beta_schedule("linear", 1000, "direct"), g["direct1000"])
    np.testing.assert_allclose(od.named_beta_schedule("cosine", 50), g["cosine50"], rtol=1e-15)
    assert sorted(od.space_timesteps(1000, "100")) == list(g["space_1000_100"])
    assert sorted(od.space_timesteps(1000, "ddim50")) == list(g["space_1000_ddim50"])
    assert sorted(od.space_timesteps(300, [10, 15, 20])) == list(g["space_300_10_15_20"])
    assert sorted(od.space_timesteps(1000, "10,10,10")) == list(g["space_1000_10_10_10"])
    betas, tmap = od.respaced_betas(od.named_beta_schedule("linear", 1000, "direct"), od.space_timesteps(1000, "100"))
    np.testing.assert_array_equal(betas, g["respaced100_betas"])
    assert tmap == list(g["respaced100_map"])
    b10, m10 = od.respaced_betas(od.named_beta_schedule("linear", 10, "sampled"), od.space_timesteps(10, [10]))
    assert m10 == list(g["d10_map"])
    tab = od.Tables(b10)
    for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                 "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                 "posterior_mean_coef2"):
        np.testing.assert_array_equal(getattr(tab, name), g["d10_" + name], err_msg=name)


def _tab10():
    b10, m10 = od.respaced_betas(od.named_beta_schedule("linear", 10, "sampled"), od.space_timesteps(10, [10]))
    return od.Tables(b10), m10


def test_p_sample_and_q_sample(golden):
    g = golden("diffusion")
    tab, tmap = _tab10()
    x, cond = torch.from_numpy(g["x"]), torch.from_numpy(g["cond"])
    for tval in (9, 4, 0):
        t = torch.tensor([tval, tval])
        torch.manual_seed(100 + tval)
        out = od.p_sample(tab, toy_model, x, t, cond=cond, timestep_map=tmap)
        np.testing.assert_allclose(out["pred_xstart"].numpy(), g[f"p_sample_t{tval}_pred_xstart"], atol=2e-6)
        np.testing.assert_allclose(out["sample"].numpy(), g[f"p_sample_t{tval}_sample"], atol=3e-6)
    q = od.q_sample(tab, x, torch.from_numpy(g["q_t"]), torch.from_numpy(g["q_noise"]))
    np.testing.assert_allclose(q.numpy(), g["q_sample"], atol=1e-6)


def test_p_sample_loop(golden):
    g = golden("diffusion")
    tab, tmap = _tab10()
    torch.manual_seed(7)
    out = od.p_sample_loop(tab, toy_model, torch.from_numpy(g["x"]), cond=torch.from_numpy(g["cond"]), timestep_map=tmap)
    np.testing.assert_allclose(out.numpy(), g["loop_final"], atol=2e-5)


def test_training_losses(golden):
    g = golden("diffusion")
    tab, tmap = _tab10()
    batch = {k: torch.from_numpy(g["tl_" + k]) for k in ("t1n", "t1c", "t2w", "t2f")}
    torch.manual_seed(11)
    terms, mo, mo_idwt = od.training_losses(tab, toy_model, batch, torch.from_numpy(g["tl_t"]), timestep_map=tmap)
    np.testing.assert_allclose(mo.numpy(), g["tl_model_output"], atol=3e-6)
    np.testing.assert_allclose(mo_idwt.numpy(), g["tl_model_output_idwt"], atol=5e-6)
    np.testing.assert_allclose(terms["mse_wav"].numpy(), g["tl_mse_wav"], rtol=1e-5)


def test_sample_postprocess(golden):
    g = golden("diffusion")
    out = od.sample_postprocess(torch.from_numpy(g["post_in"]), torch.from_numpy(g["post_cond1"]))
    assert out.shape == (1, 8, 12, 155)
    np.testing.assert_allclose(out.numpy(), g["post_out"], atol=2e-6)


def _small_sd(golden, fixture="wunet_small"):
    g = golden(fixture)
    shapes = {str(k): tuple(int(v) for v in str(s).split(",")) for k, s in zip(g["keys"], g["shapes"])}
    return ow.tie_output_blocks(ow.seeded_state_dict(shapes, seed=0), len(SMALL_CFG["channel_mult"]))


def test_wunet_small(golden):
    g = golden("wunet_small")
    sd = _small_sd(golden)
    y = ow.wunet_forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), model_channels=32,
                         channel_mult=(1, 2))
    err = (y.numpy() - g["y"])
    assert np.abs(err).max() < 5e-5 * max(1.0, np.abs(g["y"]).max()), np.abs(err).max()


def test_loop_small(golden):
    g = golden("loop_small")
    sd = _small_sd(golden)
    tab = od.Tables(g["betas"])
    model = lambda x, t: ow.wunet_forward(sd, x, t, model_channels=32, channel_mult=(1, 2))
    torch.manual_seed(5)
    img = torch.from_numpy(g["x"])
    cond = torch.from_numpy(g["cond"])
    for k, i in enumerate(reversed(range(tab.num_timesteps))):
        img = od.p_sample(tab, model, img, torch.tensor([i]), cond=cond, timestep_map=list(g["tmap"]))["sample"]
        np.testing.assert_allclose(img.numpy(), g["samples"][k], atol=2e-4)


@pytest.mark.parametrize("fixture,weights", [("train_small", "wunet_small"), ("train_small_ssn", "wunet_small_ssn")])
def test_training_step_small(golden, fixture, weights):
    """Loss, every parameter gradient and the AdamW update of one reference training step (train_small.npz; the
    _ssn pair is the same step with use_scale_shift_norm=True)."""
    from oracle import train as otr
    g = golden(fixture)
    sd = _small_sd(golden, weights)
    tab, m10 = _tab10()
    batch = {k: torch.from_numpy(g["batch_" + k]) for k in ("t1n", "t1c", "t2w", "t2f")}
    loss, mse, out, grads = otr.training_step_grads(sd, tab, batch, torch.from_numpy(g["t"]), torch.from_numpy(g["noise"]),
                                                    model_channels=32, channel_mult=(1, 2), timestep_map=m10)
    np.testing.assert_allclose(mse.numpy(), g["mse_wav"], rtol=2e-5)
    assert abs(float(loss) - float(g["loss"])) <= 2e-5 * float(g["loss"])
    np.testing.assert_allclose(out.numpy(), g["model_output"], atol=5e-5 * float(np.abs(g["model_output"]).max()))
    names = [str(n) for n in g["param_names"]]
    assert len(names) == 178
    for name, ref_norm in zip(names, g["grad_norms"]):
        gr = grads[name]
        assert abs(float(gr.double().norm()) - ref_norm) <= 1e-3 * ref_norm + 1e-6, name   # exact-zero gradients (a per-channel
        # constant in front of a 1-channel-per-group GroupNorm) are rounding noise ~1e-9 on both sides
        if "grad/" + name in g.files:
            ref = g["grad/" + name]
            np.testing.assert_allclose(gr.numpy(), ref, atol=1e-3 * float(np.abs(ref).max()) + 1e-7, err_msg=name)
        else:
            ref = g["gradslice/" + name]
            np.testing.assert_allclose(gr[:2].numpy(), ref, atol=1e-3 * float(np.abs(ref).max()) + 1e-7, err_msg=name)
    for key in g.files:
        if key.startswith("stepped/"):
            name = key[len("stepped/"):]
            p, _, _ = otr.adamw_step(sd[name], grads[name], torch.zeros_like(sd[name]), torch.zeros_like(sd[name]), 1,
                                     float(g["lr"]), weight_decay=float(g["wd"]))
            # Adam's first step moves every weight by ~lr * sign(g): where the true gradient is exactly zero the sign is
            # rounding noise, so compare only where |g| is far above eps = 1e-8
            mask = np.abs(g["grad/" + name]) > 1e-5
            np.testing.assert_allclose(p.numpy()[mask], g[key][mask], atol=2e-5, err_msg=name)


def test_unet_small(golden):
    """Plain UNetModel (run.sh's use_freq=False model) forward against the reference fixture."""
    from oracle import unet as ou
    g = golden("unet_small")
    shapes = {str(k): tuple(int(v) for v in str(s).split(",")) for k, s in zip(g["keys"], g["shapes"])}
    sd = ow.seeded_state_dict(shapes, seed=0)
    y = ou.unet_forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), model_channels=32, channel_mult=(1, 2, 2))
    assert np.abs(y.numpy() - g["y"]).max() < 5e-5 * max(1.0, np.abs(g["y"]).max())


def test_training_step_unet_small(golden):
    """One reference training step of the plain UNetModel (train_unet_small.npz): loss and every parameter gradient."""
    from oracle import train as otr
    from oracle import unet as ou
    g = golden("train_unet_small")
    gu = golden("unet_small")
    shapes = {str(k): tuple(int(v) for v in str(s).split(",")) for k, s in zip(gu["keys"], gu["shapes"])}
    sd = ow.seeded_state_dict(shapes, seed=0)
    tab, m10 = _tab10()
    batch = {k: torch.from_numpy(g["batch_" + k]) for k in ("t1n", "t1c", "t2w", "t2f")}
    loss, mse, out, grads = otr.training_step_grads(sd, tab, batch, torch.from_numpy(g["t"]), torch.from_numpy(g["noise"]),
                                                    model_channels=32, channel_mult=(1, 2, 2), timestep_map=m10,
                                                    forward=ou.unet_forward)
    assert abs(float(loss) - float(g["loss"])) <= 2e-5 * float(g["loss"])
    np.testing.assert_allclose(mse.numpy(), g["mse_wav"], rtol=2e-5)
    for name, ref_norm in zip((str(n) for n in g["param_names"]), g["grad_norms"]):
        gr = grads[name]
        assert abs(float(gr.double().norm()) - ref_norm) <= 1e-3 * ref_norm + 1e-6, name
        key = "grad/" + name if "grad/" + name in g.files else "gradslice/" + name
        got = gr.numpy() if key.startswith("grad/") else gr[:2].numpy()
        np.testing.assert_allclose(got, g[key], atol=1e-3 * float(np.abs(g[key]).max()) + 1e-7, err_msg=name)


def test_preprocess(golden):
    """clip_and_normalize of the reference's loader (bratsloader.py:107-111) on two fixtures."""
    from oracle import preprocess as op
    g = golden("preprocess")
    for k in ("vol", "neg"):
        np.testing.assert_allclose(op.clip_and_normalize(g[k]), g[k + "_out"], rtol=0, atol=1e-15)
    out = op.preprocess_volume(g["vol"], crop=4, pad_to=32)
    assert out.shape == (1, 32, 28, 32) and out.dtype == np.float32
    np.testing.assert_array_equal(out[0, :, :, :23], g["vol_out"][4:-4, 4:-4].astype(np.float32))
    assert float(np.abs(out[..., 23:]).max()) == 0.0


def test_scale_shift_norm_small(golden):
    """use_scale_shift_norm=True (wunet.py:256-260, unet.py:301-305): both oracle U-Nets against forwards of the
    unmodified reference (oracle/make_golden_ssn.py)."""
    from oracle import unet as ou
    for name, forward, mult, tie in (("wunet_small_ssn", ow.wunet_forward, (1, 2), True),
                                     ("unet_small_ssn", ou.unet_forward, (1, 2, 2), False)):
        g = golden(name)
        shapes = {str(k): tuple(int(v) for v in str(s).split(",")) for k, s in zip(g["keys"], g["shapes"])}
        sd = ow.seeded_state_dict(shapes, seed=0)
        if tie:
            sd = ow.tie_output_blocks(sd, len(mult))
        y = forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), model_channels=32, channel_mult=mult)
        assert np.abs(y.numpy() - g["y"]).max() < 5e-5 * max(1.0, np.abs(g["y"]).max()), name
