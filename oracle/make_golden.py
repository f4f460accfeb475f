"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    CUDA_VISIBLE_DEVICES="" python -m oracle.make_golden

CUDA must be hidden because the reference moves its band matrices to the GPU whenever one is visible,
even for CPU inputs (DWT_IDWT/DWT_IDWT_layer.py:505-511).  The fixtures are small (a few hundred KB in
total) and are what pins the oracle -- and through it the CUDA path -- to the reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shims import reference_modules  # noqa: E402
from oracle import wunet as owunet              # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

SMALL_CFG = dict(image_size=16, in_channels=32, model_channels=32, out_channels=8, num_res_blocks=2,
                 attention_resolutions=(), dropout=0.0, channel_mult=(1, 2), dims=3, num_groups=32,
                 bottleneck_attention=False, resblock_updown=True, use_freq=True, additive_skips=False,
                 use_scale_shift_norm=False)


def toy_model(x, t, **kw):
    """Deterministic stand-in denoiser for the diffusion-arithmetic fixtures (8 output channels)."""
    tt = t.float().reshape(-1, 1, 1, 1, 1)
    return 0.6 * x[:, :8] + 0.25 * x[:, 8:16] - 0.1 * x[:, 16:24] + 0.02 * tt + 0.1


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with reference_modules() as ref:
        # ------------------------------------------------------------------ 1. Haar DWT / IDWT
        dwt, idwt = ref.layer.DWT_3D("haar"), ref.layer.IDWT_3D("haar")
        kat = torch.arange(8.0).reshape(1, 1, 2, 2, 2)
        g = torch.Generator().manual_seed(1)
        x = torch.rand(2, 3, 4, 12, 8, generator=g)
        bands = dwt(x)
        rt = idwt(*bands)
        xg = x.clone().requires_grad_(True)
        gout = [torch.randn(b.shape, generator=g) for b in bands]
        torch.autograd.backward(dwt(xg), gout)
        np.savez(os.path.join(GOLDEN, "haar.npz"),
                 kat_bands=np.array([float(b) for b in dwt(kat)], dtype=np.float32),
                 x=x.numpy(), bands=np.stack([b.contiguous().numpy() for b in bands]), roundtrip=rt.numpy(),
                 grad_bands=np.stack([t.numpy() for t in gout]), grad_x=xg.grad.numpy())

        # ------------------------------------------------------------------ 2. schedules / respacing
        gd, respace, su = ref.gd, ref.respace, ref.script_util
        sch = {}
        sch["sampled10"] = gd.get_named_beta_schedule("linear", 10, "sampled")
        sch["direct20"] = gd.get_named_beta_schedule("linear", 20, "direct")
        sch["direct1000"] = gd.get_named_beta_schedule("linear", 1000, "direct")
        sch["cosine50"] = gd.get_named_beta_schedule("cosine", 50)
        sch["space_1000_100"] = np.array(sorted(respace.space_timesteps(1000, "100")))
        sch["space_1000_ddim50"] = np.array(sorted(respace.space_timesteps(1000, "ddim50")))
        sch["space_300_10_15_20"] = np.array(sorted(respace.space_timesteps(300, [10, 15, 20])))
        sch["space_1000_10_10_10"] = np.array(sorted(respace.space_timesteps(1000, "10,10,10")))
        d100 = su.create_gaussian_diffusion(steps=1000, predict_xstart=True, timestep_respacing="100", mode="i2i")
        sch["respaced100_betas"] = d100.betas
        sch["respaced100_map"] = np.array(d100.timestep_map)
        d10 = su.create_gaussian_diffusion(steps=10, predict_xstart=True, sample_schedule="sampled", mode="i2i")
        for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                     "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                     "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                     "posterior_mean_coef2"):
            sch["d10_" + name] = getattr(d10, name)
        sch["d10_map"] = np.array(d10.timestep_map)
        np.savez(os.path.join(GOLDEN, "schedules.npz"), **sch)

        # ------------------------------------------------------------------ 3. diffusion arithmetic (toy model)
        g = torch.Generator().manual_seed(2)
        xt = torch.randn(2, 8, 4, 6, 4, generator=g)
        cond = torch.rand(2, 24, 4, 6, 4, generator=g)
        dif = {"x": xt.numpy(), "cond": cond.numpy()}
        for tval in (9, 4, 0):
            t = torch.tensor([tval, tval])
            torch.manual_seed(100 + tval)
            with torch.no_grad():
                out = d10.p_sample(toy_model, xt, t, clip_denoised=True, model_kwargs={}, cond=cond)
            dif[f"p_sample_t{tval}_sample"] = out["sample"].numpy()
            dif[f"p_sample_t{tval}_pred_xstart"] = out["pred_xstart"].contiguous().numpy()
        t = torch.tensor([3, 7])
        nz = torch.randn(2, 8, 4, 6, 4, generator=g)
        dif["q_t"] = t.numpy()
        dif["q_noise"] = nz.numpy()
        dif["q_sample"] = d10.q_sample(xt, t, noise=nz).numpy()
        torch.manual_seed(7)
        finals = None
        for out in d10.p_sample_loop_progressive(toy_model, xt.shape, time=d10.num_timesteps, noise=xt,
                                                 clip_denoised=True, model_kwargs={}, cond=cond, progress=False,
                                                 device=torch.device("cpu")):
            finals = out
        dif["loop_final"] = finals["sample"].numpy()
        # training_losses (i2i)
        batch = {k: torch.rand(2, 1, 8, 12, 8, generator=g) for k in ("t1n", "t1c", "t2w", "t2f")}
        t = torch.tensor([2, 8])
        torch.manual_seed(11)
        terms, mo, mo_idwt = d10.training_losses(toy_model, batch, t, model_kwargs={}, mode="i2i", contr="t1n")
        for k, v in batch.items():
            dif["tl_" + k] = v.numpy()
        dif["tl_t"] = t.numpy()
        dif["tl_mse_wav"] = terms["mse_wav"].numpy()
        dif["tl_model_output"] = mo.numpy()
        dif["tl_model_output_idwt"] = mo_idwt.numpy()
        # sample.py post-processing (:113-131)
        sw = torch.randn(1, 8, 4, 6, 80, generator=g) * 0.5 + 0.3
        c1 = torch.rand(1, 1, 8, 12, 160, generator=g)
        c1[c1 < 0.3] = 0
        B, _, D, H, W = sw.size()
        s = idwt(sw[:, 0].view(B, 1, D, H, W) * 3., *[sw[:, i].view(B, 1, D, H, W) for i in range(1, 8)])
        s[s <= 0] = 0
        s[s >= 1] = 1
        s[c1 == 0] = 0
        dif["post_in"] = sw.numpy()
        dif["post_cond1"] = c1.numpy()
        dif["post_out"] = s.squeeze(1)[:, :, :, :155].numpy()
        np.savez(os.path.join(GOLDEN, "diffusion.npz"), **dif)

        # ------------------------------------------------------------------ 4. WavUNetModel (small config)
        model = ref.wunet.WavUNetModel(**SMALL_CFG)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        sd = owunet.tie_output_blocks(owunet.seeded_state_dict(shapes, seed=0), len(SMALL_CFG["channel_mult"]))
        model.load_state_dict(sd, strict=True)
        model.eval()
        g = torch.Generator().manual_seed(3)
        xin = torch.randn(2, 32, 8, 8, 8, generator=g)
        tin = torch.tensor([3, 977])
        with torch.no_grad():
            y = model(xin, tin)
        np.savez(os.path.join(GOLDEN, "wunet_small.npz"), x=xin.numpy(), t=tin.numpy(), y=y.numpy(),
                 keys=np.array(sorted(shapes)),
                 shapes=np.array([",".join(map(str, shapes[k])) for k in sorted(shapes)]),
                 n_params=sum(p.numel() for p in model.parameters()))

        # end-to-end: respaced loop through the small U-Net
        d4 = su.create_gaussian_diffusion(steps=1000, predict_xstart=True, timestep_respacing="4", mode="i2i")
        xt = torch.randn(1, 8, 8, 8, 8, generator=g)
        cond = torch.rand(1, 24, 8, 8, 8, generator=g)
        torch.manual_seed(5)
        outs = []
        for out in d4.p_sample_loop_progressive(model, xt.shape, time=d4.num_timesteps, noise=xt, clip_denoised=True,
                                                 model_kwargs={}, cond=cond, progress=False,
                                                 device=torch.device("cpu")):
            outs.append(out["sample"].numpy())
        np.savez(os.path.join(GOLDEN, "loop_small.npz"), x=xt.numpy(), cond=cond.numpy(), samples=np.stack(outs),
                 betas=d4.betas, tmap=np.array(d4.timestep_map))

        # CFG-W4 key list + shapes (346 keys / 54,285,640 params) -- checked by the drop-in model
        big = ref.wunet.WavUNetModel(image_size=224, in_channels=32, model_channels=64, out_channels=8,
                                     num_res_blocks=2, attention_resolutions=(), channel_mult=(1, 2, 2, 4), dims=3,
                                     num_groups=32, bottleneck_attention=False, resblock_updown=True, use_freq=True)
        bsd = big.state_dict()
        np.savez(os.path.join(GOLDEN, "cfg_w4_keys.npz"), keys=np.array(list(bsd.keys())),
                 shapes=np.array([",".join(map(str, v.shape)) for v in bsd.values()]),
                 n_params=sum(p.numel() for p in big.parameters()))
    print("golden fixtures written to", GOLDEN)
    for f in sorted(os.listdir(GOLDEN)):
        print(f"  {f}: {os.path.getsize(os.path.join(GOLDEN, f))} bytes")


if __name__ == "__main__":
    main()
