"""Config-2 quality check at FULL size: one 224x224x160 volume sampled with T=10 ('sampled' schedule, CFG-W4, seeded
random weights) through the fcwdm path on the GPU and through the fp32 CPU oracle with the same per-step noise, compared
in image space after the final IDWT + clamp + mask: max-abs, PSNR, SSIM (11x11x11 uniform window) and the per-step
wavelet-domain relative L2.  The oracle needs ~3 s per step on 16 cores.   usage: quality_probe.py [T]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402
from fcwdm import ops, pipeline  # noqa: E402
from oracle import diffusion as od  # noqa: E402
from oracle import wunet as ow  # noqa: E402


def ssim3d(a, b, win=11, L=1.0):
    """Mean SSIM of two (D,H,W) volumes in [0, L] with a uniform win^3 window (K1 = 0.01, K2 = 0.03)."""
    a, b = a[None, None].double(), b[None, None].double()
    c1, c2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    mu_a, mu_b = F.avg_pool3d(a, win, 1), F.avg_pool3d(b, win, 1)
    va = F.avg_pool3d(a * a, win, 1) - mu_a ** 2
    vb = F.avg_pool3d(b * b, win, 1) - mu_b ** 2
    cov = F.avg_pool3d(a * b, win, 1) - mu_a * mu_b
    return float((((2 * mu_a * mu_b + c1) * (2 * cov + c2)) / ((mu_a ** 2 + mu_b ** 2 + c1) * (va + vb + c2))).mean())


T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda")
torch.set_num_threads(os.cpu_count() or 1)
model, diffusion = bench.build_model(dev)
assert diffusion.num_timesteps == 10
vol, x_T = bench.synth_volume(3)
cond_cpu = torch.cat([od.wavelet_pack(vol[:, k:k + 1]) for k in (1, 2, 3)], dim=1)
# the per-step noise the fused sampler draws: normal_() on the CUDA generator, same order
torch.manual_seed(77)
noises = [torch.randn(x_T.shape, device=dev) for _ in range(T)]
torch.manual_seed(77)
vd = vol.to(dev)
t0 = time.time()
steps_gpu = []
cond = pipeline.build_cond(vd[:, 1:2], vd[:, 2:3], vd[:, 3:4])
with torch.no_grad():
    for out in diffusion.p_sample_loop_progressive(model, tuple(x_T.shape), time=T, noise=x_T.to(dev), cond=cond, progress=False):
        steps_gpu.append(out["sample"].cpu())
img_gpu = ops.sample_to_image(steps_gpu[-1].to(dev), vd[:, 1:2]).cpu()[0, 0]
print(f"fcwdm: {time.time() - t0:.1f} s (incl. capture)", flush=True)

sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
tab = od.Tables(diffusion.betas)
net = lambda xin, tt: ow.wunet_forward(sd, xin, tt, model_channels=64, channel_mult=(1, 2, 2, 4))
img = x_T.clone()
print("| step | t | wavelet-domain rel-L2 | max-abs |\n|---|---|---|---|")
for k, i in enumerate(reversed(range(T))):
    t1 = time.time()
    with torch.no_grad():
        img = od.p_sample(tab, net, img, torch.tensor([i]), cond=cond_cpu, timestep_map=list(diffusion.timestep_map),
                          noise=noises[k].cpu())["sample"]
    rel = float((steps_gpu[k] - img).norm() / img.norm())
    print(f"| {k} | {i} | {rel:.3e} | {float((steps_gpu[k] - img).abs().max()):.3e} |   ({time.time() - t1:.1f} s)", flush=True)
ref = od.sample_postprocess(img, vol[:, 1:2])[0]
got = img_gpu[:, :, :155]
mse = float(((got.double() - ref.double()) ** 2).mean())
print(f"\nfinal image ({tuple(ref.shape)}): max-abs {float((got - ref).abs().max()):.3e}, PSNR {10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-30))):.1f} dB, "
      f"SSIM {ssim3d(got, ref):.5f}, mean |ref| {float(ref.abs().mean()):.4f}")
