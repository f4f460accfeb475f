"""The reference's own scripts/sample.py, UNMODIFIED (byte-for-byte copy staged by oracle/ref_shims.stage_reference, or
the read-only mount), executed against the drop-in packages: setup_dist -> create_model_and_diffusion(**args_to_dict) ->
load_state_dict -> BRATSVolumes / DataLoader -> DWT conditioning -> p_sample_loop -> IDWT / clamp / mask / crop ->
nib.save (scripts/sample.py:24-149).  The files it writes are compared with the oracle's chain on the same inputs.

The denoiser is the small 2-level configuration at the script's hard-wired 112x112x80 latent, with the output conv scaled
by 0.05 so that the fp32 network is a contraction and the 10-step chain can be compared tightly (see test_configs_gpu)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import diffusion as od
from oracle import ref_shims
from oracle import wunet as ow

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = 10


def _write_cases(root, n_cases):
    from fcwdm import nifti
    g = np.random.default_rng(0)
    names = []
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, 240), np.linspace(-1, 1, 240), np.linspace(-1, 1, 155), indexing="ij")
    brain = ((zz ** 2 + yy ** 2 + xx ** 2) < 0.7).astype(np.float32)
    for c in range(n_cases):
        name = f"BraTS-GLI-{c:05d}-000"                      # 19 characters: sample.py:61 slices [:19]
        os.makedirs(os.path.join(root, name))
        for k, seq in enumerate(("t1n", "t1c", "t2w", "t2f")):
            vol = (brain * (600.0 + 300.0 * np.sin(3 * zz + k + c) * np.cos(2 * yy - k) + 40.0 * g.standard_normal(brain.shape)))
            nifti.write(os.path.join(root, name, f"{name}-{seq}.nii.gz"), vol.astype(np.float32))
        names.append(name)
    return names


def test_reference_sample_script_runs_unchanged(tmp_path):
    script = ref_shims.reference_script("sample.py")
    if script is None:
        pytest.skip("no copy of the reference's scripts/sample.py (run __graft_entry__.build() where /root/reference is mounted)")
    from fcwdm import nifti
    from guided_diffusion.bratsloader import BRATSVolumes
    from guided_diffusion.script_util import create_gaussian_diffusion
    from guided_diffusion.wunet import WavUNetModel
    data = tmp_path / "validation"
    names = _write_cases(str(data), 2)
    cfg = dict(image_size=224, in_channels=32, model_channels=32, out_channels=8, num_res_blocks=2, attention_resolutions=(),
               channel_mult=(1, 2), dims=3, num_groups=32, bottleneck_attention=False, resblock_updown=True, use_freq=True)
    shapes = {k: tuple(v.shape) for k, v in WavUNetModel(**cfg).state_dict().items()}
    sd = ow.tie_output_blocks(ow.seeded_state_dict(shapes, seed=0), 2)
    sd["out.2.weight"] = sd["out.2.weight"] * 0.05
    torch.save(sd, tmp_path / "model.pt")
    rec = tmp_path / "rec"
    rec.mkdir()
    out = tmp_path / "out"
    flags = dict(data_dir=str(data), model_path=str(tmp_path / "model.pt"), output_dir=str(out), contr="t1n", seed=3,
                 batch_size=1, dataset="brats", image_size=224, num_channels=32, num_res_blocks=2, channel_mult="1,2",
                 in_channels=32, out_channels=8, dims=3, attention_resolutions="", bottleneck_attention=False,
                 resblock_updown=True, use_freq=True, use_scale_shift_norm=False, predict_xstart=True, diffusion_steps=T,
                 sample_schedule="sampled", mode="i2i", num_groups=32, num_heads=1, learn_sigma=False, dropout=0.0,
                 class_cond=False, clip_denoised=True)
    argv = [sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), script, str(rec)]
    argv += [f"--{k}={v}" for k, v in flags.items()]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    p = subprocess.run(argv, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "__SCRIPT_OK__ cases=2" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]

    # ---- replay through the oracle: same loader tensors, same x_T (recorded), same per-step noise (seeded per case/step)
    ds = BRATSVolumes(str(data), mode="eval")
    d = create_gaussian_diffusion(steps=T, predict_xstart=True, sample_schedule="sampled", mode="i2i")
    tab = od.Tables(d.betas)
    net = lambda xin, tt: ow.wunet_forward(sd, xin, tt, model_channels=32, channel_mult=(1, 2))
    torch.set_num_threads(os.cpu_count() or 1)
    for c, name in enumerate(names):
        item = ds[c]
        assert os.path.basename(os.path.dirname(item["subj"])) == name
        cond_imgs = [item[k][None] for k in ("t1c", "t2w", "t2f")]
        cond = torch.cat([od.wavelet_pack(v) for v in cond_imgs], dim=1)
        x = torch.load(rec / f"x_T_{c}.pt")
        assert tuple(x.shape) == (1, 8, 112, 112, 80)
        for i in reversed(range(T)):
            g = torch.Generator(device="cuda").manual_seed(7000 + 100 * c + i)
            nz = torch.empty(x.shape, device="cuda").normal_(generator=g).cpu()
            with torch.no_grad():
                x = od.p_sample(tab, net, x, torch.tensor([i]), cond=cond, timestep_map=list(d.timestep_map), noise=nz)["sample"]
        want = od.sample_postprocess(x, cond_imgs[0])[0].numpy()
        got = nifti.read(str(out / name / "sample.nii.gz"), dtype=np.float32)
        assert got.shape == want.shape == (224, 224, 155)
        mse = float(((got.astype(np.float64) - want) ** 2).mean())
        psnr = 10 * np.log10(1.0 / max(mse, 1e-30))
        print(f"{name}: sample.nii.gz vs oracle chain: max-abs {np.abs(got - want).max():.3e}, PSNR {psnr:.1f} dB")
        assert psnr >= 40.0 and np.abs(got - want).max() <= 0.1
        tgt = nifti.read(str(out / name / "target.nii.gz"), dtype=np.float32)
        np.testing.assert_array_equal(tgt, item["t1n"][0, :, :, :155].numpy())
