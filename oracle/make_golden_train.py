"""TEST INFRASTRUCTURE ONLY -- training-step fixture from the UNMODIFIED reference (tests/golden/train_small.npz).

    CUDA_VISIBLE_DEVICES="" python -m oracle.make_golden_train

One call of the reference's GaussianDiffusion.training_losses (gaussian_diffusion.py:1084-1166, mode 'i2i') through
the reference's WavUNetModel (small configuration, seeded weights with the zero-initialised convs re-randomised),
loss = mean of mse_wav as in TrainLoop.forward_backward (train_util.py:447-460), loss.backward(), then one
torch.optim.AdamW step (train_util.py:75-82,391).  Stored: the batch, t, the image-space noise the reference drew,
mse_wav, the model output, per-parameter gradient norms, the gradients of all small parameters in full and a
slice of the large ones, and a few parameters after the optimizer step.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shims import reference_modules  # noqa: E402
from oracle import wunet as owunet              # noqa: E402
from oracle.make_golden import SMALL_CFG, GOLDEN  # noqa: E402

FULL_LIMIT = 4096      # parameters up to this many elements are stored in full
LR, WD = 1e-3, 0.01


def one_fixture(ref, model, sd_keys_tied, out_name):
    su = ref.script_util
    d10 = su.create_gaussian_diffusion(steps=10, predict_xstart=True, sample_schedule="sampled", mode="i2i")
    model.train()
    g = torch.Generator().manual_seed(21)
    batch = {k: torch.rand(2, 1, 16, 16, 16, generator=g) for k in ("t1n", "t1c", "t2w", "t2f")}
    t = torch.tensor([2, 8])
    # capture the noise training_losses draws internally (th.randn_like(target), :1143)
    drawn = {}
    real_randn_like = torch.randn_like

    def spy(x, *a, **k):
        out = real_randn_like(x, *a, **k)
        drawn.setdefault("noise", out.clone())
        return out

    torch.manual_seed(31)
    torch.randn_like = spy
    try:
        terms, mo, mo_idwt = d10.training_losses(model, batch, t, model_kwargs={}, mode="i2i", contr="t1n")
    finally:
        torch.randn_like = real_randn_like
    loss = (terms["mse_wav"] * torch.ones(8)).mean()
    opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD)
    loss.backward()
    out = {"t": t.numpy(), "noise": drawn["noise"].numpy(), "mse_wav": terms["mse_wav"].detach().numpy(),
           "loss": np.float64(loss.item()), "model_output": mo.detach().numpy(), "lr": LR, "wd": WD}
    for k, v in batch.items():
        out["batch_" + k] = v.numpy()
    names, norms = [], []
    for name, p in model.named_parameters():
        names.append(name)
        norms.append(float(p.grad.double().norm()))
        if p.numel() <= FULL_LIMIT:
            out["grad/" + name] = p.grad.numpy().copy()
        else:
            out["gradslice/" + name] = p.grad[:2].numpy().copy()
    out["param_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    opt.step()
    for name, p in model.named_parameters():
        if p.numel() <= 256:
            out["stepped/" + name] = p.detach().numpy().copy()
    path = os.path.join(GOLDEN, out_name)
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; loss", float(out["loss"]), "params", len(names))


def main():
    import importlib
    which = sys.argv[1:] or ["wunet", "unet", "wunet_ssn"]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with reference_modules() as ref:
        if "wunet" in which:
            model = ref.wunet.WavUNetModel(**SMALL_CFG)
            shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
            sd = owunet.tie_output_blocks(owunet.seeded_state_dict(shapes, seed=0), len(SMALL_CFG["channel_mult"]))
            model.load_state_dict(sd, strict=True)
            one_fixture(ref, model, sd, "train_small.npz")
        if "wunet_ssn" in which:                               # use_scale_shift_norm=True (wunet.py:256-260)
            cfg = dict(SMALL_CFG, use_scale_shift_norm=True)
            model = ref.wunet.WavUNetModel(**cfg)
            shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
            sd = owunet.tie_output_blocks(owunet.seeded_state_dict(shapes, seed=0), len(cfg["channel_mult"]))
            model.load_state_dict(sd, strict=True)
            one_fixture(ref, model, sd, "train_small_ssn.npz")
        if "unet" in which:
            from oracle.make_golden_unet import UNET_SMALL_CFG
            unet = importlib.import_module("guided_diffusion.unet")
            model = unet.UNetModel(**UNET_SMALL_CFG)
            shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
            model.load_state_dict(owunet.seeded_state_dict(shapes, seed=0), strict=True)
            model.to(torch.device("cpu"))                  # the reference's forward asserts x.device == self.devices[0]
            one_fixture(ref, model, None, "train_unet_small.npz")
        if "unet_ssn" in which:                                # the plain U-Net with use_scale_shift_norm=True (unet.py:297-309),
            from oracle.make_golden_unet import UNET_SMALL_CFG   # the default of script_util.model_and_diffusion_defaults()
            unet = importlib.import_module("guided_diffusion.unet")
            model = unet.UNetModel(**dict(UNET_SMALL_CFG, use_scale_shift_norm=True))
            shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
            model.load_state_dict(owunet.seeded_state_dict(shapes, seed=0), strict=True)
            model.to(torch.device("cpu"))
            one_fixture(ref, model, None, "train_unet_small_ssn.npz")


if __name__ == "__main__":
    main()
