"""TEST INFRASTRUCTURE ONLY -- three real steps of the UNMODIFIED reference ``TrainLoop`` (guided_diffusion/train_util.py:
32-462) on the small wavelet U-Net, frozen as tests/golden/trainloop_steps.npz:

    CUDA_VISIBLE_DEVICES="" python -m oracle.make_golden_trainloop_steps

What is recorded: the timesteps the loop drew, the image-space noise ``training_losses`` drew (so the GPU test can replay
it), the per-step loss and eight per-band losses from the reference's own progress.csv, the names of the files it left in
its checkpoint directory, best_losses.txt, and norms / slices of the parameters after the three AdamW steps.

The reference cannot run this loop on a CPU-only box as shipped: it calls ``.cuda()`` on the loss weights
(train_util.py:447), writes to ``/data`` (:545) and through ``blobfile``, and logs to wandb.  The harness works around
that WITHOUT touching the reference's files: ``Tensor.cuda`` is a no-op for the duration, ``get_blob_logdir`` is pointed at
a scratch directory inside the repo, ``blobfile`` is a three-function stand-in over the local file system, wandb is
initialised in disabled mode."""
import csv
import importlib
import os
import shutil
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import wunet as owunet                    # noqa: E402
from oracle.make_golden import GOLDEN, SMALL_CFG      # noqa: E402
from oracle.ref_shims import reference_modules        # noqa: E402

KEYS = ("t1n", "t1c", "t2w", "t2f")
STEPS = 4            # run_loop stops when step + resume_step reaches lr_anneal_steps: steps 1, 2, 3


class Volumes(torch.utils.data.Dataset):
    """The dataset of tests/test_trainloop_gpu.py: BRATSVolumes-shaped items of seeded random 16^3 volumes."""

    def __init__(self, n=4, seed=5):
        g = torch.Generator().manual_seed(seed)
        self.items = [{k: torch.rand(1, 16, 16, 16, generator=g) for k in KEYS} for _ in range(n)]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return dict(self.items[i], missing="none", subj="dummy_string")


def _blobfile_stub():
    bf = types.ModuleType("blobfile")
    bf.BlobFile = lambda path, mode="rb": open(path, mode)
    bf.join, bf.dirname, bf.exists = os.path.join, os.path.dirname, os.path.exists
    return bf


def main():
    scratch = os.path.join(ROOT, "gpurun_out", "_trainloop_golden")
    shutil.rmtree(scratch, ignore_errors=True)
    os.makedirs(scratch)
    os.environ["WANDB_MODE"] = "disabled"
    os.environ["WANDB_SILENT"] = "true"
    sys.modules["blobfile"] = _blobfile_stub()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    import wandb
    wandb.init(mode="disabled")
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # train_util.py:447 hard-codes .cuda() on the loss weights
    noises, ts = [], []
    real_randn_like = torch.randn_like
    try:
        with reference_modules() as ref:
            tu = importlib.import_module("guided_diffusion.train_util")
            dist_util = importlib.import_module("guided_diffusion.dist_util")
            logger = importlib.import_module("guided_diffusion.logger")
            resample = importlib.import_module("guided_diffusion.resample")
            tu.get_blob_logdir = lambda: scratch                    # train_util.py:545 returns "/data"
            dist_util.setup_dist()
            logger.configure(dir=os.path.join(scratch, "log"), format_strs=["csv"])
            model = ref.wunet.WavUNetModel(**SMALL_CFG)
            shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
            sd = owunet.tie_output_blocks(owunet.seeded_state_dict(shapes, seed=0), len(SMALL_CFG["channel_mult"]))
            model.load_state_dict(sd, strict=True)
            model.train()
            d10 = ref.script_util.create_gaussian_diffusion(steps=10, predict_xstart=True, sample_schedule="sampled", mode="i2i")
            sampler = resample.create_named_schedule_sampler("uniform", d10, maxt=d10.num_timesteps)
            orig_sample = sampler.sample

            def sample(batch_size, device):
                t, w = orig_sample(batch_size, device)
                ts.append(t.clone())
                return t, w

            sampler.sample = sample

            def randn_like(x, *a, **k):
                z = real_randn_like(x, *a, **k)
                noises.append(z.clone())
                return z

            torch.randn_like = randn_like
            np.random.seed(0)
            torch.manual_seed(0)
            data = torch.utils.data.DataLoader(Volumes(), batch_size=2, shuffle=False)
            loop = tu.TrainLoop(model=model, diffusion=d10, data=data, batch_size=2, in_channels=32, image_size=16,
                                microbatch=-1, lr=1e-3, ema_rate="0.9999", log_interval=1, contr="t1n", save_interval=2,
                                resume_checkpoint="", resume_step=0, use_fp16=False, schedule_sampler=sampler,
                                weight_decay=0.01, lr_anneal_steps=STEPS, dataset="brats", summary_writer=None, mode="i2i",
                                sample_schedule="sampled", diffusion_steps=10)
            loop.run_loop()
            final = {k: v.detach().clone() for k, v in model.state_dict().items()}
            opt_state = loop.opt.state_dict()
    finally:
        torch.randn_like = real_randn_like
        torch.Tensor.cuda = real_cuda
    rows = list(csv.DictReader(open(os.path.join(scratch, "log", "progress.csv"))))
    ckpt = sorted(os.listdir(os.path.join(scratch, "checkpoints")))
    out = {
        "t": np.stack([t.numpy() for t in ts]), "noise": np.stack([z.numpy() for z in noises]),
        # log_loss_dict (train_util.py:554-560) logs the mean over the 8 equally weighted bands under "mse_wav" = the loss
        "loss": np.array([float(r["mse_wav"]) for r in rows]),
        "csv_columns": np.array(sorted(rows[0].keys())),
        "step": np.array([int(float(r["step"])) for r in rows]), "samples": np.array([int(float(r["samples"])) for r in rows]),
        "checkpoint_files": np.array(ckpt), "best_losses_txt": np.array(open(os.path.join(scratch, "checkpoints", "best_losses.txt")).read()),
        "opt_state_keys": np.array(sorted(opt_state.keys())), "opt_step": np.array(float(opt_state["state"][0]["step"])),
        "final_lr": np.array(loop.opt.param_groups[0]["lr"]),
        "param_names": np.array(list(final.keys())),
        "param_norms": np.array([float(v.double().norm()) for v in final.values()]),
        "param_delta_norms": np.array([float((v.double() - sd[k].double()).norm()) for k, v in final.items()]),
    }
    for k in ("out.2.bias", "time_embed.0.bias", "input_blocks.0.0.bias", "middle_block.0.in_layers.0.weight"):
        out["param/" + k] = final[k].numpy()
        out["delta/" + k] = (final[k] - sd[k]).numpy()
    np.savez(os.path.join(GOLDEN, "trainloop_steps.npz"), **out)
    print("steps", out["step"], "loss", out["loss"], "t", out["t"].tolist(), "files", ckpt)
    print(out["best_losses_txt"])
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
