"""``TrainLoop`` for the fcwdm training path -- the reference's training driver (guided_diffusion/train_util.py:32-462)
re-hosted on the B200 kernels and made data-parallel.  Constructor keywords, method
names, checkpoint file names and the order of operations in a step are the reference's, so ``scripts/train.py`` runs
against this module unchanged; what is different, and why:

* **one process per GPU.**  ``dist_util.setup_dist`` honours the launcher; rank 0's weights are broadcast at start
  (``_load_and_sync_parameters``) and the backward all-reduces gradients in buckets over NCCL while it is still running
  (``fcwdm.ddp.attach``).  The reference trains a world of one (dist_util.py:41-43) with ``sync_params`` commented out.
* **optimizer** = ``fcwdm.optim.FusedAdamW``: one kernel over the flat fp32 master weights instead of
  ``torch.optim.AdamW``'s per-tensor loops (train_util.py:110); same update rule, same hyper-parameters.
* **no per-step host synchronisation.**  The reference reads back a loss ``.item()`` per step, eight ``mse_wav`` items
  and two ``.item()`` per PARAMETER for the norms (train_util.py:215-224,371-375,426-441).  Here the loss, the norms and
  a non-finite flag stay on the device and are read back when the logger dumps (every ``log_interval`` steps) or a
  checkpoint decision needs the number (every ``save_interval`` steps).
* ``use_fp16`` / GradScaler: the fcwdm path computes in bf16 with fp32 master weights and fp32 gradients, so there is
  no loss scale to manage; ``use_fp16=True`` is refused instead of silently meaning something else.
* wandb / tensorboard are optional (logged to when a run / writer exists), and the checkpoint root is
  ``$FCWDM_CHECKPOINT_ROOT``, else the reference's ``/data`` when that directory is writable, else the logger directory.
"""
import functools
import os
import re
import time

import numpy as np
import torch as th
import torch.distributed as dist

from . import dist_util, logger
from .resample import LossAwareSampler, UniformSampler
from DWT_IDWT.DWT_IDWT_layer import DWT_3D, IDWT_3D

try:                                                    # optional, as is tensorboard (a writer is passed in or not)
    import wandb as _wandb
except Exception:                                       # pragma: no cover - absent in the build image
    _wandb = None

INITIAL_LOG_LOSS_SCALE = 20.0
BAND_NAMES = ("LLL", "LLH", "LHL", "LHH", "HLL", "HLH", "HHL", "HHH")
MODALITIES = ("t1n", "t1c", "t2w", "t2f")


def visualize(img):
    """Min-max normalise an array to [0, 1] (all zeros for a constant image)."""
    lo, hi = img.min(), img.max()
    if hi > lo:
        return (img - lo) / (hi - lo)
    return np.zeros_like(img)


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_initialized() else 0


class TrainLoop:
    def __init__(self, *, model, diffusion, data, batch_size, in_channels, image_size, microbatch, lr, ema_rate,
                 log_interval, contr, save_interval, resume_checkpoint, resume_step, use_fp16=False,
                 fp16_scale_growth=1e-3, schedule_sampler=None, weight_decay=0.0, lr_anneal_steps=0, dataset='brats',
                 summary_writer=None, mode='default', loss_level='image', sample_schedule='direct',
                 diffusion_steps=1000):
        if use_fp16:
            raise NotImplementedError("use_fp16=True: the fcwdm path computes in bf16 with fp32 master weights and needs no "
                                      "loss scaling; pass use_fp16=False")
        if not hasattr(model, "train_engine"):
            raise TypeError("TrainLoop needs an fcwdm model (WavUNetModel / UNetModel from this package); there is no "
                            "PyTorch fallback on this path")
        self.summary_writer = summary_writer
        self.mode = mode
        self.model = model
        self.diffusion = diffusion
        self.datal = data
        self.dataset = dataset
        self.iterdatal = iter(data)
        self.batch_size = batch_size
        self.in_channels = in_channels
        self.image_size = image_size
        self.contr = contr
        self.microbatch = microbatch if microbatch > 0 else batch_size
        self.lr = lr
        self.ema_rate = [ema_rate] if isinstance(ema_rate, float) else [float(x) for x in str(ema_rate).split(",")]
        self.log_interval = log_interval
        self.save_interval = save_interval
        self.resume_checkpoint = resume_checkpoint
        self.use_fp16 = False
        self.schedule_sampler = schedule_sampler or UniformSampler(diffusion, diffusion.num_timesteps)
        self.weight_decay = weight_decay
        self.lr_anneal_steps = lr_anneal_steps
        self.dwt = DWT_3D('haar')
        self.idwt = IDWT_3D('haar')
        self.loss_level = loss_level
        self.step = 1
        self.resume_step = resume_step
        self.global_batch = self.batch_size * _world()
        self.sync_cuda = th.cuda.is_available()
        self.sample_schedule = sample_schedule
        self.diffusion_steps = diffusion_steps
        if not th.cuda.is_available():
            raise RuntimeError("Training requires CUDA: the fcwdm kernels have no CPU path")
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("move the model to its CUDA device before building TrainLoop (train.py:61 does)")
        diffusion.sync_timestep_check = False            # t is validated on the device (no host read-back per step)

        self.best_losses = {}
        self.best_checkpoints = {}
        self.checkpoint_dir = os.path.join(get_blob_logdir(), 'checkpoints')
        if _rank() == 0:
            os.makedirs(self.checkpoint_dir, exist_ok=True)
        self._load_best_losses()

        self._load_and_sync_parameters()
        from fcwdm import ddp
        from fcwdm.optim import FusedAdamW
        self.opt = FusedAdamW(self.model, lr=self.lr, weight_decay=self.weight_decay)
        self.grad_sync = ddp.attach(self.model) if _world() > 1 else None
        if self.resume_step or self.resume_checkpoint or find_resume_checkpoint():
            logger.log(f"Resume Step: {self.resume_step}")
            self._load_optimizer_state()         # also for a BEST file resumed at step 0 (the reference skips it then)
        self._ones = th.ones(8, device=self.device)
        self._nonfinite = th.zeros((), dtype=th.bool, device=self.device)
        self.last_info = {}

    # ------------------------------------------------------------------------------------------ best-loss bookkeeping
    def _best_file(self):
        return os.path.join(self.checkpoint_dir, 'best_losses.txt')

    def _load_best_losses(self):
        self.best_losses = {}
        try:
            with open(self._best_file()) as f:
                for line in f:
                    if line.strip():
                        modality, value = line.strip().split(':')
                        self.best_losses[modality] = float(value)
            logger.log(f"Loaded best losses: {self.best_losses}")
        except FileNotFoundError:
            pass
        except Exception as exc:                          # a damaged file must not stop training (reference :128-130)
            logger.warn(f"Error loading best losses: {exc}")
            self.best_losses = {}

    def _save_best_losses(self):
        try:
            with open(self._best_file(), 'w') as f:
                for modality, value in self.best_losses.items():
                    f.write(f"{modality}:{value}\n")
        except Exception as exc:
            logger.warn(f"Error saving best losses: {exc}")

    # ------------------------------------------------------------------------------------------ parameters / optimizer
    def _load_and_sync_parameters(self):
        resume_checkpoint = find_resume_checkpoint() or self.resume_checkpoint
        if resume_checkpoint:
            self.resume_step = resume_step_of_checkpoint(resume_checkpoint, self.resume_step)
            logger.log(f"loading model from checkpoint: {resume_checkpoint}...")
            self.model.load_state_dict(dist_util.load_state_dict(resume_checkpoint, map_location=self.device))
        dist_util.sync_params(self.model.parameters())
        self._decorrelate_ranks()
        for name in ("_engine", "_train_engine"):        # packed weights follow the new values
            eng = getattr(self.model, name, None)
            if eng is not None:
                eng.invalidate()

    def _decorrelate_ranks(self):
        """scripts/train.py seeds torch / numpy / random with the SAME ``args.seed`` in every process (train.py:26-29) and
        builds a shuffling DataLoader without a DistributedSampler, so under torchrun every rank would draw the same
        cases, timesteps and noise and the all-reduce would average identical gradients.  After the weight broadcast the
        generators are therefore advanced by a rank-dependent offset (seed + rank); rank 0 keeps
        the single-process stream, so a world of one is unchanged."""
        world, rank = _world(), _rank()
        if world <= 1:
            return
        if rank:
            import random
            th.manual_seed(th.initial_seed() + rank)
            th.cuda.manual_seed(th.cuda.initial_seed() + rank)
            np.random.seed((int(np.random.get_state()[1][0]) + rank) % (2 ** 32))
            random.seed(random.getrandbits(32) + rank)
        sampler = getattr(self.datal, "sampler", None)
        from torch.utils.data.distributed import DistributedSampler
        if sampler is not None and not isinstance(sampler, DistributedSampler) and rank == 0:
            logger.warn("the data loader has no DistributedSampler: ranks draw independently shuffled cases (re-seeded "
                        "per rank), so one epoch visits each case ~world_size times instead of once")

    def _load_optimizer_state(self):
        main_checkpoint = find_resume_checkpoint() or self.resume_checkpoint
        folder = os.path.dirname(main_checkpoint) if main_checkpoint else self.checkpoint_dir
        for name in (f"opt{self.resume_step:06}.pt", f"opt_best_{self.contr}.pt"):
            path = os.path.join(folder, name)
            if os.path.exists(path):
                logger.log(f"loading optimizer state from checkpoint: {path}")
                self.opt.load_state_dict(dist_util.load_state_dict(path, map_location=self.device))
                return
        logger.log('no optimizer checkpoint exists')

    # ------------------------------------------------------------------------------------------ the loop
    def _next_batch(self):
        try:
            return next(self.iterdatal)
        except StopIteration:
            self.iterdatal = iter(self.datal)
            return next(self.iterdatal)

    def _to_device(self, batch):
        if self.mode == 'i2i':
            for k in MODALITIES:
                batch[k] = batch[k].to(self.device, non_blocking=True)
            return batch
        return batch.to(self.device, non_blocking=True)

    def run_loop(self):
        t_data = t_step = t_log = t_save = 0.0
        start = last = time.time()
        lossmse = None
        # The host issues ~700 launches per step and stays about one step ahead of the GPU; a full (generation-2) Python
        # garbage collection walks every object torch / numpy / the model tree ever created -- 32 ms here, a whole step --
        # and the GPU drains meanwhile (measured: one 65 ms step in ~25, tools/train_stall_probe.py).  Everything alive now
        # is long-lived, so it is moved out of the collector's reach; the per-step garbage (tape closures) stays collectable.
        import gc
        gc.collect()
        gc.freeze()
        while not self.lr_anneal_steps or self.step + self.resume_step < self.lr_anneal_steps:
            now = time.time()
            t_total, last = now - last, now
            batch = self._next_batch()
            cond = {}
            batch = self._to_device(batch)
            t_data += time.time() - now

            t0 = time.time()
            lossmse, sample, sample_idwt = self.run_step(batch, cond)
            t_step += time.time() - t0

            t0 = time.time()
            gstep = self.step + self.resume_step
            if self.step % self.log_interval == 0:
                scalars = {'time/load': t_data, 'time/forward': t_step, 'time/total': t_total, 'loss/MSE': float(lossmse)}
                self._log_scalars(scalars, gstep)
            if self.step % 200 == 0:
                self._log_images(batch, sample, sample_idwt, gstep)
            if self.step % self.log_interval == 0:
                if bool(self._nonfinite):
                    logger.warn("a non-finite loss was seen since the last dump")
                    self._nonfinite.zero_()
                logger.dumpkvs()
            t_log += time.time() - t0

            if self.step % self.save_interval == 0:
                t0 = time.time()
                self.save_if_best(float(lossmse))
                t_save += time.time() - t0
                if os.environ.get("DIFFUSION_TRAINING_TEST", "") and self.step > 0:
                    return
            self.step += 1
            if self.step % self.log_interval == 0:
                logger.log(f"[PROFILE] Step {self.step}: Data {t_data:.2f}s, Step {t_step:.2f}s, Log {t_log:.2f}s, "
                           f"Save {t_save:.2f}s, Total {time.time() - start:.2f}s")
                t_data = t_step = t_log = t_save = 0.0
        if lossmse is not None and (self.step - 1) % self.save_interval != 0:
            self.save_if_best(float(lossmse))

    def _log_scalars(self, scalars, gstep):
        if self.summary_writer is not None:
            for k, v in scalars.items():
                self.summary_writer.add_scalar(k, v, global_step=gstep)
        if _wandb is not None and getattr(_wandb, "run", None) is not None and _rank() == 0:
            _wandb.log(dict(scalars, step=gstep), step=gstep)

    def _log_images(self, batch, sample, sample_idwt, gstep):
        """Mid-plane previews of the prediction, its eight bands and the conditioning modalities (reference :226-271)."""
        use_wandb = _wandb is not None and getattr(_wandb, "run", None) is not None and _rank() == 0
        if self.summary_writer is None and not use_wandb:
            return
        planes = {'sample/x_0': sample_idwt[0, 0, :, :, sample_idwt.size(2) // 2]}
        mid = sample.size(2) // 2
        for ch, name in enumerate(BAND_NAMES):
            planes[f'sample/{name}'] = sample[0, ch, :, :, mid]
        if self.mode == 'i2i':
            for k in MODALITIES:
                if k != self.contr:
                    planes[f'source/{k}'] = batch[k][0, 0, :, :, batch[k].size(2) // 2]
        images = {}
        for key, plane in planes.items():
            plane = plane.detach().float()
            if self.summary_writer is not None:
                self.summary_writer.add_image(key, plane.unsqueeze(0), global_step=gstep)
            if use_wandb:
                img = (visualize(plane.cpu().numpy()) * 255).astype('uint8')
                images[key] = _wandb.Image(img, caption=key)
        if use_wandb:
            _wandb.log(images, step=gstep)

    # ------------------------------------------------------------------------------------------ one step
    def run_step(self, batch, cond, label=None, info=None):
        info = {} if info is None else info
        lossmse, sample, sample_idwt = self.forward_backward(batch, cond, label)
        with th.no_grad():                               # two reductions over the flat buffers; results stay on the GPU
            info['norm/param_max'] = self.opt.flat.abs().max()
            flat_grad = getattr(self.model.train_engine(), "last_flat", None)
            if flat_grad is not None:
                info['norm/grad_max'] = flat_grad.abs().max()
            self._nonfinite |= ~th.isfinite(lossmse)
        self.last_info = info
        self.opt.step()
        self._anneal_lr()
        self.log_step()
        return lossmse, sample, sample_idwt

    def forward_backward(self, batch, cond, label=None):
        self.opt.zero_grad()
        batch_size = batch['t1n'].shape[0] if self.mode == 'i2i' else batch.shape[0]
        t, _ = self.schedule_sampler.sample(batch_size, self.device)
        compute_losses = functools.partial(self.diffusion.training_losses, self.model, x_start=batch, t=t,
                                           model_kwargs=cond, labels=label, mode=self.mode, contr=self.contr)
        losses, sample, sample_idwt = compute_losses()
        if isinstance(self.schedule_sampler, LossAwareSampler):
            # training_losses reports one batch-mean value per band, not one loss per sample (gaussian_diffusion.py
            # :1160); the reference indexes its return tuple with a string here and cannot run either
            raise NotImplementedError("loss-aware timestep sampling needs per-sample losses; training_losses returns "
                                      "per-band means -- use schedule_sampler='uniform'")
        mse_wav = losses["mse_wav"]
        loss = (mse_wav * self._ones).mean()             # all bands weighted equally (reference :447-449)
        lossmse = loss.detach()
        for i, name in enumerate(BAND_NAMES):
            logger.logkv_mean(f"mse_wav_{name.lower()}", mse_wav[i].detach())
        logger.logkv_mean("loss", lossmse)
        # the reference's log_loss_dict columns (train_util.py:554-560): the band mean under "mse_wav", and -- it zips the B
        # timesteps with the 8 band values -- band i under the quartile of t[i]; the timesteps come from the sampler's host
        # copy of its draw, the values stay on the device until the logger dumps
        logger.logkv_mean("mse_wav", lossmse)
        t_host = getattr(self.schedule_sampler, "last_indices", None)
        if t_host is not None:
            for i, ti in enumerate(t_host[:len(BAND_NAMES)]):
                logger.logkv_mean(f"mse_wav_q{int(4 * int(ti) / self.diffusion.num_timesteps)}", mse_wav[i].detach())
        loss.backward()
        return lossmse, sample, sample_idwt

    def _anneal_lr(self):
        if not self.lr_anneal_steps:
            return
        frac_done = (self.step + self.resume_step) / self.lr_anneal_steps
        for group in self.opt.param_groups:
            group["lr"] = self.lr * (1 - frac_done)

    def log_step(self):
        logger.logkv("step", self.step + self.resume_step)
        logger.logkv("samples", (self.step + self.resume_step + 1) * self.global_batch)
        for k, v in self.last_info.items():
            logger.logkv(k, v)

    # ------------------------------------------------------------------------------------------ checkpoints
    def save_if_best(self, current_loss):
        """Keep one checkpoint per target modality: the one with the lowest loss seen at a save point."""
        modality = self.contr
        if _world() > 1:                                  # every rank must take the same decision
            box = th.tensor([current_loss], device=self.device, dtype=th.float64)
            dist.all_reduce(box, op=dist.ReduceOp.SUM)
            current_loss = float(box.item()) / _world()
        best = self.best_losses.get(modality)
        if best is not None and not current_loss < best:
            logger.log(f"Loss {current_loss:.6f} not better than best {best:.6f} for {modality}")
            return False
        self.best_losses[modality] = current_loss
        if _rank() != 0:
            return True
        logger.log(f"NEW BEST for {modality}! Loss: {current_loss:.6f}")
        old = self.best_checkpoints.get(modality)
        path = os.path.join(self.checkpoint_dir, f"brats_{modality}_BEST_{self.sample_schedule}_{self.diffusion_steps}.pt")
        if old and old != path and os.path.exists(old):
            os.remove(old)
        try:
            _atomic_save(self.model.state_dict(), path)
            self.best_checkpoints[modality] = path
            self._save_best_losses()
            _atomic_save(self.opt.state_dict(), os.path.join(self.checkpoint_dir, f"opt_best_{modality}.pt"))
            logger.log(f"Saved new best checkpoint: {path}")
        except Exception as exc:
            logger.error(f"Error saving checkpoint: {exc}")
        return True

    def save(self):
        """Step-numbered checkpoint + optimizer state (the reference's legacy ``save``, :472-513)."""
        if self.dataset not in ('brats', 'lidc-idri', 'brats_inpainting', 'synthrad'):
            raise ValueError(f'dataset {self.dataset} not implemented')
        if _rank() != 0:
            return
        gstep = self.step + self.resume_step
        name = f"{self.dataset}_{self.contr}_{gstep:06d}_{self.sample_schedule}_{self.diffusion_steps}.pt"
        path = os.path.join(self.checkpoint_dir, name)
        logger.log(f"Saving model to: {path}")
        _atomic_save(self.model.state_dict(), path)
        _atomic_save(self.opt.state_dict(), os.path.join(self.checkpoint_dir, f"opt{gstep:06d}.pt"))


def _atomic_save(obj, path):
    tmp = path + ".tmp"
    th.save(obj, tmp)
    os.replace(tmp, path)


def parse_resume_step_from_filename(filename):
    """Step count encoded in a checkpoint name: the digits that end the last '_'-separated word of the stem
    (``.../brats_t1n_005000.pt`` -> 5000, ``model012000.pt`` -> 12000); 0 when there are none."""
    stem = os.path.basename(filename).split(".")[-2] if "." in os.path.basename(filename) else os.path.basename(filename)
    word = stem.split("_")[-1]
    digits = ""
    for c in reversed(word):
        if not c.isdigit():
            break
        digits = c + digits
    return int(digits) if digits else 0


_STEP_FIELD = re.compile(r"_(\d{6})(?:_|\.|$)")


def resume_step_of_checkpoint(filename, fallback=0):
    """Global step a checkpoint was written at, for the names THIS class writes: ``<dataset>_<contr>_<NNNNNN>_<schedule>_
    <T>.pt`` (``save``) carries it as its 6-digit field; ``..._BEST_<schedule>_<T>.pt`` (``save_if_best``) carries none,
    so the caller's ``resume_step`` stands.  ``parse_resume_step_from_filename`` -- kept as the reference defines it --
    would return the diffusion step count T for both (their last '_' word); it is only used for other names
    (``model012000.pt``)."""
    base = os.path.basename(filename)
    m = _STEP_FIELD.search(base)
    if m:
        return int(m.group(1))
    if "_BEST_" in base:
        return fallback
    return parse_resume_step_from_filename(filename) or fallback


def get_blob_logdir():
    """Root under which ``checkpoints/`` lives: $FCWDM_CHECKPOINT_ROOT, else the reference's ``/data`` volume when it is
    there and writable, else the logger's directory."""
    root = os.environ.get("FCWDM_CHECKPOINT_ROOT")
    if root:
        return root
    if os.path.isdir("/data") and os.access("/data", os.W_OK):
        return "/data"
    return logger.get_dir()


def find_resume_checkpoint():
    """Hook for infrastructure that can discover the newest checkpoint on its own; none by default."""
    return None


def log_loss_dict(diffusion, ts, losses):
    """Mean of every loss term plus per-quartile-of-t means (reference :553-559).  `ts` / values may live on the GPU:
    they are read back once here, so call it at logging cadence, not per step."""
    ts_host = ts.detach().cpu().numpy()
    for key, values in losses.items():
        vals = values.detach().float().cpu().numpy()
        logger.logkv_mean(key, float(vals.mean()))
        if vals.shape[:1] == ts_host.shape[:1]:
            for sub_t, sub_loss in zip(ts_host, vals):
                quartile = int(4 * sub_t / diffusion.num_timesteps)
                logger.logkv_mean(f"{key}_q{quartile}", float(np.mean(sub_loss)))
