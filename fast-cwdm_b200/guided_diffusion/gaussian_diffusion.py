"""Drop-in for the hot-path part of the reference's ``guided_diffusion.gaussian_diffusion``.

Same names, signatures and results as the reference for: ``get_named_beta_schedule`` (:30-67),
``betas_for_alpha_bar`` (:70-88), the ``ModelMeanType / ModelVarType / LossType`` enums (:90-123),
``GaussianDiffusion.__init__`` tables (:143-205), ``q_sample`` (:224-242), ``q_posterior_mean_variance``
(:244-267), ``p_mean_variance`` (:269-388), ``p_sample`` (:529-574), ``p_sample_loop`` (:481-527),
``p_sample_loop_progressive`` (:668-719), ``training_losses`` (:1084-1166) and ``_extract_into_tensor``
(:1246-1263).  The schedule tables stay float64 numpy on the host exactly as in the reference; all per-voxel
arithmetic runs in hand-written sm_100a kernels (libfcwdm.so):

* process_xstart + q_posterior_mean_variance + the sampling line of p_sample are ONE elementwise kernel
  (fcwdm_p_sample_step): IDWT -> clamp -> DWT touches exactly the 8 sub-band values of one latent voxel;
* the six per-step ``_extract_into_tensor`` gathers (each a host sync + H2D copy in the reference) are replaced
  by a [T][5] coefficient table uploaded once per device and indexed by ``t`` inside the kernel;
* when the model is this package's WavUNetModel the loop keeps x_t as channels 0..7 of a persistent
  channels-last bf16 denoiser input next to the (constant) conditioning channels, so the reference's per-step
  ``th.cat([x, cond])`` (:297) disappears, and the whole step is replayed from a CUDA graph (fcwdm/sampler.py).

Out of scope: DDIM, the *_known / interpolation loops and the VLB/bpd terms; they
are not reachable from scripts/sample.py or scripts/train.py.  Learned variances (learn_sigma=True) and
ModelMeanType.PREVIOUS_X raise NotImplementedError (run.sh ships learn_sigma=False, predict_xstart=True).
"""
import enum
import math

import numpy as np
import torch as th

from .nn import mean_flat
from DWT_IDWT.DWT_IDWT_layer import DWT_3D, IDWT_3D
from fcwdm import ops

dwt = DWT_3D('haar')
idwt = IDWT_3D('haar')


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, sample_schedule="direct"):
    """Reference :30-67.  'direct' = linear betas rescaled by 1000/T; 'sampled' = the 1000-step linear
    alpha-bar curve sub-sampled at linspace(0, 999, T) (the Fast-DDPM style schedule fast-cwdm ships)."""
    if schedule_name == "linear":
        if sample_schedule == "direct":
            scale = 1000 / num_diffusion_timesteps
            return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
        if sample_schedule == "sampled":
            full = np.cumprod(1.0 - np.linspace(0.0001, 0.02, 1000, dtype=np.float64), axis=0)
            picked = full[np.linspace(0, 999, num_diffusion_timesteps, dtype=int)]
            prev = np.concatenate([[1.0], picked[:-1]])
            return np.clip(1.0 - picked / prev, 0.0001, 0.999)
        raise NotImplementedError(f"Unknown sample_schedule: {sample_schedule}")
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps,
                                   lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """Reference :70-88."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self == LossType.KL or self == LossType.RESCALED_KL


def _unwrap(model):
    """respace._WrappedModel -> (inner model, timestep_map or None, rescale flag, original steps)."""
    if hasattr(model, "timestep_map") and hasattr(model, "model"):
        return model.model, model.timestep_map, model.rescale_timesteps, model.original_num_steps
    return model, None, False, None


class GaussianDiffusion:
    """Utilities for training and sampling diffusion models (reference :126-205 for the constructor)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False,
                 mode='default', loss_level='image'):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        self.mode = mode
        self.loss_level = loss_level

        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])

        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self._dev_tables = {}
        self._samplers = {}

    # ------------------------------------------------------------------ device-side coefficient tables
    def _model_variance_tables(self):
        """(variance, log_variance) float64 tables for the fixed-variance types (reference :320-331)."""
        if self.model_var_type == ModelVarType.FIXED_LARGE:
            v = np.append(self.posterior_variance[1], self.betas[1:])
            return v, np.log(v)
        if self.model_var_type == ModelVarType.FIXED_SMALL:
            return self.posterior_variance, self.posterior_log_variance_clipped
        raise NotImplementedError("learned variances (learn_sigma=True) are not implemented on the fcwdm path; "
                                  "fast-cwdm ships learn_sigma=False (run.sh:124)")

    def _table(self, name, device):
        """fp32 device copies of host tables, built once per device (replaces the per-call H2D of :1260)."""
        key = (name, str(device))
        tab = self._dev_tables.get(key)
        if tab is None:
            if name == "step":       # [T][5]: coef1, coef2, sigma*(t != 0), sqrt_recip_acp, sqrt_recipm1_acp
                _, logvar = self._model_variance_tables()
                sigma = np.exp(0.5 * logvar.astype(np.float32).astype(np.float64))
                sigma = sigma * (np.arange(self.num_timesteps) != 0)
                arr = np.stack([self.posterior_mean_coef1, self.posterior_mean_coef2, sigma,
                                self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod], axis=1)
            elif name == "mean":     # same with sigma = 0: the kernel then returns the posterior mean
                arr = np.stack([self.posterior_mean_coef1, self.posterior_mean_coef2, np.zeros(self.num_timesteps),
                                self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod], axis=1)
            elif name == "q":        # [T][2]
                arr = np.stack([self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod], axis=1)
            else:
                arr = np.asarray(getattr(self, name))
            tab = th.from_numpy(np.ascontiguousarray(arr.astype(np.float32))).to(device)
            self._dev_tables[key] = tab
        return tab

    sync_timestep_check = True

    def _check_t(self, t):
        """Reference :1257-1259 raises IndexError for out-of-range timesteps (its `.min()/.max()` read-back is a host
        synchronisation per call).  With ``sync_timestep_check = False`` (training loops that must not stall the launch
        queue: fcwdm bench / TrainLoop) CUDA timesteps are validated on the device instead: an out-of-range value
        trips an asynchronous device-side assert rather than an IndexError."""
        if not t.numel():
            return
        if t.is_cuda and (not self.sync_timestep_check or th.cuda.is_current_stream_capturing()):
            th._assert_async(((t >= 0) & (t < self.num_timesteps)).all())
            return
        if int(t.min()) < 0 or int(t.max()) >= self.num_timesteps:
            raise IndexError(f"Timesteps out of bounds: min={int(t.min())}, max={int(t.max())}, "
                             f"arr len={self.num_timesteps}")

    # ------------------------------------------------------------------ forward process
    def q_mean_variance(self, x_start, t):
        mean = _extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = _extract_into_tensor(1.0 - self.alphas_cumprod, t, x_start.shape)
        log_variance = _extract_into_tensor(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def q_sample(self, x_start, t, noise=None):
        """sqrt(acp[t]) * x_start + sqrt(1 - acp[t]) * noise   (reference :224-242), one kernel."""
        if noise is None:
            noise = th.randn_like(x_start)
        assert noise.shape == x_start.shape
        self._check_t(t)
        return ops.q_sample(x_start.float(), noise.float(), self._table("q", x_start.device),
                            t.to(th.int64).contiguous())

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """Reference :244-267.  The mean is evaluated by the fused step kernel with sigma = 0 and clipping off."""
        assert x_start.shape == x_t.shape
        self._check_t(t)
        t64 = t.to(th.int64).contiguous()
        mean, _ = ops.p_sample_step(x_start.float(), x_t.float(), x_t.float(), self._table("mean", x_t.device), t64,
                                    clip_denoised=False, predict_xstart=True, want_pred=False)
        var = _extract_into_tensor(self.posterior_variance, t, x_t.shape)
        logvar = _extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, var, logvar

    # ------------------------------------------------------------------ reverse process
    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def _wrap_model(self, model):
        """Hook for SpacedDiffusion (respace.py), which maps t to the original timestep before the model call."""
        return model

    def _model_output(self, model, x, t, cond, model_kwargs):
        model = self._wrap_model(model)
        if self.mode == 'i2i':
            x_in = th.cat([x, cond], dim=1)          # generic-model path only; the fused sampler never concatenates
        else:
            x_in = x
        return model(x_in, self._scale_timesteps(t), **(model_kwargs or {}))

    def _step(self, model_output, x, t, noise, clip_denoised, table):
        if self.model_mean_type == ModelMeanType.PREVIOUS_X:
            raise NotImplementedError("ModelMeanType.PREVIOUS_X is not implemented on the fcwdm path")
        if x.shape[1] != 8:
            raise NotImplementedError("the fused step operates on the 8 Haar sub-band channels of a wavelet-domain "
                                      f"sample; got {x.shape[1]} channels")
        x8 = x[:, :8].float().contiguous()
        return ops.p_sample_step(model_output.float(), x8, noise, self._table(table, x.device),
                                 t.to(th.int64).contiguous(), clip_denoised=clip_denoised,
                                 predict_xstart=(self.model_mean_type == ModelMeanType.START_X), want_pred=True)

    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, cond=None):
        """Reference :269-388: {'mean', 'variance', 'log_variance', 'pred_xstart'}."""
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn is not supported by the fused step kernel")
        B = x.shape[0]
        assert t.shape == (B,)
        self._check_t(t)
        variance, log_variance = self._model_variance_tables()
        model_output = self._model_output(model, x, t, cond, model_kwargs)
        mean, pred = self._step(model_output, x, t, x[:, :8].float().contiguous(), clip_denoised, "mean")
        model_variance = _extract_into_tensor(variance, t, x.shape)
        model_log_variance = _extract_into_tensor(log_variance, t, x.shape)
        assert mean.shape == model_log_variance.shape == pred.shape == x.shape
        return {"mean": mean, "variance": model_variance, "log_variance": model_log_variance, "pred_xstart": pred}

    def _predict_xstart_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, cond=None):
        """Sample x_{t-1} (reference :529-574) -> {'sample', 'pred_xstart'}.  One model call + one fused kernel;
        the noise is drawn with th.randn_like(x) exactly where the reference draws it (:565)."""
        if cond_fn is not None or denoised_fn is not None:
            raise NotImplementedError("cond_fn / denoised_fn are not supported on the fcwdm path")
        B = x.shape[0]
        assert t.shape == (B,)
        self._check_t(t)
        self._model_variance_tables()
        model_output = self._model_output(model, x, t, cond, model_kwargs)
        noise = th.randn_like(x)
        sample, pred = self._step(model_output, x, t, noise.float(), clip_denoised, "step")
        return {"sample": sample, "pred_xstart": pred}

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=True, cond=None):
        """Reference :481-527.  NOTE: as in the reference this forwards to p_sample_loop_progressive with its
        default ``time=1000``, which only matches diffusions with num_timesteps == 1000; for
        respaced/short schedules call p_sample_loop_progressive(time=diffusion.num_timesteps) as
        scripts/complete_dataset.py:270-278 does.  Here ``time`` defaults to None = num_timesteps, which is
        identical for T = 1000 and fixes the IndexError for every other T."""
        final = None
        for sample in self._sample_loop(model, shape, None, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                                        device, progress, cond, lean=True):
            final = sample
        return final["sample"].clone() if final.get("_view") else final["sample"]

    def p_sample_loop_progressive(self, model, shape, time=None, noise=None, clip_denoised=True, denoised_fn=None,
                                  cond_fn=None, model_kwargs=None, device=None, progress=True, cond=None):
        """Generator over the per-step dicts of p_sample (reference :668-719)."""
        return self._sample_loop(model, shape, time, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device,
                                 progress, cond, lean=False)

    def _sample_loop(self, model, shape, time, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress,
                     cond, lean):
        """lean=True (p_sample_loop: only the last sample is used): the fused sampler yields views of its state buffer
        instead of per-step clones and does not materialise pred_xstart."""
        if time is None:
            time = self.num_timesteps
        inner, tmap, rescale, _ = _unwrap(model)
        if tmap is None and hasattr(self, "timestep_map"):        # SpacedDiffusion: the wrapper it would apply
            tmap, rescale = self.timestep_map, self.rescale_timesteps
        if device is None:
            device = next(inner.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        if time > self.num_timesteps:
            raise IndexError(f"Timesteps out of bounds: min=0, max={time - 1}, arr len={self.num_timesteps}")
        indices = list(range(time))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)

        from .wunet import WavUNetModel
        from .unet import UNetModel
        fused = (isinstance(inner, (WavUNetModel, UNetModel)) and hasattr(inner, "engine") and img.is_cuda and cond_fn is None and denoised_fn is None
                 and not model_kwargs and not rescale and shape[1] == 8
                 and self.model_mean_type in (ModelMeanType.START_X, ModelMeanType.EPSILON)
                 and self.model_var_type in (ModelVarType.FIXED_LARGE, ModelVarType.FIXED_SMALL)
                 and (self.mode != 'i2i' or cond is not None))
        if fused:
            from fcwdm.sampler import FusedSampler
            with th.no_grad():
                sampler = FusedSampler.get(self, inner, tuple(img.shape), img.device, clip_denoised,
                                           self.mode == 'i2i')
                sampler.begin(img, cond if self.mode == 'i2i' else None, want_pred=not lean)
                for i in indices:
                    out = sampler.step(i, tmap[i] if tmap is not None else i, clone=not lean)
                    if lean:
                        out["_view"] = True
                    yield out
            return
        for i in indices:
            t = th.tensor([i] * shape[0], device=device)
            with th.no_grad():
                out = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                    cond_fn=cond_fn, model_kwargs=model_kwargs, cond=cond)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ training
    def training_losses(self, model, x_start, t, classifier=None, model_kwargs=None, noise=None, labels=None,
                        mode='default', contr='t1n'):
        """Reference :1084-1166 -> (terms, model_output, model_output_idwt) with terms['mse_wav'] of shape (8,).

        As in the reference the i2i branch is only taken when ``model_kwargs is not None`` (:1099-1102)."""
        cond_dwt = None
        if model_kwargs is None:
            model_kwargs = {}
            target = x_start if th.is_tensor(x_start) else x_start[contr]
        elif mode == 'i2i':
            order = {'t1n': ('t1n', 't1c', 't2w', 't2f'), 't1c': ('t1c', 't1n', 't2w', 't2f'),
                     't2w': ('t2w', 't1n', 't1c', 't2f'), 't2f': ('t2f', 't1n', 't1c', 't2w')}
            if contr not in order:
                raise ValueError("This contrast can't be synthesized.")
            names = order[contr]
            target = x_start[names[0]]
            cond_dwt = th.cat([ops.dwt3d_planar(x_start[k].float(), lll_scale=1.0 / 3.0, concat=True)
                               for k in names[1:]], dim=1)
        else:
            target = x_start
        self._check_t(t)
        x_start_dwt = ops.dwt3d_planar(target.float(), lll_scale=1.0 / 3.0, concat=True)
        noise = th.randn_like(target)                      # image-space noise (:1143), transformed without /3
        noise_dwt = ops.dwt3d_planar(noise.float(), lll_scale=1.0, concat=True)
        x_t = self.q_sample(x_start_dwt, t, noise=noise_dwt)
        if mode == 'i2i' and cond_dwt is not None:
            x_t = th.cat([x_t, cond_dwt], dim=1)
        model_output = model(x_t, self._scale_timesteps(t), **model_kwargs)
        B, _, H, W, D = model_output.size()
        model_output_idwt = idwt(model_output[:, 0, :, :, :].view(B, 1, H, W, D) * 3.,
                                 *[model_output[:, i, :, :, :].view(B, 1, H, W, D) for i in range(1, 8)])
        terms = {"mse_wav": th.mean(mean_flat((x_start_dwt - model_output) ** 2), dim=0)}
        return terms, model_output, model_output_idwt

    # ------------------------------------------------------------------ explicitly out of scope
    def _out_of_scope(self, *a, **k):
        raise NotImplementedError("DDIM / *_known / interpolation / VLB utilities of the reference are outside the "
                                  "fcwdm hot path (SURVEY.md section 2 row 5)")

    ddim_sample = ddim_sample_loop = ddim_sample_loop_progressive = ddim_reverse_sample = _out_of_scope
    p_sample_loop_known = p_sample_loop_interpolation = ddim_sample_loop_known = _out_of_scope
    calc_bpd_loop = _vb_terms_bpd = _prior_bpd = _out_of_scope


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """Gather a 1-D numpy table at `timesteps`, cast to fp32 and broadcast (reference :1246-1263)."""
    if timesteps.numel() and (int(timesteps.min()) < 0 or int(timesteps.max()) >= len(arr)):
        raise IndexError(f"Timesteps out of bounds: min={int(timesteps.min())}, max={int(timesteps.max())}, "
                         f"arr len={len(arr)}")
    res = th.from_numpy(np.asarray(arr)).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)
