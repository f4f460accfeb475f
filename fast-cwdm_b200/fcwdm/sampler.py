"""Fused reverse-diffusion loop for the WavUNetModel denoiser.

State kept resident in HBM across the T steps of one volume batch (last rows):

* ``x_in``  -- the channels-last bf16 denoiser input (N*S, 64): channels 0..7 = x_t (rewritten in place by the
  step kernel), channels 8..31 = the conditioning sub-bands (written once per volume), 32..63 = zero padding
  that meets zero weights; the reference re-concatenates 128 MB per step instead (gaussian_diffusion.py:297);
* ``x_t``   -- the fp32 planar (N, 8, d, h, w) chain state the posterior mean is computed from (bf16 is only
  what the network sees);
* ``noise`` -- refilled per step with the same torch Philox draw the reference makes (randn_like, :565).

One step = [time embedding, 74 conv3d, 65 GroupNorm+SiLU, 20 DWT/IDWT, 1 fused posterior step] -- about 300
launches with no host synchronisation -- captured once into a CUDA graph and replayed per step with only the
two timestep scalars and the noise buffer changing.
"""
import os

import torch

from . import ops


class FusedSampler:
    _cache_attr = "_samplers"

    noise_hook = None       # optional callable(noise_buffer, diffusion_index) that fills the step's noise in place
    MAX_CACHED = 2          # per diffusion object: each entry pins ~0.3 GB of state + a CUDA graph's private pool (GBs)

    @classmethod
    def get(cls, diffusion, model, shape, device, clip_denoised, i2i):
        """One sampler per (model object, shape, device, clip, i2i).  Weight updates (load_state_dict, optimizer steps)
        do NOT create a new entry: the engine re-packs into its persistent buffers and the captured graph stays valid;
        only when the engine's buffers were re-allocated is the graph dropped and re-captured (see ``_sync_weights``).
        The cache is an LRU of MAX_CACHED entries; an evicted sampler's graph and buffers are released."""
        key = (id(model), tuple(shape), str(device), bool(clip_denoised), bool(i2i))
        cache = getattr(diffusion, cls._cache_attr)
        s = cache.pop(key, None)
        if s is not None and s.model is not model:          # id() reuse after the old model was collected
            s.release()
            s = None
        if s is None:
            s = cls(diffusion, model, shape, device, clip_denoised, i2i)
        cache[key] = s                                      # most recently used last
        while len(cache) > cls.MAX_CACHED:
            cache.pop(next(iter(cache))).release()
        return s

    def release(self):
        self.graph = None
        self.x_in = self.x_t = self.pred = self.noise = None

    def _sync_weights(self):
        """Bring the engine's packed operand copies up to date OUTSIDE the graph (the captured launch sequence only reads
        them) and drop the graph if the buffers it references were re-allocated."""
        eng = self.engine
        eng.prepare(self.device)
        if eng.buffers_id != self._buffers_id:
            self.graph = None
            self._eager_steps = 0
            self._buffers_id = eng.buffers_id

    def __init__(self, diffusion, model, shape, device, clip_denoised, i2i):
        from guided_diffusion.gaussian_diffusion import ModelMeanType
        N, C, d, h, w = shape
        assert C == 8
        self.diffusion, self.model, self.shape, self.device = diffusion, model, tuple(shape), device
        self.clip = bool(clip_denoised)
        self.i2i = i2i
        self.predict_xstart = diffusion.model_mean_type == ModelMeanType.START_X
        self.N, self.S, self.dims = N, d * h * w, (d, h, w)
        self.engine = model.engine()
        ld_in = (model.in_channels + 63) // 64 * 64
        self.x_in = torch.zeros((N * self.S, ld_in), dtype=torch.bfloat16, device=device)
        # the step kernel is per-voxel local, so x_{t-1} overwrites x_t in place: no second state buffer, no copy
        self.x_t = torch.empty(shape, dtype=torch.float32, device=device)
        self.pred = None                                  # allocated on the first step that asks for pred_xstart
        self.want_pred = True
        self.noise = torch.empty(shape, dtype=torch.float32, device=device)
        self.t_diff = torch.zeros((N,), dtype=torch.int64, device=device)
        self.t_model = torch.zeros((N,), dtype=torch.int64, device=device)
        self.coef = diffusion._table("step", device)
        self.graph = None
        self._eager_steps = 0
        self._buffers_id = -1
        self.use_graph = os.environ.get("FCWDM_NO_GRAPH", "0") != "1"
        self.launches_per_step = None

    # ------------------------------------------------------------------
    def begin(self, noise, cond, want_pred=True):
        """Load a new volume batch: x_T and (i2i) the 24 conditioning channels.  want_pred=False (p_sample_loop, which
        only returns the final sample): the step kernel does not materialise pred_xstart (32 MB less per step)."""
        self._sync_weights()
        if bool(want_pred) != self.want_pred:
            self.want_pred = bool(want_pred)
            self.graph = None                             # the captured step kernel carries the pred pointer
            self._eager_steps = min(self._eager_steps, 1)
        if self.want_pred and self.pred is None:
            self.pred = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        self.x_t.copy_(noise)
        ops.planar_to_cl(self.x_t, self.x_in, 8)
        if self.i2i:
            nc = cond.shape[1]
            if nc + 8 != self.model.in_channels:
                raise ValueError(f"cond has {nc} channels; model expects {self.model.in_channels - 8}")
            ops.planar_to_cl(cond.float(), self.x_in[:, 8:], nc)

    def _step_body(self):
        from . import native
        n0 = native.launch_count
        out_cl = self.engine.forward_cl(self.x_in, self.t_model, self.N, self.dims)
        d, h, w = self.dims
        with ops._on(self.device) as st:
            native.call("fcwdm_p_sample_step", ops._ptr(out_cl), out_cl.stride(0), ops._ptr(self.x_t),
                        ops._ptr(self.noise), ops._ptr(self.x_t), ops._ptr(self.pred if self.want_pred else None),
                        ops._ptr(self.x_in),
                        self.x_in.stride(0), ops._ptr(self.coef), ops._ptr(self.t_diff), self.coef.shape[0], self.N, d, h,
                        w, 1 if self.clip else 0, 1 if self.predict_xstart else 0, st)
        self.launches_per_step = native.launch_count - n0

    def step(self, i, t_model, clone=True):
        """One reverse step at diffusion index i (the model sees t_model = timestep_map[i]).

        The first step of a sampler's life runs eagerly (lazy init, weight packing); the launch sequence is then
        captured once (capture records, it does not execute) and every later step is a graph replay."""
        self.t_diff.fill_(int(i))
        self.t_model.fill_(int(t_model))
        if self.noise_hook is not None:
            self.noise_hook(self.noise, int(i))   # tests / reproducibility: the caller supplies this step's noise
        else:
            self.noise.normal_()                  # == th.randn_like(x): same generator, same draw order
        if not self.use_graph or self._eager_steps == 0:
            self._step_body()
            self._eager_steps += 1
        else:
            if self.graph is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_body()
                self.graph = g
            self.graph.replay()
        if clone:                                 # the generator API hands out tensors the caller may keep
            return {"sample": self.x_t.clone(), "pred_xstart": self.pred.clone() if self.want_pred else None}
        return {"sample": self.x_t, "pred_xstart": self.pred if self.want_pred else None}
