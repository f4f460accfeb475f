"""Timestep samplers for training (reference guided_diffusion/resample.py:8-154): ``create_named_schedule_sampler``,
``UniformSampler`` (what run.sh uses; BASELINE config 4 draws ``t`` from it with numpy seed 0) and the loss-aware
second-moment resampler.

``sample`` consumes numpy's global generator exactly as the reference does -- one ``np.random.choice(T, size=(B,), p=p)``
per call (resample.py:54) -- so a seeded run draws the same timesteps.  The two small result tensors are staged in
pinned memory and copied without blocking the host.
"""
import numpy as np
import torch as th
import torch.distributed as dist


def create_named_schedule_sampler(name, diffusion, maxt):
    if name == "uniform":
        return UniformSampler(diffusion, maxt)
    if name == "loss-second-moment":
        return LossSecondMomentResampler(diffusion)
    raise NotImplementedError(f"unknown schedule sampler: {name}")


class ScheduleSampler:
    """A distribution over timesteps; ``sample`` importance-samples from it and returns the weights that keep the
    objective's mean unchanged (1 / (T * p[t]))."""

    def weights(self):
        raise NotImplementedError

    def sample(self, batch_size, device):
        w = np.asarray(self.weights(), dtype=np.float64)
        p = w / w.sum()
        picked = np.random.choice(len(p), size=(batch_size,), p=p)
        self.last_indices = picked                      # host copy of the draw (logging keys that depend on t need no read-back)
        scale = 1.0 / (len(p) * p[picked])
        t_host = th.from_numpy(picked).long()
        w_host = th.from_numpy(scale).float()
        device = th.device(device)
        if device.type == "cuda":
            return (t_host.pin_memory().to(device, non_blocking=True),
                    w_host.pin_memory().to(device, non_blocking=True))
        return t_host.to(device), w_host.to(device)


class UniformSampler(ScheduleSampler):
    def __init__(self, diffusion, maxt=None):
        self.diffusion = diffusion
        self._weights = np.ones([diffusion.num_timesteps if maxt is None else maxt])

    def weights(self):
        return self._weights


class LossAwareSampler(ScheduleSampler):
    def update_with_local_losses(self, local_ts, local_losses):
        """Every rank contributes its (timestep, loss) pairs; all ranks then apply the identical update."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            self.update_with_all_losses([int(v) for v in local_ts.tolist()], [float(v) for v in local_losses.tolist()])
            return
        world = dist.get_world_size()
        n = th.tensor([len(local_ts)], dtype=th.int64, device=local_ts.device)
        counts = [th.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        width = max(counts)
        ts_pad = th.zeros(width, dtype=local_ts.dtype, device=local_ts.device)
        ls_pad = th.zeros(width, dtype=local_losses.dtype, device=local_losses.device)
        ts_pad[:len(local_ts)] = local_ts
        ls_pad[:len(local_losses)] = local_losses
        all_ts = [th.zeros_like(ts_pad) for _ in range(world)]
        all_ls = [th.zeros_like(ls_pad) for _ in range(world)]
        dist.all_gather(all_ts, ts_pad)
        dist.all_gather(all_ls, ls_pad)
        ts, ls = [], []
        for c, a, b in zip(counts, all_ts, all_ls):
            ts.extend(int(v) for v in a[:c].tolist())
            ls.extend(float(v) for v in b[:c].tolist())
        self.update_with_all_losses(ts, ls)

    def update_with_all_losses(self, ts, losses):
        raise NotImplementedError


class LossSecondMomentResampler(LossAwareSampler):
    """p[t] proportional to sqrt(E[loss_t^2]) over the last `history_per_term` losses of each timestep, mixed with a
    small uniform floor; uniform until every timestep has a full history."""

    def __init__(self, diffusion, history_per_term=10, uniform_prob=0.001):
        self.diffusion = diffusion
        self.history_per_term = history_per_term
        self.uniform_prob = uniform_prob
        self._loss_history = np.zeros([diffusion.num_timesteps, history_per_term], dtype=np.float64)
        self._loss_counts = np.zeros([diffusion.num_timesteps], dtype=np.int64)

    def weights(self):
        T = self.diffusion.num_timesteps
        if not self._warmed_up():
            return np.ones([T], dtype=np.float64)
        w = np.sqrt((self._loss_history ** 2).mean(axis=-1))
        w = w / w.sum() * (1.0 - self.uniform_prob) + self.uniform_prob / T
        return w

    def update_with_all_losses(self, ts, losses):
        H = self.history_per_term
        for t, loss in zip(ts, losses):
            n = self._loss_counts[t]
            if n == H:
                self._loss_history[t] = np.append(self._loss_history[t, 1:], loss)
            else:
                self._loss_history[t, n] = loss
                self._loss_counts[t] = n + 1

    def _warmed_up(self):
        return bool((self._loss_counts == self.history_per_term).all())
