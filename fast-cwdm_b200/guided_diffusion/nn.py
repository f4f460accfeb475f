"""Helpers mirrored from the reference's guided_diffusion/nn.py (names, arguments and results identical; the
heavy ones are only parameter containers here -- the denoiser's forward runs on fcwdm kernels)."""
import torch as th
import torch.nn as nn


class SiLU(nn.Module):
    def forward(self, x):
        return x * th.sigmoid(x)


class GroupNorm32(nn.GroupNorm):
    """fp32-compute GroupNorm (reference nn.py:17-19)."""

    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def conv_nd(dims, *args, **kwargs):
    """reference nn.py:22-32."""
    try:
        return {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[dims](*args, **kwargs)
    except KeyError:
        raise ValueError(f"unsupported dimensions: {dims}")


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def avg_pool_nd(dims, *args, **kwargs):
    try:
        return {1: nn.AvgPool1d, 2: nn.AvgPool2d, 3: nn.AvgPool3d}[dims](*args, **kwargs)
    except KeyError:
        raise ValueError(f"unsupported dimensions: {dims}")


def update_ema(target_params, source_params, rate=0.99):
    for targ, src in zip(target_params, source_params):
        targ.detach().mul_(rate).add_(src, alpha=1 - rate)


def zero_module(module):
    """reference nn.py:68-74."""
    for p in module.parameters():
        p.detach().zero_()
    return module


def scale_module(module, scale):
    for p in module.parameters():
        p.detach().mul_(scale)
    return module


def mean_flat(tensor):
    """Mean over dims 2.. (NOT 1..: reference nn.py:86-90 keeps the channel axis)."""
    return tensor.mean(dim=list(range(2, len(tensor.shape))))


def normalization(channels, groups=32):
    return GroupNorm32(groups, channels)


def timestep_embedding(timesteps, dim, max_period=10000):
    """Sinusoidal embedding, reference nn.py:103-121, computed by fcwdm_timestep_embedding (CUDA only)."""
    from fcwdm import ops
    from fcwdm.native import FcwdmError
    if not timesteps.is_cuda:
        raise FcwdmError("timestep_embedding: CUDA tensor required (no CPU fallback)")
    if timesteps.is_floating_point():
        raise NotImplementedError("fractional timesteps (rescale_timesteps=True) are not implemented; the shipped "
                                  "configuration uses rescale_timesteps=False (run.sh:129)")
    out = th.empty((timesteps.shape[0], dim), dtype=th.float32, device=timesteps.device)
    ops.timestep_embedding(timesteps.to(th.int64).contiguous(), out, dim, float(max_period))
    return out
