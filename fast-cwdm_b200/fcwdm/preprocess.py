"""GPU version of the reference's per-case preprocessing (guided_diffusion/bratsloader.py:40-109): the loader's
``clip_and_normalize`` (0.1 % / 99.9 % quantile clip + min-max to [0, 1]) followed by the zero-padding of the slice
axis to 160 and the 8-voxel crop of the in-plane axes (240 -> 224).  Raw volumes are uploaded once; everything else is
``fcwdm_clip_normalize`` (radix-select order statistics + one elementwise pass), so a case goes
disk -> GPU -> conditioning sub-bands without a host-side sort."""
import ctypes

import torch

from . import native, ops

_VP = ctypes.c_void_p


def clip_and_normalize(raw, crop=8, pad_to=160, q_lo=0.001, q_hi=0.999, return_quantiles=False):
    """raw: (V, X, Y, Z) float32 CUDA raw intensities (e.g. (4, 240, 240, 155): the four modalities of one case, each
    normalised on its own).  Returns (V, 1, X - 2*crop, Y - 2*crop, pad_to) float32 in [0, 1]."""
    ops._need_cuda(raw, "clip_and_normalize")
    if raw.dim() != 4:
        raise ValueError(f"expected (V, X, Y, Z), got {tuple(raw.shape)}")
    raw = raw.float().contiguous()
    V, X, Y, Z = raw.shape
    out = torch.empty((V, 1, X - 2 * crop, Y - 2 * crop, max(pad_to, Z)), dtype=torch.float32, device=raw.device)
    q = torch.empty((V, 2), dtype=torch.float32, device=raw.device)
    nbytes = native.load().fcwdm_clip_normalize_workspace_bytes(V)
    ws = torch.empty(max(nbytes, 64), dtype=torch.uint8, device=raw.device)
    with ops._on(raw.device) as st:
        native.call("fcwdm_clip_normalize", ops._ptr(raw), ops._ptr(out), ops._ptr(q), ops._ptr(ws), ws.numel(), V, X, Y, Z,
                    crop, crop, max(pad_to, Z), float(q_lo), float(q_hi), st)
    return (out, q) if return_quantiles else out


def preprocess_case(t1n=None, t1c=None, t2w=None, t2f=None):
    """The dict BRATSVolumes.__getitem__ builds (bratsloader.py:44-100) from raw (240, 240, 155) CUDA volumes: every
    present modality -> (1, 224, 224, 160) float32; a missing one -> torch.zeros(1) and its name under 'missing'."""
    given = {k: v for k, v in (("t1n", t1n), ("t1c", t1c), ("t2w", t2w), ("t2f", t2f)) if v is not None}
    if not given:
        raise ValueError("at least one modality is required")
    names = list(given)
    stacked = clip_and_normalize(torch.stack([given[k].float() for k in names], dim=0))
    out = {k: stacked[i] for i, k in enumerate(names)}
    missing = "none"
    for k in ("t1n", "t1c", "t2w", "t2f"):
        if k not in out:
            missing = k
            out[k] = torch.zeros(1)
    out["missing"] = missing
    return out
