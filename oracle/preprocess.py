"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's per-case preprocessing.

clip_and_normalize (guided_diffusion/bratsloader.py:107-111): clip to the 0.1 % / 99.9 % quantiles (np.quantile,
linear interpolation, float64), then min-max to [0, 1].  BRATSVolumes.__getitem__ (:44-50): place the (240, 240, 155)
result in a zero (1, 240, 240, 160) tensor and crop 8 voxels off both ends of the two in-plane axes, cast to float32.
"""
import numpy as np


def clip_and_normalize(img):
    img = np.asarray(img, dtype=np.float64)
    lo, hi = np.quantile(img, 0.001), np.quantile(img, 0.999)
    clipped = np.clip(img, lo, hi)
    return (clipped - np.min(clipped)) / (np.max(clipped) - np.min(clipped))


def pad_crop(vol, crop=8, pad_to=160):
    """(X, Y, Z) -> (1, X - 2 crop, Y - 2 crop, pad_to) float32 (bratsloader.py:47-50)."""
    X, Y, Z = vol.shape
    out = np.zeros((1, X, Y, max(pad_to, Z)), dtype=np.float32)
    out[:, :, :, :Z] = vol.astype(np.float32)
    return out[:, crop:X - crop, crop:Y - crop, :]


def preprocess_volume(raw, crop=8, pad_to=160):
    return pad_crop(clip_and_normalize(raw), crop, pad_to)
