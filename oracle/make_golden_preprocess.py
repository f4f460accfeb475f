"""TEST INFRASTRUCTURE ONLY -- fixture of the reference's clip_and_normalize (tests/golden/preprocess.npz).

    CUDA_VISIBLE_DEVICES="" python -m oracle.make_golden_preprocess
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shims import reference_modules  # noqa: E402
from oracle.make_golden import GOLDEN           # noqa: E402


def main():
    rng = np.random.default_rng(17)
    # MRI-like: zero background, gamma-distributed foreground, a few hot outliers; values exactly representable in fp32
    vol = rng.gamma(2.0, 180.0, size=(40, 36, 23)).astype(np.float32)
    vol[:6] = 0
    vol[:, :5] = 0
    vol[rng.random(vol.shape) < 0.0005] *= 12.0
    neg = (rng.standard_normal((24, 20, 9)) * 3.0).astype(np.float32)          # signed data, ties at the quantile positions
    neg[::3] = np.round(neg[::3])
    with reference_modules():
        bl = importlib.import_module("guided_diffusion.bratsloader")
        out = {"vol": vol, "vol_out": bl.clip_and_normalize(vol.astype(np.float64)),
               "vol_q": np.array([np.quantile(vol.astype(np.float64), 0.001), np.quantile(vol.astype(np.float64), 0.999)]),
               "neg": neg, "neg_out": bl.clip_and_normalize(neg.astype(np.float64)),
               "neg_q": np.array([np.quantile(neg.astype(np.float64), 0.001), np.quantile(neg.astype(np.float64), 0.999)])}
    path = os.path.join(GOLDEN, "preprocess.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", out["vol_q"], out["neg_q"])


if __name__ == "__main__":
    main()
