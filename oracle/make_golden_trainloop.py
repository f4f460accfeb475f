"""TEST INFRASTRUCTURE ONLY -- host-side training-driver fixtures from the UNMODIFIED reference
(tests/golden/trainloop_host.json).

    CUDA_VISIBLE_DEVICES="" python -m oracle.make_golden_trainloop

* the timesteps / weights the reference's UniformSampler draws after ``np.random.seed(0)`` (resample.py:42-67), for
  the (T, batch) pairs BASELINE config 4 uses, several consecutive calls each;
* ``parse_resume_step_from_filename`` (train_util.py:516-538) on a list of checkpoint names, including the reference's
  own naming schemes (train_util.py:343,480-487) -- quirks included (a ``BEST_sampled_10`` name parses to 10).
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shims import reference_modules  # noqa: E402
from oracle.make_golden import GOLDEN           # noqa: E402

NAMES = ["brats_t1n_005000_sampled_10.pt", "brats_t1n_BEST_sampled_10.pt", "/data/checkpoints/model012000.pt",
         "opt000250.pt", "ema_0.9999_004000.pt", "brats_t2f_000123_direct_1000.pt", "weights.pt", "a_b_c7.pt",
         "/tmp/run.1/brats_t1c_BEST_direct_100.pt"]


class _D:
    def __init__(self, T):
        self.num_timesteps = T


def main():
    out = {"uniform": [], "parse": {}}
    with reference_modules():
        resample = importlib.import_module("guided_diffusion.resample")
        for T, B in ((10, 2), (10, 1), (1000, 2), (100, 8)):
            np.random.seed(0)
            s = resample.UniformSampler(_D(T), T)
            calls = []
            for _ in range(4):
                t, w = s.sample(B, "cpu")
                calls.append({"t": t.tolist(), "w": w.tolist()})
            out["uniform"].append({"T": T, "B": B, "seed": 0, "calls": calls})
        try:
            tu = importlib.import_module("guided_diffusion.train_util")
            for n in NAMES:
                out["parse"][n] = tu.parse_resume_step_from_filename(n)
        except Exception as exc:                           # wandb / tensorboard import problems: say so, keep going
            out["parse_error"] = repr(exc)
    path = os.path.join(GOLDEN, "trainloop_host.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, os.path.getsize(path), "bytes", out.get("parse_error", ""))


if __name__ == "__main__":
    main()
