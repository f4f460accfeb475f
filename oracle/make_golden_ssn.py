"""TEST INFRASTRUCTURE ONLY -- fixtures for use_scale_shift_norm=True from the UNMODIFIED reference
(tests/golden/wunet_small_ssn.npz, tests/golden/unet_small_ssn.npz).

    CUDA_VISIBLE_DEVICES="" python -m oracle.make_golden_ssn

`use_scale_shift_norm=True` is the default of the reference's model_and_diffusion_defaults() (script_util.py) although
run.sh passes False: ResBlock then computes out_norm(h) * (1 + scale) + shift with (scale, shift) = chunk(emb_out, 2)
instead of h + emb_out (wunet.py:256-260, unet.py:301-305).  One forward of each small model with seeded weights.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shims import reference_modules  # noqa: E402
from oracle import wunet as owunet              # noqa: E402
from oracle.make_golden import SMALL_CFG, GOLDEN  # noqa: E402
from oracle.make_golden_unet import UNET_SMALL_CFG  # noqa: E402


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with reference_modules() as ref:
        unet = importlib.import_module("guided_diffusion.unet")
        for name, build, cfg, tie in (("wunet_small_ssn", ref.wunet.WavUNetModel, SMALL_CFG, True),
                                      ("unet_small_ssn", unet.UNetModel, UNET_SMALL_CFG, False)):
            cfg = dict(cfg, use_scale_shift_norm=True)
            model = build(**cfg)
            shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
            sd = owunet.seeded_state_dict(shapes, seed=0)
            if tie:
                sd = owunet.tie_output_blocks(sd, len(cfg["channel_mult"]))
            model.load_state_dict(sd, strict=True)
            if hasattr(model, "devices"):
                model.to(torch.device("cpu"))
            model.eval()
            g = torch.Generator().manual_seed(23)
            xin = torch.randn(2, 32, 8, 8, 8, generator=g)
            tin = torch.tensor([5, 641])
            with torch.no_grad():
                y = model(xin, tin)
            path = os.path.join(GOLDEN, name + ".npz")
            np.savez_compressed(path, x=xin.numpy(), t=tin.numpy(), y=y.numpy(), keys=np.array(sorted(shapes)),
                                shapes=np.array([",".join(map(str, shapes[k])) for k in sorted(shapes)]))
            print(path, os.path.getsize(path), "bytes; |y| max", float(y.abs().max()))


if __name__ == "__main__":
    main()
