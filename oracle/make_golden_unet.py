"""TEST INFRASTRUCTURE ONLY -- fixture of the reference's plain UNetModel (tests/golden/unet_small.npz).

    CUDA_VISIBLE_DEVICES="" python -m oracle.make_golden_unet

Small configuration of the run.sh flag set (dims=3, resblock_updown=True, resample_2d=False, no attention), seeded
weights (zero-initialised convs re-randomised), one forward; plus the state-dict keys / shapes of the full run.sh
configuration (channel_mult 1,2,2,4,4, 64 base channels) for the drop-in's key check.
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
from oracle.make_golden import GOLDEN           # noqa: E402

UNET_SMALL_CFG = dict(image_size=16, in_channels=32, model_channels=32, out_channels=8, num_res_blocks=2,
                      attention_resolutions=(), dropout=0.0, channel_mult=(1, 2, 2), dims=3, num_groups=32,
                      bottleneck_attention=False, resblock_updown=True, resample_2d=False, additive_skips=False,
                      use_scale_shift_norm=False)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with reference_modules():
        unet = importlib.import_module("guided_diffusion.unet")
        model = unet.UNetModel(**UNET_SMALL_CFG)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        sd = owunet.seeded_state_dict(shapes, seed=0)
        model.load_state_dict(sd, strict=True)
        model.to(torch.device("cpu"))
        model.eval()
        g = torch.Generator().manual_seed(13)
        xin = torch.randn(2, 32, 8, 8, 8, generator=g)
        tin = torch.tensor([7, 431])
        with torch.no_grad():
            y = model(xin, tin)
        big = unet.UNetModel(image_size=224, in_channels=32, model_channels=64, out_channels=8, num_res_blocks=2,
                             attention_resolutions=(), channel_mult=(1, 2, 2, 4, 4), dims=3, num_groups=32,
                             bottleneck_attention=False, resblock_updown=True, resample_2d=False)
        bsd = big.state_dict()
        path = os.path.join(GOLDEN, "unet_small.npz")
        np.savez_compressed(path, x=xin.numpy(), t=tin.numpy(), y=y.numpy(), keys=np.array(sorted(shapes)),
                            shapes=np.array([",".join(map(str, shapes[k])) for k in sorted(shapes)]),
                            n_params=sum(p.numel() for p in model.parameters()),
                            big_keys=np.array(list(bsd.keys())),
                            big_shapes=np.array([",".join(map(str, v.shape)) for v in bsd.values()]),
                            big_n_params=sum(p.numel() for p in big.parameters()))
    print(path, os.path.getsize(path), "bytes; out absmax", float(y.abs().max()))


if __name__ == "__main__":
    main()
