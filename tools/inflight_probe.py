"""Development probe: throughput with K volumes in flight on K CUDA streams (batch 1 each, one graph per stream)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from fcwdm import pipeline  # noqa: E402
from guided_diffusion.script_util import create_gaussian_diffusion  # noqa: E402

dev = torch.device("cuda")
model, diffusion0 = bench.build_model(dev)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = 6
vol, noise = bench.synth_volume(1)
vol, noise = vol.to(dev), noise.to(dev)
diffs = [create_gaussian_diffusion(steps=10, predict_xstart=True, sample_schedule="sampled", mode="i2i") for _ in range(K)]
streams = [torch.cuda.Stream(dev) for _ in range(K)]
for d, s in zip(diffs, streams):                       # warm-up + capture, one stream at a time
    with torch.cuda.stream(s):
        for _ in range(2):
            pipeline.synthesize(d, model, vol[:, 1:2], vol[:, 2:3], vol[:, 3:4], noise)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    for d, s in zip(diffs, streams):
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            out = pipeline.synthesize(d, model, vol[:, 1:2], vol[:, 2:3], vol[:, 3:4], noise)
for s in streams:
    torch.cuda.current_stream(dev).wait_stream(s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{K} in flight: {reps * K / (ms * 1e-3):.2f} volumes/s ({ms / (reps * K):.1f} ms per volume), finite={bool(torch.isfinite(out).all())}")
