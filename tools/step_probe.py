"""Development probe: N eager (un-graphed) denoising steps of CFG-W4 at full size -- the target command for the
ncu launch list (profiles/).  usage: step_probe.py [steps] [batch]"""
import os
import sys

os.environ["FCWDM_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1                 # volumes per call (BASELINE config 3: 8)
dev = torch.device("cuda")
model, diffusion = bench.build_model(dev)
noise = torch.randn(B, 8, 112, 112, 80, device=dev)
cond = torch.rand(B, 24, 112, 112, 80, device=dev)
it = diffusion.p_sample_loop_progressive(model, noise.shape, time=steps, noise=noise, cond=cond, progress=False)
for k, out in enumerate(it):
    torch.cuda.synchronize()
print("finite", bool(torch.isfinite(out["sample"]).all()))
