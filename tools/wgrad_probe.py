"""Development probe: time conv3d wgrad on one shape (target for ncu).  usage: wgrad_probe.py D H W Cin Cout k [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402

D, H, W, ci, co, k = [int(v) for v in (sys.argv[1:7] if len(sys.argv) >= 7 else (112, 112, 80, 64, 64, 3))]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 10
dev = torch.device("cuda")
S = D * H * W
x = torch.randn((S, max(64, ci)), device=dev).to(torch.bfloat16)
dy = torch.randn((S, max(64, co)), device=dev).to(torch.bfloat16)
dw = torch.zeros((co, ci, k, k, k), device=dev)
for _ in range(3):
    ops.conv3d_wgrad(x, dy, dw, (1, D, H, W), ci, co, k, accumulate=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv3d_wgrad(x, dy, dw, (1, D, H, W), ci, co, k, accumulate=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"wgrad {D}x{H}x{W} {ci}->{co} k{k}: {ms*1e3:.1f} us (kernel + finalize) {2.0*S*ci*co*k**3/ms/1e9:.1f} TFLOP/s")
