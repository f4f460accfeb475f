"""Development probe: launch one conv3d configuration a few times (target for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402

D, H, W, ci, co, k = [int(v) for v in (sys.argv[1:7] if len(sys.argv) >= 7 else (112, 112, 80, 64, 64, 3))]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
with_stats = len(sys.argv) > 8 and sys.argv[8] == "stats"
dev = torch.device("cuda")
S = D * H * W
x = torch.randn((S, max(64, ci)), device=dev).to(torch.bfloat16)
w = torch.randn((co, ci, k, k, k), device=dev) * 0.05
wp = ops.conv3d_pack_weights(w)
b = torch.zeros(co, device=dev)
y = torch.empty((S, co), dtype=torch.bfloat16, device=dev)
stats = torch.zeros((1, 16, 32, 2), dtype=torch.float64, device=dev) if with_stats else None
kw = dict(gn_stats=stats, gn_groups=32) if with_stats else {}
for _ in range(iters):
    ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, k, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, k, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"conv {D}x{H}x{W} {ci}->{co} k{k} stats={with_stats}: {ms*1e3:.1f} us {2.0*S*ci*co*k**3/ms/1e9:.1f} TFLOP/s")
