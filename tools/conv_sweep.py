"""Development probe: time every conv3d shape of CFG-W4 through fcwdm_conv3d_fwd (general single-CTA kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402
from perf_probe import timeit  # noqa: E402

dev = torch.device("cuda")
convs = [(112, 112, 80, 64, 64, 3), (56, 56, 40, 128, 128, 3), (56, 56, 40, 256, 64, 3), (56, 56, 40, 64, 128, 3),
         (28, 28, 20, 128, 128, 3), (28, 28, 20, 512, 128, 3), (14, 14, 10, 128, 256, 3), (14, 14, 10, 256, 256, 3),
         (14, 14, 10, 1024, 128, 3), (7, 7, 5, 256, 256, 3), (7, 7, 5, 1024, 256, 3), (56, 56, 40, 64, 128, 1),
         (14, 14, 10, 128, 256, 1)]
tot = 0.0
for (D, H, W, ci, co, k) in convs:
    S = D * H * W
    x = torch.randn((S, max(64, ci)), device=dev).to(torch.bfloat16)
    w = torch.randn((co, ci, k, k, k), device=dev) * 0.05
    wp = ops.conv3d_pack_weights(w)
    b = torch.zeros(co, device=dev)
    y = torch.empty((S, co), dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, k), iters=20)
    tot += ms
    print(f"conv {D}x{H}x{W} {ci}->{co} k{k}: {ms*1e3:8.1f} us  {2.0*S*ci*co*k**3/ms/1e9:8.1f} TFLOP/s", flush=True)
print(f"sum {tot*1e3:.1f} us  (FCWDM_CONV_BSTAGES={os.environ.get('FCWDM_CONV_BSTAGES')})")
