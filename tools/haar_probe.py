"""Development probe: a few planar fp32 DWT_3D / IDWT_3D launches on 16 x 224x224x160 (target for ncu --set full)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402

v = torch.rand((1, 16, 224, 224, 160), device="cuda")
for _ in range(3):
    b = ops.dwt3d_planar(v)
    y = ops.idwt3d_planar(b)
torch.cuda.synchronize()
print("round trip max err", float((y - v).abs().max()))
