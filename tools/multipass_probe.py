"""Development probe: a C_in / C_out > 64 full-resolution conv as several 64x64 CTA-pair launches (channel-block passes
accumulating through the residual input, output halves written to column slices) against the general kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402
from perf_probe import timeit  # noqa: E402

dev = torch.device("cuda")
shapes = [(112, 112, 80, 128, 64), (112, 112, 80, 192, 64), (112, 112, 80, 128, 128), (56, 56, 40, 128, 128),
          (56, 56, 40, 256, 128)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1:6])]
for (D, H, W, ci, co) in shapes:
    S = D * H * W
    x = torch.randn((S, ci), device=dev).to(torch.bfloat16)
    w = torch.randn((co, ci, 3, 3, 3), device=dev) * 0.02
    b = torch.randn(co, device=dev)
    wp = ops.conv3d_pack_weights(w)
    y0 = torch.empty((S, co), dtype=torch.bfloat16, device=dev)
    y1 = torch.empty((S, co), dtype=torch.bfloat16, device=dev)
    blocks = [[ops.conv3d_pair_pack_weights(w[o:o + 64, i:i + 64].contiguous()) for i in range(0, ci, 64)]
              for o in range(0, co, 64)]

    def general():
        ops.conv3d_cl(x, wp, b, y0, (1, D, H, W), ci, co, 3)

    def multipass():
        for oi, o in enumerate(range(0, co, 64)):
            yv = y1[:, o:o + 64]
            for ii, i in enumerate(range(0, ci, 64)):
                ops.conv3d_pair_cl(x[:, i:i + 64], blocks[oi][ii], b[o:o + 64] if ii == 0 else None, yv, (1, D, H, W), 64, 64,
                                   residual=yv if ii else None)

    general()
    multipass()
    torch.cuda.synchronize()
    err = float((y1.float() - y0.float()).norm() / y0.float().norm())
    mg = timeit(general, iters=10)
    mm = timeit(multipass, iters=10)
    fl = 2.0 * S * ci * co * 27
    print(f"conv {D}x{H}x{W} {ci}->{co}: general {mg*1e3:7.1f} us ({fl/mg/1e9:6.0f} TF/s)   pair x{(ci//64)*(co//64)} passes "
          f"{mm*1e3:7.1f} us ({fl/mm/1e9:6.0f} TF/s)   rel diff {err:.2e}", flush=True)
