"""BASELINE config 5: DWT_3D / IDWT_3D standalone bandwidth sweep, cubes 64^3 .. 256^3 plus 224x224x160, fp32 and
bf16, channel count chosen so each launch moves >= 512 MB (>> 126 MB L2); CUDA events, best of 10 after 3 warm-ups.
Prints a markdown table (committed as profiles/r01_haar_sweep.md)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def best_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print(f"| shape (D,H,W) | dtype | C | MB/launch | DWT us | DWT GB/s | frac | IDWT us | IDWT GB/s | frac |")
print("|---|---|---|---|---|---|---|---|---|---|")
shapes = [(n, n, n) for n in (64, 96, 128, 160, 192, 224, 256)] + [(224, 224, 160)]
for dt in (torch.float32, torch.bfloat16):
    for (D, H, W) in shapes:
        vox = D * H * W
        esz = 4 if dt == torch.float32 else 2
        C = max(1, -(-(256 * 2 ** 20) // (vox * esz)))          # >= 256 MB in + 256 MB out
        x = torch.rand((1, C, D, H, W), device="cuda").to(dt)
        bands = ops.dwt3d_planar(x)
        nbytes = 2.0 * x.numel() * esz
        t_d = best_ms(lambda: ops.dwt3d_planar(x))
        t_i = best_ms(lambda: ops.idwt3d_planar(bands))
        gd, gi = nbytes / t_d / 1e6, nbytes / t_i / 1e6
        print(f"| {D}x{H}x{W} | {'fp32' if esz == 4 else 'bf16'} | {C} | {nbytes/2**20:.0f} | {t_d*1e3:.1f} | {gd:.0f} | "
              f"{gd/peak:.3f} | {t_i*1e3:.1f} | {gi:.0f} | {gi/peak:.3f} |", flush=True)
        del x, bands
print(f"\nHBM peak (measured copy, MEASURED_PEAKS.json): {peak} GB/s; bytes = 2 * numel * sizeof (read once, write once).")
