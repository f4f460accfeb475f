"""Development probe: throughput of fcwdm.sample_driver.SamplingDriver (disk -> GPU -> disk) on synthetic BraTS-shaped
cases whose voxel values are quantised to 12 bits with a zero background (so gzip behaves roughly as on MR volumes).

    python tools/driver_probe.py [n_cases] [reader_threads] [writer_threads]
"""
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from fcwdm import nifti  # noqa: E402
from fcwdm.sample_driver import SamplingDriver  # noqa: E402
from guided_diffusion.bratsloader import BRATSVolumes  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rt = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    wt = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    root = tempfile.mkdtemp(prefix="fcwdm_driver_")
    try:
        g = np.random.default_rng(0)
        xs = np.linspace(-1, 1, 240, dtype=np.float32)
        ball = (xs[:, None, None] ** 2 + xs[None, :, None] ** 2 + np.linspace(-1, 1, 155, dtype=np.float32)[None, None, :] ** 2) < 0.6
        t0 = time.time()
        for i in range(n):
            subj = f"BraTS-GLI-{i:05d}-000"
            os.makedirs(os.path.join(root, "validation", subj))
            for m in ("t1n", "t1c", "t2w", "t2f"):
                v = np.floor(g.random((240, 240, 155), dtype=np.float32) * 64.0) * 16.0 + 200.0
                nifti.write(os.path.join(root, "validation", subj, f"{subj}-{m}.nii.gz"), (v * ball).astype(np.float32))
        size = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(root) for f in fs)
        print(f"wrote {n} cases, {size / 1e6:.0f} MB, in {time.time() - t0:.1f} s; host cores {os.cpu_count()}", flush=True)
        model, diffusion = bench.build_model(torch.device("cuda"))
        ds = BRATSVolumes(os.path.join(root, "validation"), mode="eval", raw=True)
        for rep in range(2):                                  # the first pass includes kernel / graph warm-up
            out = os.path.join(root, f"out{rep}")
            drv = SamplingDriver(diffusion, model, ds.database, output_dir=out, mode="sample", contr="t1n",
                                 reader_threads=rt, writer_threads=wt, depth=rt)
            st = drv.run()
            drv.close()
            print(f"pass {rep}: {st['cases']} cases in {st['wall_s']:.2f} s = {st['cases'] / st['wall_s']:.2f} cases/s "
                  f"(reader thread-seconds {st['read_s']:.1f}, writer thread-seconds {st['write_s']:.1f}, "
                  f"{st['bytes_written'] / 1e6:.0f} MB written; readers {rt}, writers {wt})", flush=True)
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
