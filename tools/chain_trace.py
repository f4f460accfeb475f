"""Development probe: where does a layer of the persistent conv chain spend its time?  %globaltimer stamps per (CTA, layer)
(fcwdm_debug_set_chain_trace), printed per layer as microseconds after the moment the previous layer's last CTA arrived at
the grid barrier (median and max over the CTAs that had work).  Needs a trace build:
    FCWDM_CONV_TRACE=1 python fast-cwdm_b200/fcwdm/build.py --force ; FCWDM_LIB_PATH=... python tools/chain_trace.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200"), os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fcwdm import native, ops  # noqa: E402
import chain_probe as cp  # noqa: E402  (prints its own table first)

NAMES = ["A prod at layer", "barrier passed", "sgn ready", "plane0 landed", "plane0 handed", "first MMA", "last MMA issued",
         "acc complete", "exchanged", "stored", "stats flushed", "arrived", "chunk0 tmem", "chunk0 +partials"]
dev = torch.device("cuda")
for name, N, dims, widths in (("7^3 256 x6", 1, (5, 7, 7), (256,) * 7), ("14^3 256 x6", 1, (10, 14, 14), (256,) * 7),
                              ("28^3 128 x6", 1, (20, 28, 28), (128,) * 7)):
    specs = cp.build(N, dims, widths)
    dims4 = (N,) + dims
    counter = torch.zeros(2, dtype=torch.int64, device=dev)
    trace = torch.zeros((160, 32, 16), dtype=torch.int64, device=dev)
    for _ in range(3):
        cp.run_chain(specs, dims4, counter)
    torch.cuda.synchronize()
    native.call("fcwdm_debug_set_chain_trace", ops._ptr(trace))
    cp.run_chain(specs, dims4, counter)
    torch.cuda.synchronize()
    native.call("fcwdm_debug_set_chain_trace", None)
    t = trace.cpu().numpy().astype(np.float64)
    L = len(specs)
    print(f"\n== {name}: us after the previous layer's last arrival (median / max over busy CTAs; n = busy CTAs)")
    print("| layer | n | " + " | ".join(NAMES) + " |")
    print("|---|---|" + "---|" * len(NAMES))
    prev_release = None
    for li in range(L):
        busy = t[:, li, 7] > 0
        alive = t[:, li, 0] > 0
        if li == 0:
            origin = t[alive, li, 0].min()
        else:
            origin = prev_release
        cells = []
        for s in range(len(NAMES)):
            sel = (busy if s >= 2 else alive) & (t[:, li, s] > 0)
            if not sel.any():
                cells.append("-")
                continue
            v = (t[sel, li, s] - origin) / 1e3
            cells.append(f"{np.median(v):.1f} / {v.max():.1f}")
        print(f"| {li} | {int(busy.sum())} | " + " | ".join(cells) + " |")
        arr = t[alive, li, 11]
        prev_release = arr.max() if (arr > 0).any() else t[busy, li, 9].max()
    total = (t[:, L - 1, 9].max() - t[t[:, 0, 0] > 0, 0, 0].min()) / 1e3
    print(f"total {total:.1f} us for {L} layers = {total / L:.1f} us per layer")
