"""Development probe: where does a low-resolution conv3d spend its time?  Per-CTA clock64 stamps (fcwdm_debug_set_conv_trace)
of one launch with a cold L2, printed as medians over CTAs in microseconds from kernel entry.
Needs a trace build of the library: FCWDM_CONV_TRACE=1 python fast-cwdm_b200/fcwdm/build.py --force"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import native, ops  # noqa: E402

NAMES = ["entry", "after PDL wait", "first plane requested", "first planes ready", "first weights ready",
         "tile-0 MMAs issued", "tile-0 MMAs retired", "tile-0 stored", "exit"]
dev = torch.device("cuda")
shapes = [(7, 7, 5, 256, 256), (14, 14, 10, 256, 256), (28, 28, 20, 128, 128), (56, 56, 40, 128, 128)]
flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=dev)
clock_ghz = 1.965
for gn_in in (False, True):
    for (D, H, W, ci, co) in shapes:
        S = D * H * W
        x = torch.randn((S, ci), device=dev).to(torch.bfloat16)
        w = torch.randn((co, ci, 3, 3, 3), device=dev) * 0.05
        wp = ops.conv3d_pack_weights(w)
        b = torch.zeros(co, device=dev)
        y = torch.empty((S, co), dtype=torch.bfloat16, device=dev)
        stats = torch.empty((1, ops.GN_STAT_REPLICAS, 32, 2), dtype=torch.float64, device=dev)
        ops.groupnorm_stats(x, stats, 1, S, ci, 32)
        gi = (stats, torch.ones(ci, device=dev), torch.zeros(ci, device=dev), 32, 1e-5) if gn_in else None
        trace = torch.zeros((148, 16), dtype=torch.int64, device=dev)
        for _ in range(2):
            ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, 3, gn_in=gi)
        flush.zero_()
        torch.cuda.synchronize()
        native.load().fcwdm_debug_set_conv_trace(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, 3, gn_in=gi)
        e1.record()
        torch.cuda.synchronize()
        native.load().fcwdm_debug_set_conv_trace(None)
        t = trace.cpu()
        used = t[t[:, 0] != 0]
        rel = (used - used[:, :1]).double() / (clock_ghz * 1e3)
        med = rel.median(dim=0).values
        print(f"\nconv {D}x{H}x{W} {ci}->{co} gn_in={gn_in}: {used.shape[0]} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us (cold L2)")
        for k, name in enumerate(NAMES):
            print(f"   {name:24s} {float(med[k]):8.2f} us")
        acc = used[:, 9:12].double() / (clock_ghz * 1e3)                  # accumulated waits of the MMA-issuing thread
        print(f"   MMA thread waited (median over CTAs): weights {float(acc[:, 0].median()):.2f} us, planes "
              f"{float(acc[:, 1].median()):.2f} us, free accumulator {float(acc[:, 2].median()):.2f} us")
