"""Development probe: the general conv kernel at 56x56x40, 128 -> 128 with fused input GroupNorm, with and without the fused
OUTPUT statistics -- does the epilogue hold the MMAs back?  Needs a trace build (tools/build_variant.sh trace -DFCWDM_CONV_TRACE,
FCWDM_LIB_PATH=tools/_bin/libfcwdm_trace.so)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import native, ops  # noqa: E402

NAMES = ["entry", "after PDL wait", "first plane requested", "first planes ready", "first weights ready",
         "tile-0 MMAs issued", "tile-0 MMAs retired", "tile-0 stored", "exit"]
dev = torch.device("cuda")
flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=dev)
clock_ghz = 1.965
for (D, H, W, ci, co) in [(40, 56, 56, 128, 128), (40, 56, 56, 256, 64)]:
    for want_stats in (False, True):
        S = D * H * W
        x = torch.randn((S, ci), device=dev).to(torch.bfloat16)
        w = torch.randn((co, ci, 3, 3, 3), device=dev) * 0.05
        wp = ops.conv3d_pack_weights(w)
        b = torch.zeros(co, device=dev)
        y = torch.empty((S, co), dtype=torch.bfloat16, device=dev)
        stats = torch.empty((1, ops.GN_STAT_REPLICAS, 32, 2), dtype=torch.float64, device=dev)
        ops.groupnorm_stats(x, stats, 1, S, ci, 32)
        gi = (stats, torch.ones(ci, device=dev), torch.zeros(ci, device=dev), 32, 1e-5) if ci <= 256 and co >= 64 else None
        ost = torch.zeros((1, ops.GN_STAT_REPLICAS, 32, 2), dtype=torch.float64, device=dev) if want_stats else None
        trace = torch.zeros((148, 16), dtype=torch.int64, device=dev)
        kw = dict(gn_in=gi, gn_stats=ost, gn_groups=32 if want_stats else 0)
        for _ in range(2):
            ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, 3, **kw)
        ts = []
        for rep in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            if rep == 4:
                native.load().fcwdm_debug_set_conv_trace(trace.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, 3, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        native.load().fcwdm_debug_set_conv_trace(None)
        t = trace.cpu()
        used = t[t[:, 0] != 0]
        rel = (used - used[:, :1]).double() / (clock_ghz * 1e3)
        med = rel.median(dim=0).values
        print(f"\nconv {D}x{H}x{W} {ci}->{co} output stats={want_stats}: {used.shape[0]} CTAs, event times {[round(v, 1) for v in ts]} us (cold L2)")
        for k, name in enumerate(NAMES):
            print(f"   {name:24s} {float(med[k]):8.2f} us   (max {float(rel[:, k].max()):8.2f})")
        acc = used[:, 9:12].double() / (clock_ghz * 1e3)
        print(f"   MMA thread waited (median / max over CTAs): weights {float(acc[:, 0].median()):.2f} / {float(acc[:, 0].max()):.2f} us, planes "
              f"{float(acc[:, 1].median()):.2f} / {float(acc[:, 1].max()):.2f} us, free accumulator {float(acc[:, 2].median()):.2f} / {float(acc[:, 2].max()):.2f} us")
