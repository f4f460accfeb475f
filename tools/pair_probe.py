"""Development probe: time the CTA-pair conv (and the single-CTA kernel) on one shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402

D, H, W, ci, co = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (112, 112, 80, 64, 64))]
iters = 10
dev = torch.device("cuda")
S = D * H * W
x = torch.randn((S, 64), device=dev).to(torch.bfloat16)
w = torch.randn((co, ci, 3, 3, 3), device=dev) * 0.05
b = torch.zeros(co, device=dev)
y = torch.empty((S, max(8, co)), dtype=torch.bfloat16, device=dev)
for name, wp, fn in (("pair", ops.conv3d_pair_pack_weights(w), ops.conv3d_pair_cl),
                     ("single", ops.conv3d_pack_weights(w), None)):
    def run():
        if fn is not None:
            fn(x, wp, b, y, (1, D, H, W), ci, co)
        else:
            ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, 3)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:6s} conv {D}x{H}x{W} {ci}->{co}: {ms*1e3:.1f} us {2.0*S*ci*co*27/ms/1e9:.1f} TFLOP/s", flush=True)

# fused input GroupNorm variant
G = 32
stats = torch.empty((1, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device=dev)
ops.groupnorm_stats(x, stats, 1, S, 64, G)
gamma = torch.ones(64, device=dev)
beta = torch.zeros(64, device=dev)
wp = ops.conv3d_pair_pack_weights(w)
ostats = torch.zeros((1, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device=dev)
for name, kw in (("pair+gn_in", dict(gn_in=(stats, gamma, beta, G, 1e-5))),
                 ("pair+gn_in+stats", dict(gn_in=(stats, gamma, beta, G, 1e-5), gn_stats=ostats, gn_groups=G) if co == 64 else None),
                 ("pair+stats", dict(gn_stats=ostats, gn_groups=G) if co == 64 else None)):
    if kw is None:
        continue
    for _ in range(3):
        ops.conv3d_pair_cl(x, wp, b, y, (1, D, H, W), ci, co, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv3d_pair_cl(x, wp, b, y, (1, D, H, W), ci, co, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:18s} conv {D}x{H}x{W} {ci}->{co}: {ms*1e3:.1f} us {2.0*S*ci*co*27/ms/1e9:.1f} TFLOP/s", flush=True)

# residual / channel-bias variants (the ResBlock's second conv: GN_IN + residual + output statistics)
res = torch.randn((S, 64), device=dev).to(torch.bfloat16)
cbias = torch.randn((1, 64), device=dev)
for name, kw in (("pair+res", dict(residual=res)),
                 ("pair+gn_in+res+stats", dict(gn_in=(stats, gamma, beta, G, 1e-5), residual=res, gn_stats=ostats, gn_groups=G)),
                 ("pair+gn_in+cb+stats", dict(gn_in=(stats, gamma, beta, G, 1e-5), chan_bias=cbias, gn_stats=ostats, gn_groups=G))):
    if co != 64:
        continue
    for _ in range(3):
        ops.conv3d_pair_cl(x, wp, b, y, (1, D, H, W), ci, co, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv3d_pair_cl(x, wp, b, y, (1, D, H, W), ci, co, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:22s} conv {D}x{H}x{W} {ci}->{co}: {ms*1e3:.1f} us {2.0*S*ci*co*27/ms/1e9:.1f} TFLOP/s", flush=True)
