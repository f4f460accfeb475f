"""Development probe: a run of L ResBlock convs at one low resolution, as ONE chain launch vs one launch per layer.
usage: chain_probe.py            (env FCWDM_CHAIN_CLUSTER=2|4, FCWDM_CHAIN_SPLIT=1|2|4 select variants)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402

G, EPS = 32, 1e-5
dev = torch.device("cuda")


def build(N, dims, widths):
    D, H, W = dims
    S = D * H * W
    gen = torch.Generator().manual_seed(0)
    cur = (torch.randn(N * S, widths[0], generator=gen)).to(dev).to(torch.bfloat16)
    prev = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device=dev)
    ops.groupnorm_stats(cur, prev, N, S, widths[0], G) if widths[0] <= 256 else None
    specs = []
    cin = widths[0]
    for li, cout in enumerate(widths[1:]):
        w = (torch.randn(cout, cin, 3, 3, 3, generator=gen) / np.sqrt(cin * 27)).to(dev)
        wp = ops.conv3d_pack_weights(w)
        bias = torch.randn(cout, generator=gen).to(dev)
        gamma, beta = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
        cb = torch.randn(N, cout, generator=gen).to(dev)
        y = torch.zeros((N * S, cout), dtype=torch.bfloat16, device=dev)
        st = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device=dev)
        specs.append(dict(x=cur, wp=wp, bias=bias, y=y, cin=cin, cout=cout, st=st, gn=(prev, gamma, beta, G, EPS) if cin <= 256 else None,
                          cb=cb if li % 2 == 0 else None, res=(specs[-1]["x"] if (li % 2 == 1 and specs[-1]["cin"] == cout) else None)))
        cur, prev, cin = y, st, cout
    return specs


def run_chain(specs, dims4, counter):
    counter.zero_()
    for s in specs:
        s["st"].zero_()
    layers = [ops.conv3d_chain_layer(s["x"], s["wp"], s["bias"], s["y"], dims4, s["cin"], s["cout"], chan_bias=s["cb"],
                                     residual=s["res"], gn_stats=s["st"], gn_groups=G, gn_in=s["gn"]) for s in specs]
    ops.conv3d_chain([l for l, _ in layers], counter)


def run_layers(specs, dims4):
    for s in specs:
        s["st"].zero_()
    for s in specs:
        ops.conv3d_cl(s["x"], s["wp"], s["bias"], s["y"], dims4, s["cin"], s["cout"], 3, chan_bias=s["cb"], residual=s["res"],
                      gn_stats=s["st"], gn_groups=G, gn_in=s["gn"])


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    print(f"cluster={os.environ.get('FCWDM_CHAIN_CLUSTER', 'auto')} split cap={os.environ.get('FCWDM_CHAIN_SPLIT', '-')}")
    for name, N, dims, widths in (
            ("28^3 128 x8", 1, (20, 28, 28), (128,) * 9),
            ("28^3 512->128 +6", 1, (20, 28, 28), (512,) + (128,) * 7),
            ("14^3 256 x6", 1, (10, 14, 14), (256,) * 7),
            ("14^3 1024->128,128->256,256 x4", 1, (10, 14, 14), (1024, 128, 256, 256, 256, 256)),
            ("7^3 256 x12", 1, (5, 7, 7), (256,) * 13),
            ("7^3 1024->256 +11", 1, (5, 7, 7), (1024,) + (256,) * 12),
            ("14^3 256 x6 batch 8", 8, (10, 14, 14), (256,) * 7)):
        specs = build(N, dims, widths)
        dims4 = (N,) + dims
        counter = torch.zeros(2, dtype=torch.int64, device=dev)
        flop = sum(2.0 * N * dims[0] * dims[1] * dims[2] * s["cin"] * s["cout"] * 27 for s in specs)
        t_chain = timeit(lambda: run_chain(specs, dims4, counter))
        ya = [s["y"].float().clone() for s in specs]
        t_layers = timeit(lambda: run_layers(specs, dims4))
        diff = max(float((a - s["y"].float()).abs().max() / s["y"].float().abs().max().clamp_min(1e-6)) for a, s in zip(ya, specs))
        L = len(specs)
        print(f"{name:32s} L={L:2d}: chain {t_chain:7.1f} us ({t_chain / L:5.1f}/layer, {flop / t_chain / 1e6:6.0f} TFLOP/s) | per-layer launches "
              f"{t_layers:7.1f} us ({t_layers / L:5.1f}/layer, {flop / t_layers / 1e6:6.0f} TFLOP/s) | max rel diff {diff:.1e}", flush=True)


if __name__ == "__main__":
    main()
