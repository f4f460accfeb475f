"""Development probe for csrc/conv3d_chain.cu: small chains with one feature switched on at a time; prints per-layer
max error against torch fp32."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from fcwdm import ops  # noqa: E402
from gpu_util import bf16_round, from_cl, to_cl  # noqa: E402

G, EPS = 32, 1e-5
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run(name, N, dims, widths, gn=False, cb=False, res=False, stats=True, seed=0):
    D, H, W = dims
    S = D * H * W
    gen = torch.Generator().manual_seed(seed)
    x0 = bf16_round(torch.randn(N, widths[0], D, H, W, generator=gen)).cuda()
    cur, cur_c, prev = to_cl(x0), widths[0], None
    layers, meta = [], []
    for li, cout in enumerate(widths[1:]):
        cin = cur_c
        w = bf16_round(torch.randn(cout, cin, 3, 3, 3, generator=gen) / np.sqrt(cin * 27)).cuda()
        bias = torch.randn(cout, generator=gen).cuda()
        wp = ops.conv3d_pack_weights(w)
        y = torch.zeros((N * S, cout), dtype=torch.bfloat16, device="cuda")
        st = torch.zeros((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device="cuda") if stats else None
        g = gi = c = r = None
        if gn and li > 0:
            gamma, beta = (torch.rand(cin, generator=gen) + 0.5).cuda(), (torch.randn(cin, generator=gen) * 0.2).cuda()
            g, gi = (gamma, beta), (prev, gamma, beta, G, EPS)
        if cb and li > 0:
            c = torch.randn(N, cout, generator=gen).cuda()
        if res and li > 0:
            r = bf16_round(torch.randn(N, cout, D, H, W, generator=gen)).cuda()
        rc = to_cl(r) if r is not None else None
        layers.append(ops.conv3d_chain_layer(cur, wp, bias, y, (N, D, H, W), cin, cout, chan_bias=c, residual=rc, gn_stats=st,
                                             gn_groups=G if stats else 0, gn_in=gi))
        meta.append((cur, cin, w, bias, c, r, g, y, cout, st))
        cur, cur_c, prev = y, cout, st
    counter = torch.zeros(2, dtype=torch.int64, device="cuda")
    ops.conv3d_chain([l for l, _ in layers], counter)
    torch.cuda.synchronize()
    line = [f"{name:34s}"]
    for li, (x_cl, cin, w, bias, c, r, g, y, cout, st) in enumerate(meta):
        x = from_cl(x_cl, (N, cin, D, H, W))
        a = x
        if g is not None:
            a = bf16_round(F.silu(F.group_norm(x, G, g[0], g[1], EPS)))
        ref = F.conv3d(a, w, bias, padding=1)
        if c is not None:
            ref = ref + c[:, :, None, None, None]
        if r is not None:
            ref = ref + r
        got = from_cl(y, (N, cout, D, H, W))
        err = float((got - ref).abs().max()) / float(ref.abs().max())
        zeros = float((got == 0).float().mean())
        serr = ""
        if st is not None:
            yy = got.double().reshape(N, G, -1)
            s = st.sum(dim=1)
            serr = f" st {float(((s[..., 0] - yy.sum(-1)).abs() / (yy.abs().sum(-1) + 1)).max()):.1e}"
        line.append(f"L{li}: {err:.2e} z{zeros:.2f}{serr}")
    print(" | ".join(line), f"| counter {int(counter[0])}", flush=True)


bott = (1, (5, 7, 7))
run("1 layer", *bott, (256, 256))
run("2 plain", *bott, (256, 256, 256))
run("3 plain", *bott, (256, 256, 256, 256))
run("2 plain nostats", *bott, (256, 256, 256), stats=False)
run("2 cb", *bott, (256, 256, 256), cb=True)
run("2 res", *bott, (256, 256, 256), res=True)
run("2 gn", *bott, (256, 256, 256), gn=True)
run("2 gn cb res", *bott, (256, 256, 256), gn=True, cb=True, res=True)
run("3 gn", *bott, (256, 256, 256, 256), gn=True)
run("2 plain 128 (split 2)", *bott, (128, 128, 128))
run("2 gn 128 (split 2)", *bott, (128, 128, 128), gn=True)
run("2 plain 64 (split 1)", *bott, (64, 128, 128))
run("2 gn 64->128->128", *bott, (64, 128, 128), gn=True)
run("14^3 2 plain", 1, (10, 14, 14), (128, 256, 256))
run("14^3 2 gn", 1, (10, 14, 14), (128, 256, 256), gn=True)
run("28^3 2 plain", 1, (20, 28, 28), (128, 128, 128))
run("28^3 2 gn", 1, (20, 28, 28), (128, 128, 128), gn=True)
run("N=2 ragged gn cb res", 2, (4, 9, 11), (64, 128, 128, 256), gn=True, cb=True, res=True)
