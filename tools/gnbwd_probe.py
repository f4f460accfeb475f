"""Development probe: GroupNorm backward (reduce + apply) and the column-sum pass at the training shapes
(FCWDM_TR_BLOCKS_PER_SM sets the grid cap of these slab kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402
from perf_probe import timeit  # noqa: E402

dev = torch.device("cuda")
print("FCWDM_TR_BLOCKS_PER_SM =", os.environ.get("FCWDM_TR_BLOCKS_PER_SM", "8 (default)"))
for (N, C, S, G) in [(2, 64, 112 * 112 * 80, 32), (2, 128, 56 * 56 * 40, 32), (2, 128, 28 * 28 * 20, 32), (2, 256, 14 * 14 * 10, 32)]:
    x = torch.randn((N * S, C), device=dev).to(torch.bfloat16)
    dy = torch.randn((N * S, C), device=dev).to(torch.bfloat16)
    dx = torch.empty_like(x)
    y = torch.empty_like(x)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    stats = torch.empty((N, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device=dev)
    ops.groupnorm_stats(x, stats, N, S, C, G)
    emb = torch.zeros((N, C), device=dev)
    bias = torch.zeros(C, device=dev)
    t_bwd = timeit(lambda: ops.groupnorm_bwd(x, dy, stats, gamma, beta, dx, dg, db, N, S, C, G), iters=10)
    t_col = timeit(lambda: ops.colsum_cl(dx, N, S, C, out_sample=emb, out_total=bias), iters=10)
    t_fwd = timeit(lambda: ops.groupnorm_silu(x, y, stats, gamma, beta, N, S, C, G, 1e-5, True, have_stats=True), iters=10)
    gb = N * S * C * 2 / 1e9
    print(f"N{N} C{C} S{S}: gn_bwd {t_bwd*1e3:7.1f} us ({5*gb/t_bwd:6.0f} GB/s of 5 tensor passes)   colsum {t_col*1e3:6.1f} us "
          f"({gb/t_col:6.0f} GB/s)   gn_fwd apply {t_fwd*1e3:6.1f} us ({2*gb/t_fwd:6.0f} GB/s)", flush=True)
