"""Development probe: per-step wall / device times of N training steps, with Python GC collections and caching-allocator
events logged -- what stalls one step in ~25 for a whole step time?   usage: train_stall_probe.py [steps] [batch]"""
import gc
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from fcwdm.optim import FusedAdamW  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 80
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
model, diffusion = bench.build_model(dev)
model.train()
diffusion.sync_timestep_check = False
opt = FusedAdamW(model, lr=1e-5, weight_decay=0.0)
g = torch.Generator().manual_seed(0)
batch = {k: torch.rand((B, 1) + bench.IMAGE, generator=g).to(dev) for k in ("t1n", "t1c", "t2w", "t2f")}
ones = torch.ones(8, device=dev)
gc_log = []
t_gc = [0.0]


def on_gc(phase, info):
    if phase == "start":
        t_gc[0] = time.perf_counter()
    else:
        gc_log.append((info["generation"], info["collected"], 1e3 * (time.perf_counter() - t_gc[0])))


gc.callbacks.append(on_gc)
events = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
walls, issue = [], []
events[0].record()
for i in range(steps):
    t0 = time.perf_counter()
    n_gc = len(gc_log)
    opt.zero_grad()
    t = torch.randint(0, diffusion.num_timesteps, (B,), device=dev)
    terms, _, _ = diffusion.training_losses(model, batch, t, model_kwargs={}, mode="i2i", contr="t1n")
    loss = (terms["mse_wav"] * ones).mean()
    loss.backward()
    opt.step()
    events[i + 1].record()
    issue.append(1e3 * (time.perf_counter() - t0))
    st = torch.cuda.memory_stats()
    walls.append((i, issue[-1], gc_log[n_gc:], torch.cuda.memory_reserved() >> 20, st.get("num_alloc_retries", 0),
                  st.get("num_device_alloc", 0), st.get("num_device_free", 0)))
    if i % 4 == 3:
        events[i - 1].synchronize()          # keep the host ~2 steps ahead, as a loop that logs its loss does
torch.cuda.synchronize()
for i, iss, gcs, res, retries, nalloc, nfree in walls:
    dt = events[i].elapsed_time(events[i + 1])
    flag = "  <-- STALL" if dt > 45 else ""
    print(f"step {i:3d}: device {dt:6.1f} ms  host issue {iss:6.1f} ms  reserved {res} MiB  retries {retries}  cudaMalloc {nalloc} cudaFree {nfree}  gc {gcs}{flag}")
