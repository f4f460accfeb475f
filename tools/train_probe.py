"""Development probe: full-size CFG-W4 training steps (training_losses forward, fcwdm backward, FusedAdamW) -- the
target command for the ncu launch list of the training path (profiles/).  usage: train_probe.py [steps] [batch]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from fcwdm import native  # noqa: E402
from fcwdm.optim import FusedAdamW  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda")
model, diffusion = bench.build_model(dev)
model.train()
diffusion.sync_timestep_check = False
opt = FusedAdamW(model, lr=1e-5, weight_decay=0.0)
g = torch.Generator().manual_seed(0)
batch = {k: torch.rand((B, 1) + bench.IMAGE, generator=g).to(dev) for k in ("t1n", "t1c", "t2w", "t2f")}
for i in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n0 = native.launch_count
    e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    opt.zero_grad()
    t = torch.randint(0, diffusion.num_timesteps, (B,), device=dev)
    e0.record()
    terms, _, _ = diffusion.training_losses(model, batch, t, model_kwargs={}, mode="i2i", contr="t1n")
    loss = (terms["mse_wav"] * torch.ones(8, device=dev)).mean()
    e1.record()
    loss.backward()
    e2.record()
    opt.step()
    e3.record()
    torch.cuda.synchronize()
    print(f"step {i}: loss {float(loss.detach()):.5f}  fwd {e0.elapsed_time(e1):.2f} ms  bwd {e1.elapsed_time(e2):.2f} ms  "
          f"opt {e2.elapsed_time(e3):.2f} ms  wall {1e3 * (time.perf_counter() - t0):.1f} ms  "
          f"launches {native.launch_count - n0}  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
