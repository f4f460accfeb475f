"""Config 2 at the three step counts SURVEY.md section 8d names: T=10 ('sampled' schedule, the fast-cwdm setting),
T=100 (diffusion_steps=1000, timestep_respacing='100') and T=1000 (scripts/sample.py default).  Batch 1,
224x224x160, CFG-W4, inputs resident, CUDA events.  Prints a markdown table (-> profiles/)."""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from fcwdm import pipeline  # noqa: E402
from guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults  # noqa: E402

dev = torch.device("cuda")
rows = []
for name, over, reps in (("T=10 sampled", dict(diffusion_steps=10, sample_schedule="sampled"), 5),
                         ("T=100 (1000 respaced to 100)", dict(diffusion_steps=1000, sample_schedule="direct",
                                                               timestep_respacing="100"), 2),
                         ("T=1000", dict(diffusion_steps=1000, sample_schedule="direct"), 1)):
    args = model_and_diffusion_defaults()
    args.update(bench.CFG_W4)
    args.update(over)
    with contextlib.redirect_stdout(io.StringIO()):
        model, diffusion = create_model_and_diffusion(**args)
    g = torch.Generator().manual_seed(0)
    for p in model.parameters():
        if float(p.detach().abs().max()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model.to(dev).eval()
    vol, noise = bench.synth_volume(1)
    vol, noise = vol.to(dev), noise.to(dev)
    run = lambda: pipeline.synthesize(diffusion, model, vol[:, 1:2], vol[:, 2:3], vol[:, 3:4], noise)
    out = run()                                        # eager first step + graph capture
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    T = diffusion.num_timesteps
    rows.append(f"| {name} | {T} | {ms:.1f} | {1e3 / ms:.3f} | {ms / T:.2f} | {bool(torch.isfinite(out).all())} |")
    print(rows[-1], flush=True)
    del model, diffusion
    torch.cuda.empty_cache()
print("\n| schedule | steps | ms / volume | volumes/s (1 x B200) | ms / step | finite |\n|---|---|---|---|---|---|")
print("\n".join(rows))
