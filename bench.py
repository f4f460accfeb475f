#!/usr/bin/env python
"""Headline benchmark of the fast-cwdm hot path on B200 (contract: one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[1]): full respaced `p_sample_loop` synthesising one modality from three, batch 1,
224x224x160 BraTS-shaped synthetic volume, WavUNetModel CFG-W4 (the only WavUNetModel configuration that runs at
this size) with random-init weights (zero-init convs re-randomised), T = 10 steps of the
`sampled` schedule (the fast-cwdm setting, tr_job.yml:122).  One "step" = one whole volume: conditioning DWTs,
T denoising steps (each 74 tcgen05 conv3d + 65 GroupNorm/SiLU + 20 DWT/IDWT + 1 fused posterior update), final
IDWT + clamp + mask.

  value : volumes/s with the inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e   : the same through the public per-volume API with pinned HOST buffers: H2D of the three conditioning
          modalities and the CPU-drawn noise (scripts/sample.py:100) and D2H of the synthesised volume are inside
          the timed region.
  roofline : dominant kernel = conv3d 64->64 @112x112x80 (10 launches per denoising step, 60% of the FLOPs), timed
          alone with CUDA events; peak = measured cuBLAS bf16 burst figure (MEASURED_PEAKS.json).
  cpu_baseline : the oracle port (oracle/*.py, torch CPU fp32 restatement of the reference) on the host cores.

`--impl reference` times the reference's algorithm on the CPU (the reference is pure Python and does not travel to
the GPU box, so the oracle port stands in; /root/reference is never read here).
Multi-GPU: one process per GPU (torchrun), volumes partitioned by rank, no collective on the data path; the only
communication is the timing barrier / max-reduction.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]

if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"]:
    # the reference moves its Haar band matrices to the GPU whenever one is visible, even for CPU inputs
    # (DWT_IDWT/DWT_IDWT_layer.py:505-511): the CPU arm must not see a device
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np  # noqa: E402
import torch  # noqa: E402

T_STEPS = 10
LATENT = (112, 112, 80)
IMAGE = (224, 224, 160)
CFG_W4 = dict(image_size=224, in_channels=32, num_channels=64, out_channels=8, channel_mult="1,2,2,4", dims=3,
              attention_resolutions="", bottleneck_attention=False, resblock_updown=True, use_freq=True,
              use_scale_shift_norm=False, predict_xstart=True, diffusion_steps=T_STEPS, sample_schedule="sampled",
              mode="i2i", num_groups=32, num_heads=1, num_res_blocks=2)
CONV_FLOP_PER_STEP = 3665.0e9            # 74 conv3d per denoiser call, CFG-W4, batch 1
WORKLOAD = "p_sample_loop T=10 'sampled' i2i, WavUNetModel CFG-W4 (1,2,2,4), batch 1, 224x224x160 volume"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        """enabled=False on ranks > 0: the JSON line reports rank 0's GPU, and N concurrent nvidia-smi loops slow each
        other down enough to miss a short timed region."""
        self.index, self.rows, self.proc, self.t_mark, self.enabled = index, [], None, 0.0, enabled

    def mark(self):
        """The timed region starts now: only samples arriving from here on are reported.  (The sampler is started
        before the warm-up steps: nvidia-smi needs 0.3-1 s to produce its first line, longer than a short timed region.)"""
        self.t_mark = time.time()

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            # let the sample covering the end of the region arrive; on a busy 8-GPU box nvidia-smi's loop can take a few
            # hundred ms per line, so wait (bounded) until at least one line has arrived after mark()
            deadline = time.time() + 1.5
            time.sleep(0.12)
            while time.time() < deadline and not any(t >= self.t_mark for t, _ in self.rows):
                time.sleep(0.05)
            self.proc.terminate()
        rows = [r for t, r in self.rows if t >= self.t_mark]
        during = True
        if not rows and self.rows:              # region shorter than the sampling period: fall back to the warm-up samples
            rows, during = [r for _, r in self.rows[-3:]], False
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "during_timed_region": during}


def ncu_traffic(name):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summary (profiles/)."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return None
    tot = 0.0
    for line in open(path):
        parts = line.split()
        if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[1] == "Mbyte":
            tot += float(parts[2]) * 1e6
    return tot or None


CFG_UNET = dict(CFG_W4, use_freq=False, channel_mult="1,2,2,4,4", resample_2d=False)      # run.sh:59-66,109-133
UNET_FLOP_PER_STEP = 7168.7e9           # SURVEY.md section 8f row 1 [probed]: plain UNetModel, 81.5 M parameters
MODEL = {"name": "wunet"}


def build_model(device):
    from guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults
    args = model_and_diffusion_defaults()
    args.update(CFG_UNET if MODEL["name"] == "unet" else CFG_W4)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model, diffusion = create_model_and_diffusion(**args)
    g = torch.Generator().manual_seed(0)
    for p in model.parameters():                       # re-randomise the zero-initialised convs
        if float(p.detach().abs().max()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model.to(device).eval()
    return model, diffusion


def synth_volume(seed, batch=1):
    """BraTS-shaped synthetic case: 4 modalities in [0,1) with a zero background border (exercises the mask)."""
    g = torch.Generator().manual_seed(seed)
    vol = torch.rand((batch, 4) + IMAGE, generator=g)
    vol[:, :, :8] = 0
    vol[:, :, :, :8] = 0
    noise = torch.randn((batch, 8) + LATENT, generator=g)
    return vol, noise


# ----------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm (torch CPU fp32), bounded sample
# ----------------------------------------------------------------------------------------------------
def oracle_state():
    from oracle import wunet as ow
    from guided_diffusion.wunet import WavUNetModel
    m = WavUNetModel(image_size=224, in_channels=32, model_channels=64, out_channels=8, num_res_blocks=2,
                     attention_resolutions=(), channel_mult=(1, 2, 2, 4), dims=3, num_groups=32,
                     bottleneck_attention=False, resblock_updown=True, use_freq=True)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    return ow.tie_output_blocks(ow.seeded_state_dict(shapes, seed=0, std=0.02), 4)


class ReferenceStep:
    """One denoising step of config 2 (GaussianDiffusion.p_sample: WavUNetModel forward + IDWT / clamp / DWT + posterior
    + noise, gaussian_diffusion.py:529-574) on the host CPU.  kind='reference': the UNMODIFIED reference's own classes
    (oracle/_ref staged at build time, or the mount of the build container) through oracle/ref_shims; kind='port': the
    oracle restatement, only when no copy of the reference is present."""

    def __init__(self):
        from oracle import ref_shims
        torch.set_num_threads(os.cpu_count() or 1)
        self.sd = oracle_state()
        self.kind = "reference" if ref_shims.reference_available() else "port"
        self.ctx = None
        if self.kind == "reference":
            import contextlib
            import io
            self.ctx = ref_shims.reference_modules()
            ref = self.ctx.__enter__()
            args = ref.script_util.model_and_diffusion_defaults()
            args.update(CFG_W4)
            with contextlib.redirect_stdout(io.StringIO()):          # the reference prints its schedule diagnostics
                self.model, self.diffusion = ref.script_util.create_model_and_diffusion(**args)
            self.model.load_state_dict(self.sd, strict=True)
            self.model.eval()
            self.root = ref_shims.REFERENCE_ROOT

    def close(self):
        if self.ctx is not None:
            self.ctx.__exit__(None, None, None)
            self.ctx = None

    def seconds(self, depth):
        """Wall time of one step on a `depth`-plane slab of the 112-plane latent (depth % 16 == 0)."""
        g = torch.Generator().manual_seed(1)
        x = torch.randn((1, 8, depth) + LATENT[1:], generator=g)
        cond = torch.rand((1, 24, depth) + LATENT[1:], generator=g)
        t = torch.tensor([T_STEPS - 1])
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                out = self.diffusion.p_sample(self.model, x, t, clip_denoised=True, model_kwargs={}, cond=cond)
            else:
                from oracle import diffusion as od
                from oracle import wunet as ow
                tab = od.Tables(od.named_beta_schedule("linear", T_STEPS, "sampled"))
                net = lambda xin, tt: ow.wunet_forward(self.sd, xin, tt, model_channels=64, channel_mult=(1, 2, 2, 4))
                out = od.p_sample(tab, net, x, t, cond=cond)
        dt = time.perf_counter() - t0
        assert bool(torch.isfinite(out["sample"]).all())
        return dt


def cpu_baseline(budget_s=25.0):
    """cpu_baseline of the GPU arm's line: a bounded sample (one denoising step; a 16-plane slab first, the full 112
    planes when that fits the budget) of the reference on the host cores, scaled to volumes/s."""
    ref = ReferenceStep()
    try:
        depth = 16                                           # 1/7 of the latent depth; must be divisible by 2^4
        dt = ref.seconds(depth)
        if dt * (LATENT[0] / depth) <= budget_s:             # fast host: time the full-depth step as well
            depth = LATENT[0]
            dt = ref.seconds(depth)
    finally:
        ref.close()
    step_full = dt * (LATENT[0] / depth)
    return {"value": 1.0 / (T_STEPS * step_full), "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": ref.kind,
            "sample": f"one denoising step of the {'unmodified reference (GaussianDiffusion.p_sample, fp32 torch CPU)' if ref.kind == 'reference' else 'oracle port'} "
                      f"on a {depth}-plane slab of the 112-plane latent, {dt:.2f} s, extrapolated x{LATENT[0] // depth} "
                      f"x T={T_STEPS} identical steps to one volume; host {os.cpu_count()} logical cores"}


def run_reference(args):
    """The reference arm: rank 0 alone times the reference's own CPU implementation (other ranks exit without work).
    One "step" of this arm = ONE denoising step (1/T of a volume): `ms_per_step` is that measured unit, and `value`
    (volumes/s) = 1 / (T x step time) is an extrapolation over the T identical steps of a volume -- stated in config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = ReferenceStep()
    try:
        total = args.steps + args.warmup
        depth = 16
        probe = ref.seconds(depth)                           # also serves as the page-in warm-up
        if probe * (LATENT[0] / depth) * total <= 240.0:
            depth = LATENT[0]
        for _ in range(args.warmup):
            ref.seconds(depth)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ref.seconds(depth)
        per_step = (time.perf_counter() - t0) / max(1, args.steps)
    finally:
        ref.close()
    scale = LATENT[0] / depth
    vol_s = 1.0 / (T_STEPS * per_step * scale)
    what = ("unmodified reference (GaussianDiffusion.p_sample + WavUNetModel, fp32 torch CPU)" if ref.kind == "reference"
            else "oracle port (no copy of the reference present)")
    sample = (f"each timed step = one denoising step of the {what} on a {depth}-plane slab of the 112-plane latent "
              f"({per_step:.2f} s); volumes/s = 1 / (T={T_STEPS} x {scale:.0f} x step time), an extrapolation over identical steps")
    n_gpus = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    print(json.dumps({"impl": "reference", "metric": "sampled 224x224x160 volumes/sec", "value": vol_s,
                      "unit": "volumes/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "arm": what,
                                 "step_unit": f"one denoising step = 1/{T_STEPS} of a volume"
                                              + ("" if depth == LATENT[0] else f", on {depth}/{LATENT[0]} of the depth"),
                                 "host_arm_note": ("ONE host runs this arm whatever N is: at N > 1 a GPU/CPU ratio divides N "
                                                   "GPUs by one host's cores and says nothing about scaling")},
                      "cpu_baseline": {"value": vol_s, "unit": "volumes/s", "cores": torch.get_num_threads(),
                                       "kind": ref.kind, "sample": sample},
                      "e2e": {"value": vol_s, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def time_dominant_kernel(device, iters=20):
    """conv3d 64->64 3x3x3 @112x112x80 alone, in the two forms the denoising step actually launches (5 + 5 of its 10
    full-resolution launches; `conv3d_pair_kernel<64, GN_IN>`): (A) a ResBlock's first conv = fused input GroupNorm+SiLU
    + bias + per-sample timestep embedding + output statistics; (B) its second conv = fused input GroupNorm+SiLU + bias
    + residual + output statistics.  CUDA events on the launching stream; a 256 MB write flushes L2 before every timed
    launch.  Returns ({variant: ms}, flop per launch)."""
    from fcwdm import ops
    S = LATENT[0] * LATENT[1] * LATENT[2]
    G = 32
    x = torch.randn((S, 64), device=device).to(torch.bfloat16)
    res = torch.randn((S, 64), device=device).to(torch.bfloat16)
    w = torch.randn((64, 64, 3, 3, 3), device=device) * 0.02
    wp = ops.conv3d_pair_pack_weights(w)
    b = torch.randn(64, device=device) * 0.1
    emb = torch.randn((1, 64), device=device) * 0.1
    gamma, beta = torch.rand(64, device=device) + 0.5, torch.randn(64, device=device) * 0.1
    stats = torch.empty((1, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device=device)
    ops.groupnorm_stats(x, stats, 1, S, 64, G)
    ostats = torch.zeros((1, ops.GN_STAT_REPLICAS, G, 2), dtype=torch.float64, device=device)
    y = torch.empty((S, 64), dtype=torch.bfloat16, device=device)
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=device)
    gn_in = (stats, gamma, beta, G, 1e-5)
    variants = {"gn_in+emb+stats": dict(gn_in=gn_in, chan_bias=emb, gn_stats=ostats, gn_groups=G),
                "gn_in+residual+stats": dict(gn_in=gn_in, residual=res, gn_stats=ostats, gn_groups=G)}
    out = {}
    for name, kw in variants.items():
        for _ in range(3):
            ops.conv3d_pair_cl(x, wp, b, y, (1,) + LATENT, 64, 64, **kw)
        total = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv3d_pair_cl(x, wp, b, y, (1,) + LATENT, 64, 64, **kw)
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        out[name] = total / iters
    return out, 2.0 * S * 64 * 64 * 27


def time_haar(device, iters=10):
    """DWT_3D / IDWT_3D standalone bandwidth (BASELINE metric part 2): 16 x 224x224x160 fp32 = 514 MB in + 514 MB
    out per launch (>> L2)."""
    from fcwdm import ops
    v = torch.rand((1, 16) + IMAGE, device=device)
    out = {}
    for name, fn in (("dwt3d", lambda: ops.dwt3d_planar(v)),):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        e1.synchronize()
        out[name] = 2.0 * v.numel() * 4 / (e0.elapsed_time(e1) / iters * 1e-3) / 1e9
    bands = ops.dwt3d_planar(v)
    for _ in range(3):
        ops.idwt3d_planar(bands)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.idwt3d_planar(bands)
    e1.record()
    e1.synchronize()
    out["idwt3d"] = 2.0 * v.numel() * 4 / (e0.elapsed_time(e1) / iters * 1e-3) / 1e9
    return out


def run_gpu(args):
    import torch.distributed as dist
    from fcwdm import native, pipeline
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fcwdm hot path has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = measured_peaks()
    model, diffusion = build_model(device)

    # every rank owns its own volumes (weak scaling): volume index = rank * (K + W) + i, no collective
    n_local = args.steps + args.warmup
    my_ids = pipeline.shard_indices(n_local * world, rank, world)
    host_vol, host_noise = synth_volume(1000 + my_ids[0], args.batch)
    host_vol = host_vol.pin_memory()
    host_noise = host_noise.pin_memory()
    host_out = torch.empty((args.batch,) + IMAGE[:2] + (155,), dtype=torch.float32).pin_memory()
    dev_vol = host_vol.to(device)
    dev_noise = host_noise.to(device)

    def volume_resident():
        return pipeline.synthesize(diffusion, model, dev_vol[:, 1:2], dev_vol[:, 2:3], dev_vol[:, 3:4], dev_noise)

    # end-to-end: host-resident inputs / outputs through the package's own streaming API (fcwdm.pipeline.VolumeStream):
    # the H2D copy of volume i+1 and the D2H copy of volume i-1 run on a copy stream underneath the compute of volume i
    # (all inside the timed region).
    stream = pipeline.VolumeStream(diffusion, model, device)
    copy_stream = stream.copy_stream

    def volume_e2e():
        return stream.submit(host_vol, host_noise, host_out, next_case=(host_vol, host_noise))

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, k, w, join=None):
        for _ in range(w):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        if join is not None:
            torch.cuda.current_stream(device).wait_stream(join)   # side-stream copies count towards the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    torch.manual_seed(1234 + rank)
    clocks = ClockSampler(local, enabled=(rank == 0))
    warm = max(args.warmup, 3)
    clocks.start()
    timed(volume_resident, 0, warm)                      # lazy init, weight packing, CUDA-graph capture
    clocks.mark()
    ms_res = timed(volume_resident, args.steps, 0)
    clk = clocks.stop()
    sampler = list(diffusion._samplers.values())[0]
    launches_per_volume = sampler.launches_per_step * T_STEPS + 3 + 2 + 1     # + cond DWTs, x_T/cond layout, image
    ms_e2e = timed(volume_e2e, args.steps, 1, join=copy_stream)
    out = volume_resident()
    finite = bool(torch.isfinite(out).all())

    result = None
    flop_step = UNET_FLOP_PER_STEP if MODEL["name"] == "unet" else CONV_FLOP_PER_STEP
    if rank == 0:
        vols = args.steps * world * args.batch
        value = vols / (ms_res * 1e-3)
        e2e = vols / (ms_e2e * 1e-3)
        h2d = host_vol[:, 1:].numel() * 4 + host_noise.numel() * 4       # 3 modalities actually used + noise
        d2h = host_out.numel() * 4
        result = {
            "metric": "sampled 224x224x160 volumes/sec", "value": value, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (WORKLOAD if MODEL["name"] == "wunet" else WORKLOAD.replace(
                           "WavUNetModel CFG-W4 (1,2,2,4)", "plain UNetModel (run.sh: 1,2,2,4,4, resample_2d=False)")
                                    ).replace("batch 1,", f"batch {args.batch},"),
                       "parallelism": f"volumes sharded over {world} GPU(s), no collective",
                       "l2": "per-step activations ~10 GB >> 126 MB L2 (no flush needed)", "T": T_STEPS, "volumes_per_step": args.batch,
                       "denoiser_gflop_per_step": flop_step / 1e9, "peaks": peaks["src"],
                       "output_finite": finite},
            "e2e": {"value": e2e, "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_per_volume * args.steps),
            "clocks": clk,
            "step_tflops": flop_step * T_STEPS * vols / (ms_res * 1e-3) / 1e12,
        }
    if world > 1:
        dist.barrier()
    # ---- config 3 (8 volumes per GPU) and config 4 (training step, data-parallel over the same N ranks), every rank
    sec_batch8 = sec_train = None
    if MODEL["name"] == "wunet" and args.batch == 1 and not args.no_secondary:
        sec_batch8 = measure_batch8(model, diffusion, device, world, rank, barrier)
        for smp in list(diffusion._samplers.values()):       # drop the captured graphs and their private pools
            smp.release()
        diffusion._samplers.clear()
        stream._slots = None
        torch.cuda.empty_cache()
        sec_train = measure_train(device, world, rank, local, steps=10, warm=3, batch=2)
    if rank == 0:
        ms_v, flop = time_dominant_kernel(device)
        ms_k = sum(ms_v.values()) / len(ms_v)                 # the step launches the two variants 5 : 5
        ach = flop / (ms_k * 1e-3) / 1e12
        result["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                              "frac": ach / peaks["tf_burst"], "traffic": ncu_traffic("r02_conv_pair64_gnin_ncu.txt"),
                              "traffic_note": "dram__bytes_read+write of one launch of the residual variant, ncu --set full "
                                              "behind a 256 MB L2 flush (profiles/r02_conv_pair64_gnin_ncu.txt); "
                                              "algorithmic 257 MB (385 MB with the residual operand)",
                              "kernel": "conv3d_pair_kernel<64, GN_IN> (cta_group::2, kd-fused) 64->64 @112x112x80, the two "
                                        "in-step forms weighted 5:5 (fused input GroupNorm+SiLU, bias, timestep embedding "
                                        "or residual, output statistics)",
                              "us_per_launch": ms_k * 1e3, "us_by_variant": {k: v * 1e3 for k, v in ms_v.items()},
                              "flop_per_launch": flop, "l2": "256 MB flush before every timed launch",
                              "peak_kind": "burst (kernel timed alone)",
                              # the whole denoising step against the sustained peak: the number that bounds volumes/s
                              "step_tflops": result["step_tflops"], "step_peak": peaks["tf_sustained"],
                              "step_frac": result["step_tflops"] / peaks["tf_sustained"]}
        hb = time_haar(device)
        result["secondary"] = {"dwt3d_gbs": hb["dwt3d"], "idwt3d_gbs": hb["idwt3d"], "hbm_peak_gbs": peaks["hbm"],
                               "dwt3d_frac": hb["dwt3d"] / peaks["hbm"], "idwt3d_frac": hb["idwt3d"] / peaks["hbm"],
                               "workload": "16 x 224x224x160 fp32 planar, 2*numel*4 bytes per launch",
                               "batch8": sec_batch8, "train": sec_train}
        if world == 1 and not args.no_cpu_baseline and MODEL["name"] == "wunet":
            result["cpu_baseline"] = cpu_baseline_subprocess()
        else:
            result["cpu_baseline"] = None
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# training arm (BASELINE config 4): one "step" = training_losses forward + fcwdm backward + fused AdamW on a batch of
# B synthetic cases per GPU; data-parallel over N GPUs with ONE bucketed NCCL gradient all-reduce per step
# ----------------------------------------------------------------------------------------------------
def time_wgrad_kernel(device, iters=10):
    """conv3d wgrad 64->64 3x3x3 @112x112x80 alone (the dominant training kernel), CUDA events, L2 flush between."""
    from fcwdm import ops
    S = LATENT[0] * LATENT[1] * LATENT[2]
    x = torch.randn((S, 64), device=device).to(torch.bfloat16)
    dy = torch.randn((S, 64), device=device).to(torch.bfloat16)
    dw = torch.zeros((64, 64, 3, 3, 3), device=device)
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=device)
    for _ in range(3):
        ops.conv3d_wgrad(x, dy, dw, (1,) + LATENT, 64, 64, 3, accumulate=False)
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3d_wgrad(x, dy, dw, (1,) + LATENT, 64, 64, 3, accumulate=False)
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters, 2.0 * S * 64 * 64 * 27


def train_measure(device, world, rank, local, steps, warm, batch, want_roofline=False):
    """Config 4 on this rank's GPU: `warm` + `steps` training steps (training_losses forward + fcwdm backward + fused AdamW
    on `batch` synthetic cases; under world > 1 one bucketed NCCL gradient all-reduce per step), resident and end to end
    (pinned-host batch uploaded and the loss read back every step).  Needs the default process group when world > 1.
    Returns the result dict on rank 0, None elsewhere."""
    import torch.distributed as dist
    from fcwdm import ddp, native
    from fcwdm.optim import FusedAdamW
    peaks = measured_peaks()
    model, diffusion = build_model(device)
    model.train()
    diffusion.sync_timestep_check = False        # validate t on the device: no host read-back stalling the launch queue
    sync = None
    if world > 1:
        ddp.broadcast_parameters(model)
        sync = ddp.attach(model)
    opt = FusedAdamW(model, lr=1e-5, weight_decay=0.0)                      # run.sh:60 lr, train_util.py:75-82
    B = batch
    g = torch.Generator().manual_seed(100 + rank)                            # per-rank data (seed + rank)
    host = {k: torch.rand((B, 1) + IMAGE, generator=g).pin_memory() for k in ("t1n", "t1c", "t2w", "t2f")}
    dev_batch = {k: v.to(device) for k, v in host.items()}
    slot = {k: torch.empty_like(v) for k, v in dev_batch.items()}
    ones = torch.ones(8, device=device)
    rng = torch.Generator(device=device).manual_seed(7 + rank)
    last = {}

    def step(batch):
        opt.zero_grad()
        t = torch.randint(0, diffusion.num_timesteps, (B,), device=device, generator=rng)
        terms, _, _ = diffusion.training_losses(model, batch, t, model_kwargs={}, mode="i2i", contr="t1n")
        loss = (terms["mse_wav"] * ones).mean()                              # train_util.py:447-449
        loss.backward()
        opt.step()
        last["loss"] = loss.detach()

    def step_resident():
        step(dev_batch)

    # e2e: every step uploads ITS batch from pinned host memory and its loss is read back on the host; the upload of
    # step i+1 and the read-back of step i-1 ride on a copy stream underneath step i's compute (two device batch slots,
    # two pinned loss cells), the way fcwdm.pipeline.VolumeStream does it for sampling
    copy_stream = torch.cuda.Stream(device)
    slots = [slot, {k: torch.empty_like(v) for k, v in dev_batch.items()}]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    slot_free = [torch.cuda.Event() for _ in range(2)]
    LAG = int(os.environ.get("FCWDM_BENCH_LAG", "2"))   # the host reads step i's loss while steps i+1 .. i+LAG are queued / running
    loss_cell = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(LAG + 1)]
    loss_done = [torch.cuda.Event() for _ in range(LAG + 1)]
    pipe = {"i": 0, "uploaded": -1}

    def upload(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot_free[s])                             # the step that last read this slot is done
            for k in slots[s]:
                slots[s][k].copy_(host[k], non_blocking=True)                # H2D of step i's batch (pinned)
            h2d_done[s].record(copy_stream)
        pipe["uploaded"] = i

    def step_e2e():
        i = pipe["i"]
        s = i % 2
        cur = torch.cuda.current_stream(device)
        if pipe["uploaded"] < i:
            upload(i)
        cur.wait_event(h2d_done[s])
        upload(i + 1)                                                        # next step's batch, under this step's compute
        step(slots[s])
        slot_free[s].record(cur)
        c = i % (LAG + 1)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot_free[s])
            loss_cell[c].copy_(last["loss"], non_blocking=True)              # D2H of this step's result
            loss_done[c].record(copy_stream)
        if i >= LAG:
            # every step's loss reaches the host, LAG steps late: the launch queue keeps LAG steps of work while the host
            # waits (one step of slack starved the GPU when two ranks share the host: 94 vs 113 samples/s at N = 2)
            p = (i - LAG) % (LAG + 1)
            loss_done[p].synchronize()
            last["host_loss"] = float(loss_cell[p])
        pipe["i"] = i + 1

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        e0.record()
        for _ in range(k):
            fn()
            if os.environ.get("FCWDM_BENCH_STEP_TRACE") == "1":      # development: per-step device times on stderr
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
        e1.record()
        barrier()
        if marks and rank == 0:
            ts = [e0.elapsed_time(m) for m in marks]
            print(fn.__name__, "per-step ms:", [round(b - a, 1) for a, b in zip([0.0] + ts[:-1], ts)], file=sys.stderr)
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    import gc
    gc.collect()
    gc.freeze()                                   # as guided_diffusion.train_util.TrainLoop.run_loop does (a generation-2
    #                                               collection over the whole heap is a 32 ms host pause: one starved step)
    clocks = ClockSampler(local, enabled=(rank == 0))
    clocks.start()
    timed(step_resident, 0, warm)
    clocks.mark()
    n0 = native.launch_count
    ms_res = timed(step_resident, steps, 0)
    launches = native.launch_count - n0
    for ev in slot_free:
        ev.record(torch.cuda.current_stream(device))
    ms_e2e = timed(step_e2e, steps, int(os.environ.get("FCWDM_BENCH_E2E_WARM", str(max(1, warm)))))   # its own warm-up
    clk = clocks.stop()                           # covers the resident and the end-to-end timed regions
    loss_done[(pipe["i"] - 1) % (LAG + 1)].synchronize()
    last["host_loss"] = float(loss_cell[(pipe["i"] - 1) % (LAG + 1)])
    finite = bool(torch.isfinite(last["loss"])) and last["host_loss"] == float(last["loss"])
    if rank != 0:
        return None
    samples = steps * world * B
    h2d = sum(v.numel() * 4 for v in host.values())
    model_name = "UNetModel" if MODEL["name"] == "unet" else "WavUNetModel"
    flop_step = 3.0 * (UNET_FLOP_PER_STEP if MODEL["name"] == "unet" else CONV_FLOP_PER_STEP) * B   # fwd + dgrad + wgrad
    result = {
        "metric": f"{model_name} training samples/sec", "value": samples / (ms_res * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_res / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"training step (training_losses i2i + backward + AdamW), {model_name} "
                               f"{'(run.sh: 1,2,2,4,4)' if MODEL['name'] == 'unet' else 'CFG-W4'}, "
                               f"batch {B} x 224x224x160 per GPU, bf16 compute / fp32 master",
                   "parallelism": f"dp{world}: one bucketed NCCL gradient all-reduce (mean) per step" if world > 1
                   else "single GPU", "l2": "per-step activations ~6 GB >> 126 MB L2 (no flush needed)",
                   "allreduce_buckets": sync.launched if sync else 0, "peaks": peaks["src"], "loss_finite": finite,
                   "final_loss": float(last["loss"])},
        "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clk,
        "step_tflops": flop_step * steps * world / (ms_res * 1e-3) / 1e12,
    }
    result["step_frac"] = result["step_tflops"] / world / peaks["tf_sustained"]
    if want_roofline:
        ms_k, flop = time_wgrad_kernel(device)
        ach = flop / (ms_k * 1e-3) / 1e12
        result["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                              "frac": ach / peaks["tf_burst"], "traffic": None,
                              "kernel": "conv3d_wgrad64_kernel (tap-packed M=128/64 x N=192) + finalize, 64->64 @112x112x80",
                              "us_per_launch": ms_k * 1e3, "flop_per_launch": flop}
    return result


def measure_train(device, world, rank, local, steps, warm, batch):
    """`secondary.train` of the default line (BASELINE config 4 at the same N): a short run of train_measure."""
    r = train_measure(device, world, rank, local, steps, warm, batch)
    if r is None:
        return None
    keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "e2e", "clocks", "gpu_launches",
            "step_tflops", "step_frac")
    out = {k: r[k] for k in keep}
    out["config"] = r["config"]
    return out


def measure_batch8(model, diffusion, device, world, rank, barrier, steps=2, warm=2, batch=8):
    """`secondary.batch8` (BASELINE config 3): `batch` volumes per GPU per p_sample_loop, inputs resident, every rank."""
    import torch.distributed as dist
    from fcwdm import pipeline
    vol, noise = synth_volume(2000 + rank, batch)
    vol, noise = vol.to(device), noise.to(device)

    def run():
        return pipeline.synthesize(diffusion, model, vol[:, 1:2], vol[:, 2:3], vol[:, 3:4], noise)

    for _ in range(warm):
        run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = run()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    finite = bool(torch.isfinite(out).all())
    for smp in [s for s in diffusion._samplers.values() if s.N == batch]:
        smp.release()                                        # the batch-8 graph's private pool is not needed any more
    return {"value": steps * world * batch / (ms * 1e-3), "unit": "volumes/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "volumes_per_step": batch, "ms_per_step": ms / steps, "output_finite": finite,
            "workload": f"p_sample_loop T={T_STEPS}, {batch} volumes per GPU per call, inputs resident"}


def cpu_baseline_subprocess():
    """cpu_baseline of the default line: `bench.py --impl reference --steps 1 --warmup 0` in a child process with CUDA
    hidden (the reference moves its band matrices to any visible GPU), i.e. ONE denoising step of the unmodified
    reference on the host cores after a 16-plane probe (about 10-20 s), scaled to volumes/s."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, timeout=600, env=env)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as exc:
        return {"value": None, "unit": "volumes/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": f"failed: {exc}"}


def run_train(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fcwdm training path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    result = train_measure(device, world, rank, local, args.steps, max(args.warmup, 3), args.batch, want_roofline=True)
    if rank == 0:
        result["cpu_baseline"] = None
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fcwdm", choices=["fcwdm", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip secondary.batch8 / secondary.train (configs 3, 4)")
    ap.add_argument("--batch", type=int, default=1, help="volumes per GPU per step (BASELINE config 3 uses 8)")
    ap.add_argument("--model", default="wunet", choices=["wunet", "unet"],
                    help="wunet = WavUNetModel CFG-W4 (the headline); unet = the plain UNetModel of run.sh (1,2,2,4,4)")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"],
                    help="sample = BASELINE config 2/3 (the headline, default); train = config 4 training step")
    args = ap.parse_args()
    MODEL["name"] = args.model
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
