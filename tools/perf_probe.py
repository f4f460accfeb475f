"""Development probe (not part of the product): per-kernel and per-step timings on a B200."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402

from fcwdm import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda")
    print(torch.cuda.get_device_name(0))
    # --- conv shapes of CFG-W4
    convs = [(112, 112, 80, 64, 64, 3), (112, 112, 80, 64, 8, 3), (56, 56, 40, 128, 128, 3), (56, 56, 40, 64, 64, 3),
             (56, 56, 40, 256, 64, 3), (56, 56, 40, 64, 128, 3), (28, 28, 20, 128, 128, 3), (28, 28, 20, 512, 128, 3),
             (14, 14, 10, 256, 256, 3), (14, 14, 10, 1024, 128, 3), (7, 7, 5, 256, 256, 3), (7, 7, 5, 1024, 256, 3),
             (56, 56, 40, 64, 128, 1), (14, 14, 10, 128, 256, 1)]
    for (D, H, W, ci, co, k) in convs:
        S = D * H * W
        x = torch.randn((S, max(64, ci)), device=dev).to(torch.bfloat16)
        w = torch.randn((co, ci, k, k, k), device=dev) * 0.05
        wp = ops.conv3d_pack_weights(w)
        b = torch.zeros(co, device=dev)
        y = torch.empty((S, max(8, co) if co < 64 else co), dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: ops.conv3d_cl(x, wp, b, y, (1, D, H, W), ci, co, k))
        fl = 2.0 * S * ci * co * k ** 3
        print(f"conv {D}x{H}x{W} {ci}->{co} k{k}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
    # --- elementwise kernels at full latent resolution
    S = 112 * 112 * 80
    x = torch.randn((S, 64), device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    stats = torch.empty((1, 16, 32, 2), dtype=torch.float64, device=dev)
    g = torch.ones(64, device=dev)
    bb = torch.zeros(64, device=dev)
    from fcwdm import native
    st = ops._stream(dev)
    ms = timeit(lambda: native.call("fcwdm_groupnorm_stats", ops._ptr(x), 64, ops._ptr(stats), 1, S, 64, 32, st))
    print(f"gn_stats 64ch full: {ms*1e3:.1f} us  {S*128/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: native.call("fcwdm_groupnorm_apply", ops._ptr(x), 64, ops._ptr(y), 64, ops._ptr(stats), ops._ptr(g),
                                    ops._ptr(bb), 1, S, 64, 32, 1e-5, 1, st))
    print(f"gn_apply 64ch full: {ms*1e3:.1f} us  {2*S*128/ms/1e6:.0f} GB/s")
    s2 = S // 8
    lll = torch.empty((s2, 64), dtype=torch.bfloat16, device=dev)
    hi = torch.empty((7, s2, 64), dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.dwt3d_cl(x, (1, 112, 112, 80), 64, lll, hi))
    print(f"dwt_cl 64ch full: {ms*1e3:.1f} us  {2*S*128/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: ops.idwt3d_cl(lll, hi, (1, 112, 112, 80), 64, y))
    print(f"idwt_cl 64ch full: {ms*1e3:.1f} us  {2*S*128/ms/1e6:.0f} GB/s")
    # planar DWT on 16 x (224,224,160) fp32 = 513 MB in, 513 MB out (> L2)
    v = torch.rand((1, 16, 224, 224, 160), device=dev)
    ms = timeit(lambda: ops.dwt3d_planar(v))
    print(f"dwt planar fp32 16x224x224x160: {ms*1e3:.1f} us  {2*v.numel()*4/ms/1e6:.0f} GB/s")
    bands = ops.dwt3d_planar(v)
    ms = timeit(lambda: ops.idwt3d_planar(bands))
    print(f"idwt planar fp32: {ms*1e3:.1f} us  {2*v.numel()*4/ms/1e6:.0f} GB/s")
    v1 = torch.rand((1, 1, 224, 224, 160), device=dev)
    ms = timeit(lambda: ops.dwt3d_planar(v1), iters=50)
    print(f"dwt planar fp32 1x224x224x160 (L2 resident): {ms*1e3:.1f} us  {2*v1.numel()*4/ms/1e6:.0f} GB/s")
    del v, bands
    # --- full sampling loop CFG-W4
    from guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults
    args = model_and_diffusion_defaults()
    args.update(image_size=224, in_channels=32, num_channels=64, out_channels=8, channel_mult="1,2,2,4", dims=3,
                attention_resolutions="", bottleneck_attention=False, resblock_updown=True, use_freq=True,
                use_scale_shift_norm=False, predict_xstart=True, diffusion_steps=10, sample_schedule="sampled",
                mode="i2i", num_groups=32, num_heads=1)
    model, diffusion = create_model_and_diffusion(**args)
    torch.manual_seed(0)
    for p in model.parameters():
        if float(p.detach().abs().max()) == 0.0:
            p.data.normal_(0, 0.02)
    model.to(dev).eval()
    noise = torch.randn(1, 8, 112, 112, 80, device=dev)
    cond = torch.rand(1, 24, 112, 112, 80, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = diffusion.p_sample_loop(model, noise.shape, noise=noise, cond=cond, progress=False)
        e1.record()
        torch.cuda.synchronize()
        print(f"p_sample_loop T=10 rep {rep}: device {e0.elapsed_time(e1):.1f} ms, wall {1e3*(time.time()-t0):.1f} ms, "
              f"finite={bool(torch.isfinite(out).all())}", flush=True)
    s = list(diffusion._samplers.values())[0]
    print("launches per step:", s.launches_per_step, " mem GB:", torch.cuda.max_memory_allocated() / 2**30)


if __name__ == "__main__":
    main()
