"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol, the
drop-in model has the reference's state-dict keys, the host-side schedule logic matches the reference fixtures,
and the product fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from fcwdm import native
    lib = native.load()
    header = open(os.path.join(ROOT, "include", "fcwdm.h")).read()
    declared = set(re.findall(r"\b(fcwdm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) == len(native.PROTOTYPES) >= 19
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fcwdm.h but not exported"
        assert name in native.PROTOTYPES, f"{name} has no ctypes prototype"
    # INTEGRATION.md section 3 names the reference interface every entry point replaces: none may be missing there
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    undocumented = sorted(n for n in declared if f"`{n}`" not in doc)
    assert not undocumented, f"entry points without a row in INTEGRATION.md: {undocumented}"
    assert lib.fcwdm_version() == 100
    assert lib.fcwdm_conv3d_packed_elems(64, 32, 3) == 27 * 64 * 64
    assert lib.fcwdm_conv3d_packed_elems(8, 64, 3) == 27 * 16 * 64
    assert lib.fcwdm_conv3d_packed_elems(8, 64, 2) == -1


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call (status codes, never exceptions across the ABI)."""
    import ctypes
    from fcwdm import native
    lib = native.load()
    z = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)
    assert lib.fcwdm_dwt3d_fwd(one, one, 0, 1, 1, 3, 4, 4, 0, 0, 0, 0, 0, 1.0, z) == -2      # odd depth: unsupported
    assert b"even" in lib.fcwdm_last_error()
    assert lib.fcwdm_dwt3d_fwd(z, z, 0, 1, 1, 4, 4, 4, 0, 0, 0, 0, 0, 1.0, z) == -1        # null pointer
    assert lib.fcwdm_dwt3d_fwd(one, one, 7, 1, 1, 4, 4, 4, 0, 0, 0, 0, 0, 1.0, z) == -1      # bad dtype
    assert lib.fcwdm_dwt3d_fwd(z, z, 0, 0, 1, 4, 4, 4, 0, 0, 0, 0, 0, 1.0, z) == 0         # empty batch is a no-op
    assert lib.fcwdm_conv3d_fwd(one, 64, one, z, z, 0, z, 0, one, 64, z, 0, 1, 4, 4, 4, 64, 64, 5, z) == -2   # ksize 5
    assert lib.fcwdm_conv3d_fwd(one, 32, one, z, z, 0, z, 0, one, 64, z, 0, 1, 4, 4, 4, 64, 64, 3, z) == -1   # x_ld < Cin_p
    assert lib.fcwdm_groupnorm_stats(one, 64, one, 1, 10, 64, 7, z) == -1                  # C % G != 0


def test_no_cpu_fallback():
    from DWT_IDWT.DWT_IDWT_layer import DWT_3D, IDWT_3D
    from fcwdm import FcwdmError
    from guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults
    with pytest.raises(FcwdmError):
        DWT_3D("haar")(torch.zeros(1, 1, 4, 4, 4))
    with pytest.raises(FcwdmError):
        IDWT_3D("haar")(*[torch.zeros(1, 1, 2, 2, 2)] * 8)
    with pytest.raises(NotImplementedError):
        DWT_3D("db2")
    args = model_and_diffusion_defaults()
    args.update(image_size=16, in_channels=32, num_channels=32, out_channels=8, channel_mult="1,2", dims=3,
                attention_resolutions="", bottleneck_attention=False, resblock_updown=True, use_freq=True,
                use_scale_shift_norm=False, predict_xstart=True, diffusion_steps=10, sample_schedule="sampled", mode="i2i")
    model, diffusion = create_model_and_diffusion(**args)
    with torch.no_grad(), pytest.raises(FcwdmError):
        model(torch.zeros(1, 32, 4, 4, 4), torch.zeros(1, dtype=torch.long))
    with pytest.raises(FcwdmError):                   # training (autograd) path: CUDA only as well, no CPU fallback
        model(torch.zeros(1, 32, 4, 4, 4), torch.zeros(1, dtype=torch.long))


def test_unsupported_flags_fail_at_construction():
    from guided_diffusion.wunet import WavUNetModel
    base = dict(image_size=16, in_channels=32, model_channels=32, out_channels=8, num_res_blocks=2,
                attention_resolutions=(), channel_mult=(1, 2), dims=3, bottleneck_attention=False,
                resblock_updown=True, use_freq=True)
    WavUNetModel(**base)
    ssn = WavUNetModel(**dict(base, use_scale_shift_norm=True))       # served (inference): emb_layers produce (scale, shift)
    assert ssn.input_blocks[1][0].emb_layers[1].out_features == 2 * ssn.input_blocks[1][0].out_channels
    for bad in (dict(dims=2), dict(use_freq=False), dict(resblock_updown=False), dict(additive_skips=True),
                dict(bottleneck_attention=True), dict(attention_resolutions=(2,))):
        with pytest.raises(NotImplementedError):
            WavUNetModel(**{**base, **bad})


def test_cfg_w4_state_dict_matches_reference(golden):
    from guided_diffusion.wunet import WavUNetModel
    g = golden("cfg_w4_keys")
    m = WavUNetModel(image_size=224, in_channels=32, model_channels=64, out_channels=8, num_res_blocks=2,
                     attention_resolutions=(), channel_mult=(1, 2, 2, 4), dims=3, num_groups=32,
                     bottleneck_attention=False, resblock_updown=True, use_freq=True)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    assert len(sd) == 346
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"]) == 54285640
    assert len({v.data_ptr() for v in sd.values()}) == 306          # weight-tied output blocks (wunet.py:647-673)
    assert m.output_blocks[1][0] is m.output_blocks[2][0]
    # zero-initialised second conv of every ResBlock (wunet.py:213); `out` conv is NOT zeroed (:701-705)
    assert float(m.input_blocks[1][0].out_layers[3].weight.detach().abs().max()) == 0.0
    assert float(m.out[2].weight.detach().abs().max()) > 0.0
    m2 = WavUNetModel(image_size=224, in_channels=32, model_channels=64, out_channels=8, num_res_blocks=2,
                      attention_resolutions=(), channel_mult=(1, 2, 2, 4), dims=3, num_groups=32,
                      bottleneck_attention=False, resblock_updown=True, use_freq=True)
    m2.load_state_dict(sd, strict=True)


def test_schedules_and_respacing_match_reference(golden):
    from guided_diffusion import gaussian_diffusion as gd
    from guided_diffusion.respace import space_timesteps
    from guided_diffusion.script_util import create_gaussian_diffusion
    g = golden("schedules")
    np.testing.assert_array_equal(gd.get_named_beta_schedule("linear", 10, "sampled"), g["sampled10"])
    np.testing.assert_array_equal(gd.get_named_beta_schedule("linear", 20, "direct"), g["direct20"])
    np.testing.assert_array_equal(gd.get_named_beta_schedule("linear", 1000, "direct"), g["direct1000"])
    np.testing.assert_allclose(gd.get_named_beta_schedule("cosine", 50), g["cosine50"], rtol=1e-15)
    assert sorted(space_timesteps(1000, "100")) == list(g["space_1000_100"])
    assert sorted(space_timesteps(1000, "ddim50")) == list(g["space_1000_ddim50"])
    assert sorted(space_timesteps(300, [10, 15, 20])) == list(g["space_300_10_15_20"])
    assert sorted(space_timesteps(1000, "10,10,10")) == list(g["space_1000_10_10_10"])
    with pytest.raises(ValueError):
        space_timesteps(10, "20")
    d100 = create_gaussian_diffusion(steps=1000, predict_xstart=True, timestep_respacing="100", mode="i2i")
    np.testing.assert_array_equal(d100.betas, g["respaced100_betas"])
    assert d100.timestep_map == list(g["respaced100_map"]) and d100.original_num_steps == 1000
    d10 = create_gaussian_diffusion(steps=10, predict_xstart=True, sample_schedule="sampled", mode="i2i")
    for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                 "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                 "posterior_mean_coef2"):
        np.testing.assert_array_equal(getattr(d10, name), g["d10_" + name], err_msg=name)
    assert d10.timestep_map == list(g["d10_map"]) and d10.mode == "i2i"
    assert d10.model_mean_type == gd.ModelMeanType.START_X and d10.model_var_type == gd.ModelVarType.FIXED_LARGE
    with pytest.raises(AssertionError):            # 'direct' with T=10 gives beta_end = 2.0 (reference :164)
        create_gaussian_diffusion(steps=10, sample_schedule="direct")


def test_flag_plumbing_matches_reference_defaults():
    import argparse
    from guided_diffusion.script_util import (add_dict_to_argparser, args_to_dict, model_and_diffusion_defaults,
                                              str2bool)
    d = model_and_diffusion_defaults()
    assert d["use_freq"] is False and d["sample_schedule"] == "direct" and d["diffusion_steps"] == 1000
    # dict printed by the reference's model_and_diffusion_defaults() in this container (script_util.py:70-104)
    ref = {'image_size': 64, 'num_channels': 128, 'num_res_blocks': 2, 'num_heads': 4, 'num_heads_upsample': -1,
           'num_head_channels': -1, 'attention_resolutions': '16,8', 'channel_mult': '', 'dropout': 0.0,
           'class_cond': False, 'use_checkpoint': False, 'use_scale_shift_norm': True, 'resblock_updown': True,
           'use_fp16': False, 'use_new_attention_order': False, 'dims': 2, 'num_groups': 32, 'in_channels': 1,
           'out_channels': 0, 'bottleneck_attention': True, 'resample_2d': True, 'additive_skips': False,
           'mode': 'default', 'use_freq': False, 'predict_xstart': False, 'sample_schedule': 'direct',
           'learn_sigma': False, 'diffusion_steps': 1000, 'noise_schedule': 'linear', 'timestep_respacing': '',
           'use_kl': False, 'rescale_timesteps': False, 'rescale_learned_sigmas': False, 'dataset': 'brats'}
    assert d == ref and list(d) == list(ref)
    p = argparse.ArgumentParser()
    add_dict_to_argparser(p, d)
    a = p.parse_args(["--use_freq=True", "--channel_mult=1,2,2,4", "--diffusion_steps=10"])
    kw = args_to_dict(a, d.keys())
    assert kw["use_freq"] is True and kw["channel_mult"] == "1,2,2,4" and kw["diffusion_steps"] == 10
    assert str2bool("yes") and not str2bool("0")


def test_unet_state_dict_matches_reference(golden):
    """The drop-in plain UNetModel exposes the reference's state-dict keys and shapes (run.sh configuration:
    channel_mult 1,2,2,4,4, 64 base channels), so released checkpoints load unchanged."""
    from guided_diffusion.script_util import create_model, model_and_diffusion_defaults
    g = golden("unet_small")
    m = create_model(image_size=224, num_channels=64, num_res_blocks=2, channel_mult="1,2,2,4,4", attention_resolutions="",
                     dims=3, in_channels=32, out_channels=8, bottleneck_attention=False, resblock_updown=True,
                     resample_2d=False, use_freq=False)
    assert type(m).__module__ == "guided_diffusion.unet" and type(m).__name__ == "UNetModel"
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["big_keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in g["big_shapes"]]
    assert sum(p.numel() for p in m.parameters()) == int(g["big_n_params"])
    with pytest.raises(NotImplementedError):
        create_model(image_size=224, num_channels=64, num_res_blocks=2, channel_mult="1,2", attention_resolutions="",
                     dims=3, in_channels=32, out_channels=8, bottleneck_attention=False, resblock_updown=False,
                     use_freq=False)
