"""TEST INFRASTRUCTURE -- freezes ``inspect.signature`` of every symbol of the drop-in boundary
as the UNMODIFIED reference defines it, into tests/golden/signatures.json.

    python oracle/make_golden_signatures.py        # needs /root/reference (build container only)

``tests/test_signatures_cpu.py`` diffs the drop-in's signatures against this file (and, when the reference is
mounted, this file against the live reference), failing on anything that is not whitelisted there."""
import importlib
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

# module -> dotted symbols whose call signature is part of the boundary
BOUNDARY = {
    "DWT_IDWT.DWT_IDWT_layer": ["DWT_3D.__init__", "DWT_3D.forward", "IDWT_3D.__init__", "IDWT_3D.forward"],
    "guided_diffusion.wunet": ["WavUNetModel.__init__", "WavUNetModel.forward", "ResBlock.__init__", "Upsample.__init__",
                               "Downsample.__init__", "WaveletDownsample.__init__"],
    "guided_diffusion.unet": ["UNetModel.__init__", "UNetModel.forward"],
    "guided_diffusion.gaussian_diffusion": [
        "get_named_beta_schedule", "betas_for_alpha_bar", "_extract_into_tensor", "GaussianDiffusion.__init__",
        "GaussianDiffusion.q_mean_variance", "GaussianDiffusion.q_sample", "GaussianDiffusion.q_posterior_mean_variance",
        "GaussianDiffusion.p_mean_variance", "GaussianDiffusion.p_sample", "GaussianDiffusion.p_sample_loop",
        "GaussianDiffusion.p_sample_loop_progressive", "GaussianDiffusion.training_losses",
        "GaussianDiffusion._predict_xstart_from_eps", "GaussianDiffusion._scale_timesteps"],
    "guided_diffusion.respace": ["space_timesteps", "SpacedDiffusion.__init__", "SpacedDiffusion.p_mean_variance",
                                 "SpacedDiffusion.training_losses", "_WrappedModel.__init__", "_WrappedModel.__call__"],
    "guided_diffusion.script_util": ["diffusion_defaults", "model_and_diffusion_defaults", "create_model_and_diffusion",
                                     "create_model", "create_gaussian_diffusion", "add_dict_to_argparser", "args_to_dict",
                                     "str2bool"],
    "guided_diffusion.nn": ["GroupNorm32.forward", "conv_nd", "linear", "normalization", "zero_module", "mean_flat",
                            "timestep_embedding"],
    "guided_diffusion.train_util": ["TrainLoop.__init__", "TrainLoop.run_loop", "TrainLoop.run_step",
                                    "TrainLoop.forward_backward", "TrainLoop.save_if_best", "TrainLoop.save",
                                    "parse_resume_step_from_filename", "get_blob_logdir", "find_resume_checkpoint"],
    "guided_diffusion.dist_util": ["setup_dist", "dev", "load_state_dict", "sync_params"],
    "guided_diffusion.resample": ["create_named_schedule_sampler", "UniformSampler.__init__", "ScheduleSampler.sample"],
    "guided_diffusion.bratsloader": ["BRATSVolumes.__init__", "BRATSVolumes.__getitem__", "clip_and_normalize"],
}


def signatures(modules=BOUNDARY):
    out = {}
    for mod_name, symbols in modules.items():
        try:
            mod = importlib.import_module(mod_name)
        except Exception as exc:                      # recorded, so a missing module is a visible diff
            out[mod_name] = f"<import failed: {type(exc).__name__}: {exc}>"
            continue
        for dotted in symbols:
            obj = mod
            try:
                for part in dotted.split("."):
                    obj = getattr(obj, part)
                out[f"{mod_name}:{dotted}"] = str(inspect.signature(obj))
            except Exception as exc:
                out[f"{mod_name}:{dotted}"] = f"<{type(exc).__name__}>"
    return out


def reference_signatures():
    from oracle import ref_shims
    with ref_shims.reference_modules():
        return signatures()


if __name__ == "__main__":
    sig = reference_signatures()
    path = os.path.join(ROOT, "tests", "golden", "signatures.json")
    with open(path, "w") as fh:
        json.dump(sig, fh, indent=1, sort_keys=True)
    print(f"wrote {len(sig)} signatures to {path}")
