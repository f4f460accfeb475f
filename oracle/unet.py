"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32 functional ops) of the reference's plain UNetModel forward
(guided_diffusion/unet.py), for the flag set run.sh ships: dims=3, resblock_updown=True, no attention,
use_scale_shift_norm=False, additive_skips=False.  Weights come from a plain state dict with the reference's keys.
"""
import torch
import torch.nn.functional as F

from .wunet import _conv, _emb_and_out_layers, _gn_silu, timestep_embedding


def _resample(x, up, resample_2d):
    if up:                                                                              # unet.py:60-70
        if resample_2d:
            return F.interpolate(x, (x.shape[2], x.shape[3] * 2, x.shape[4] * 2), mode="nearest")
        return F.interpolate(x, scale_factor=2, mode="nearest")
    k = (1, 2, 2) if resample_2d else 2                                                 # unet.py:89-96
    return F.avg_pool3d(x, kernel_size=k, stride=k)


def _resblock(sd, p, x, emb, groups, up=False, down=False, resample_2d=False):
    """ResBlock._forward (unet.py:285-311)."""
    if up or down:
        h = _gn_silu(sd, p + "in_layers.0", x, groups)                                  # in_rest        :286-288
        h = _resample(h, up, resample_2d)                                               # h_upd          :289
        x = _resample(x, up, resample_2d)                                               # x_upd          :290
        h = _conv(sd, p + "in_layers.2", h, 1)                                          # in_conv        :291
    else:
        h = _conv(sd, p + "in_layers.2", _gn_silu(sd, p + "in_layers.0", x, groups), 1)  # :293
    emb_out = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])   # :295
    h = _emb_and_out_layers(sd, p, h, emb_out, groups)                                  # :297-309
    if (p + "skip_connection.weight") in sd:
        w = sd[p + "skip_connection.weight"]
        x = F.conv3d(x, w, sd[p + "skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h                                                                        # :311


def unet_forward(sd, x, timesteps, *, model_channels, channel_mult, num_res_blocks=2, num_groups=32, resample_2d=False):
    """UNetModel.forward (unet.py:754-800)."""
    sd = {k: (v if v.dtype == torch.float32 else v.float()) for k, v in sd.items()}
    emb = timestep_embedding(timesteps, model_channels)
    emb = F.linear(F.silu(F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])                  # :777
    hs = []
    h = _conv(sd, "input_blocks.0.0", x.float(), 1)
    hs.append(h)
    idx = 1
    levels = len(channel_mult)
    for level in range(levels):                                                         # :560-620 construction order
        for _ in range(num_res_blocks):
            h = _resblock(sd, f"input_blocks.{idx}.0.", h, emb, num_groups)
            hs.append(h)
            idx += 1
        if level != levels - 1:
            h = _resblock(sd, f"input_blocks.{idx}.0.", h, emb, num_groups, down=True, resample_2d=resample_2d)
            hs.append(h)
            idx += 1
    h = _resblock(sd, "middle_block.0.", h, emb, num_groups)                            # :790
    h = _resblock(sd, "middle_block.1.", h, emb, num_groups)
    k = 0
    for level in reversed(range(levels)):                                               # :662-716, :792-798
        for i in range(num_res_blocks + 1):
            h = torch.cat([h, hs.pop()], dim=1)                                         # :796
            h = _resblock(sd, f"output_blocks.{k}.0.", h, emb, num_groups)
            if level and i == num_res_blocks:
                h = _resblock(sd, f"output_blocks.{k}.1.", h, emb, num_groups, up=True, resample_2d=resample_2d)
            k += 1
    return _conv(sd, "out.2", _gn_silu(sd, "out.0", h, num_groups), 1)                  # :800
