"""TEST INFRASTRUCTURE ONLY -- functional fp32 CPU restatement of WavUNetModel.forward.

Walks a reference-format ``state_dict`` (the 346 key names of guided_diffusion/wunet.py) with
torch.nn.functional ops, for the only flag set under which the reference model runs:
use_freq=True, resblock_updown=True, additive_skips=False, no attention, dims=3,
use_scale_shift_norm=False, dropout=0, progressive_input='residual'.

Citations are into /root/reference/guided_diffusion/wunet.py unless another file is named.
"""
import math

import torch
import torch.nn.functional as F

from . import haar


_S = 0.7071067811865476   # cast to fp32 by the tensor arithmetic below, as torch.Tensor(matrix) does (layer.py:506-518)


def _analysis(x, dim):
    """Same operations, in the same order, as oracle.haar._analysis, but in torch so that autograd differentiates
    through the wavelet up/down-sampling (the reference's backward is DWT_IDWT_Functions.py:139-156)."""
    idx0 = [slice(None)] * x.dim()
    idx1 = [slice(None)] * x.dim()
    idx0[dim], idx1[dim] = slice(0, None, 2), slice(1, None, 2)
    a, b = x[tuple(idx0)], x[tuple(idx1)]
    return _S * a + _S * b, _S * a - _S * b


def _synthesis(lo, hi, dim):
    even, odd = _S * lo + _S * hi, _S * lo - _S * hi
    out = torch.stack([even, odd], dim=dim + 1)
    shape = list(lo.shape)
    shape[dim] *= 2
    return out.reshape(shape)


def _dwt(x):
    """oracle.haar.dwt3d in torch (H, then W, then D; DWT_IDWT_Functions.py:122-135)."""
    L, H = _analysis(x, 3)
    LL, LH = _analysis(L, 4)
    HL, HH = _analysis(H, 4)
    LLL, HLL = _analysis(LL, 2)
    LLH, HLH = _analysis(LH, 2)
    LHL, HHL = _analysis(HL, 2)
    LHH, HHH = _analysis(HH, 2)
    return LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH


def _idwt(LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH):
    """oracle.haar.idwt3d in torch (D, then W, then H; DWT_IDWT_Functions.py:167-180)."""
    LL = _synthesis(LLL, HLL, 2)
    LH = _synthesis(LLH, HLH, 2)
    HL = _synthesis(LHL, HHL, 2)
    HH = _synthesis(LHH, HHH, 2)
    L = _synthesis(LL, LH, 4)
    H = _synthesis(HL, HH, 4)
    return _synthesis(L, H, 3)


def timestep_embedding(timesteps, dim, max_period=10000):
    """nn.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn_silu(sd, p, x, groups):
    """GroupNorm32 (nn.py:17-19, eps 1e-5 default) + SiLU."""
    return F.silu(F.group_norm(x.float(), groups, sd[p + ".weight"], sd[p + ".bias"], 1e-5))


def _conv(sd, p, x, pad):
    return F.conv3d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def _emb_and_out_layers(sd, p, h, emb_out, groups):
    """The tail both U-Nets' ResBlocks share (wunet.py:252-263, unet.py:297-309): h + emb_out then out_layers, or --
    use_scale_shift_norm, recognised by emb_layers producing 2 x C values -- out_norm(h) * (1 + scale) + shift with
    (scale, shift) = chunk(emb_out, 2), then SiLU and the conv."""
    C = sd[p + "out_layers.0.weight"].shape[0]
    if emb_out.shape[1] == 2 * C:
        scale = emb_out[:, :C, None, None, None]
        shift = emb_out[:, C:, None, None, None]
        hn = F.group_norm(h.float(), groups, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], 1e-5)
        return _conv(sd, p + "out_layers.3", F.silu(hn * (1 + scale) + shift), 1)
    h = h + emb_out[:, :, None, None, None]
    return _conv(sd, p + "out_layers.3", _gn_silu(sd, p + "out_layers.0", h, groups), 1)


def _resblock(sd, p, x, skip, emb, groups, up=False, down=False):
    """ResBlock.forward (:223-269).  ``x`` tensor, ``skip`` = 7-tuple of bands or None.
    Returns (out, hSkip)."""
    hskip = skip
    h = _conv(sd, p + "in_layers.2", _gn_silu(sd, p + "in_layers.0", x, groups), 1)        # :234/:247
    if down:                                                                               # :240-241, :118-121
        hb = _dwt(h)
        h, hskip = hb[0] / 3.0, tuple(hb[1:])
        x = _dwt(x)[0] / 3.0
    elif up:                                                                               # :236-241, :65-85
        h = _idwt(3.0 * h, *skip)
        x = _idwt(3.0 * x, *skip)
        hskip = None
    emb_out = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])   # :250
    h = _emb_and_out_layers(sd, p, h, emb_out, groups)                                     # :252-263 (dropout p=0)
    if (p + "skip_connection.weight") in sd:                                               # :217-220
        x = _conv(sd, p + "skip_connection", x, 0)
    return x + h, hskip                                                                    # :266-267


def wunet_forward(sd, x, timesteps, *, model_channels, channel_mult, num_res_blocks=2, num_groups=32):
    """WavUNetModel.forward (:734-795) for the module list built by __init__ (:480-705)."""
    sd = {k: (v if v.dtype == torch.float32 else v.float()) for k, v in sd.items()}   # keeps autograd leaves (oracle.train)
    L = len(channel_mult)
    emb = timestep_embedding(timesteps, model_channels)                                    # :745
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    hs = []
    pyramid = x
    h = _conv(sd, "input_blocks.0.0", x, 1)                                                # input_blocks[0] (:482-484)
    hs.append(None)
    idx = 1
    for _level in range(L):                                                                # :497-570
        for _ in range(num_res_blocks):
            h, _s = _resblock(sd, f"input_blocks.{idx}.0.", h, None, emb, num_groups)
            hs.append(None)                                                                # :752-755
            idx += 1
        h, s = _resblock(sd, f"input_blocks.{idx}.0.", h, None, emb, num_groups, down=True)
        hs.append(s)
        idx += 1
        bands = _dwt(pyramid)                                                              # WaveletDownsample :142-145
        pyramid = _conv(sd, f"input_blocks.{idx}.0.conv", torch.cat(bands, dim=1) / 3.0, 1)
        pyramid = pyramid + h                                                              # :758-760
        h = pyramid
        idx += 1

    skip = None
    for j in range(2):                                                                     # middle_block (:577-609, :762-765)
        h, skip = _resblock(sd, f"middle_block.{j}.", h, None, emb, num_groups)            # skip becomes None (:765)

    k = 0
    for _level in range(L):                                                                # output_blocks (:615-675, :767-789)
        for i in range(num_res_blocks + 1):
            new_hs = hs.pop()                                                              # :768-770
            if new_hs:
                skip = new_hs
            if i < num_res_blocks:
                h, skip_out = _resblock(sd, f"output_blocks.{k}.0.", h, skip, emb, num_groups)
            else:
                # `layers` is reused at :647-673, so block k = [the previous ResBlock again, ResBlock(up)]
                h, skip_out = _resblock(sd, f"output_blocks.{k}.0.", h, skip, emb, num_groups)
                h, skip_out = _resblock(sd, f"output_blocks.{k}.1.", h, skip_out, emb, num_groups, up=True)
            # forward() re-inserts `skip` into the tuple before the next module (:778-783); a standard
            # ResBlock passes its input skip through (:225-228,:267) and an up block returns None.
            k += 1

    for i in range(num_res_blocks):                                                        # out_res (:680-696, :791-792)
        h, _s = _resblock(sd, f"out_res.{i}.0.", h, None, emb, num_groups)
    return _conv(sd, "out.2", _gn_silu(sd, "out.0", h, num_groups), 1)                     # :794-795, :701-705


def seeded_state_dict(shapes, seed=0, std=0.05):
    """Deterministic weights independent of module construction order: every tensor is drawn from its own
    generator seeded by (seed, crc32(key)).  GroupNorm scales are centred on 1.  This also re-randomises
    the zero-initialised convs (wunet.py:213), without which every ResBlock branch is multiplied by 0
    (SURVEY.md fact 5).  Tied keys (output_blocks.{3l+1}.0 == output_blocks.{3l+2}.0) must be given the
    same tensor by the caller; see :func:`tie_output_blocks`."""
    import zlib
    out = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        w = torch.randn(tuple(shape), generator=g) * std
        is_norm_scale = key.endswith(".weight") and len(shape) == 1
        out[key] = (1.0 + w) if is_norm_scale else w
    return out


def tie_output_blocks(sd, num_levels, num_res_blocks=2):
    """Mirror the construction quirk at wunet.py:647-673 on a plain dict."""
    for l in range(num_levels):
        a = l * (num_res_blocks + 1) + num_res_blocks - 1
        b = a + 1
        for key in list(sd):
            if key.startswith(f"output_blocks.{a}.0."):
                sd[f"output_blocks.{b}.0." + key[len(f"output_blocks.{a}.0."):]] = sd[key]
    return sd
