"""Drop-in for the reference's ``guided_diffusion.wunet``: the wavelet U-Net denoiser.

``WavUNetModel`` keeps the reference constructor signature (guided_diffusion/wunet.py:435-440), builds the
same module tree -- hence the same ``state_dict`` keys (346 in CFG-W4, including the weight-tied ResBlocks that
the reference creates by re-using its ``layers`` list at :647-673) and the same default initialisation -- and
keeps ``forward(x, timesteps)``, ``.to()``, ``.eval()``, ``load_state_dict`` semantics.  The sub-modules are
parameter containers: the forward pass is not executed module by module but by ``fcwdm.engine.WavUNetEngine``
as a sequence of hand-written sm_100a kernels on channels-last bf16 activations (tcgen05 implicit-GEMM conv3d,
fused GroupNorm+SiLU, one-pass Haar DWT/IDWT with the /3, x3 and timestep-embedding adds folded in).

Supported flag set = the only one under which the reference model itself runs:
dims=3, use_freq=True, resblock_updown=True, additive_skips=False, no attention, progressive_input='residual';
use_scale_shift_norm either way (run.sh ships False).  Anything else raises NotImplementedError at construction instead of failing
deep inside forward as the reference does.
"""
from abc import abstractmethod

import torch as th
import torch.nn as nn

from .nn import conv_nd, linear, normalization, zero_module
from DWT_IDWT.DWT_IDWT_layer import DWT_3D, IDWT_3D

_FUSED_MSG = "this block runs inside WavUNetModel's fused fcwdm plan; call WavUNetModel.forward"


class TimestepBlock(nn.Module):
    """Any module whose forward() takes timestep embeddings as a second argument."""

    @abstractmethod
    def forward(self, x, emb):
        """Apply the module to `x` given `emb` timestep embeddings."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Sequential container that routes the timestep embedding to the children that take it."""

    def forward(self, x, emb):
        raise NotImplementedError(_FUSED_MSG)


class Upsample(nn.Module):
    """Wavelet up-sampling: x = IDWT(3*x, *skip7) (reference wunet.py:40-85)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, resample_2d=True, use_freq=True):
        super().__init__()
        if use_conv:
            raise NotImplementedError("Upsample(use_conv=True) (grouped conv on the skip bands) is never built by "
                                      "WavUNetModel (wunet.py:194-195 pass False)")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.resample_2d = resample_2d
        self.use_freq = use_freq
        self.idwt = IDWT_3D("haar")

    def forward(self, x):
        raise NotImplementedError(_FUSED_MSG)


class Downsample(nn.Module):
    """Wavelet down-sampling: (LLL/3, 7 high bands) = DWT(x) (reference wunet.py:88-124)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, resample_2d=True, use_freq=True):
        super().__init__()
        if use_conv or not use_freq:
            raise NotImplementedError("only the wavelet Downsample (use_conv=False, use_freq=True) is supported")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.use_freq = use_freq
        self.dwt = DWT_3D("haar")
        self.op = self.dwt

    def forward(self, x):
        raise NotImplementedError(_FUSED_MSG)


class WaveletDownsample(nn.Module):
    """Input-pyramid block: conv(cat(DWT(x)) / 3) (reference wunet.py:127-145)."""

    def __init__(self, in_ch=None, out_ch=None):
        super().__init__()
        out_ch = out_ch if out_ch else in_ch
        self.in_ch = in_ch
        self.out_ch = out_ch
        self.conv = conv_nd(3, self.in_ch * 8, self.out_ch, 3, stride=1, padding=1)
        self.dwt = DWT_3D('haar')

    def forward(self, x):
        raise NotImplementedError(_FUSED_MSG)


class ResBlock(TimestepBlock):
    """Residual block with optional wavelet up/down-sampling (reference wunet.py:148-269)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=True, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False, num_groups=32, resample_2d=True, use_freq=False):
        super().__init__()
        if dims != 3:
            raise NotImplementedError("only dims=3 is implemented")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_scale_shift_norm = use_scale_shift_norm
        self.use_checkpoint = use_checkpoint
        self.up = up
        self.down = down
        self.num_groups = num_groups
        self.use_freq = use_freq

        self.in_layers = nn.Sequential(
            normalization(channels, self.num_groups),
            nn.SiLU(),
            conv_nd(dims, channels, self.out_channels, 3, padding=1),
        )
        self.updown = up or down
        if up:
            self.h_upd = Upsample(channels, False, dims, resample_2d=resample_2d, use_freq=self.use_freq)
            self.x_upd = Upsample(channels, False, dims, resample_2d=resample_2d, use_freq=self.use_freq)
        elif down:
            self.h_upd = Downsample(channels, False, dims, resample_2d=resample_2d, use_freq=self.use_freq)
            self.x_upd = Downsample(channels, False, dims, resample_2d=resample_2d, use_freq=self.use_freq)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        # use_scale_shift_norm: emb_out = (scale, shift), h = out_norm(h) * (1 + scale) + shift (reference wunet.py:256-260)
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, (2 if use_scale_shift_norm else 1) * self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels, self.num_groups),
            nn.SiLU(),
            nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)),
        )
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward(self, x, temb):
        raise NotImplementedError(_FUSED_MSG)


class WavUNetModel(nn.Module):
    """The wavelet U-Net with timestep embedding (reference wunet.py:410-795)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False, num_groups=32,
                 bottleneck_attention=True, resample_2d=True, additive_skips=False, decoder_device_thresh=0,
                 use_freq=False, progressive_input='residual'):
        super().__init__()
        unsupported = []
        if dims != 3:
            unsupported.append("dims != 3")
        if not use_freq:
            unsupported.append("use_freq=False (that is UNetModel's job, script_util.py:243-267)")
        if not resblock_updown:
            unsupported.append("resblock_updown=False (crashes in the reference: wunet.py:120 unpacks a conv output)")
        if additive_skips:
            unsupported.append("additive_skips=True (crashes in the reference: wunet.py:773-774)")
        if tuple(attention_resolutions) or bottleneck_attention:
            unsupported.append("attention (AttentionBlock receives a tuple in the reference: wunet.py:314,763)")
        if num_classes is not None:
            unsupported.append("class conditioning")
        if progressive_input != 'residual':
            unsupported.append("progressive_input != 'residual'")
        if unsupported:
            raise NotImplementedError("WavUNetModel (fcwdm B200 path) does not support: " + "; ".join(unsupported))

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = tuple(channel_mult)
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_groups = num_groups
        self.bottleneck_attention = bottleneck_attention
        self.devices = None
        self.decoder_device_thresh = decoder_device_thresh
        self.additive_skips = additive_skips
        self.use_freq = use_freq
        self.progressive_input = progressive_input

        emb_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, emb_dim), nn.SiLU(), linear(emb_dim, emb_dim))

        def block(cin, cout=None, **kw):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, num_groups=num_groups,
                            resample_2d=resample_2d, use_freq=use_freq, **kw)

        # ---- encoder: stem conv, then per level [ResBlock x num_res_blocks, ResBlock(down), WaveletDownsample]
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        self._feature_size = model_channels
        chans = [model_channels]
        ch = model_channels
        pyramid_ch = in_channels
        for mult in self.channel_mult:
            width = mult * model_channels
            for _ in range(num_res_blocks):
                self.input_blocks.append(TimestepEmbedSequential(block(ch, width)))
                ch = width
                self._feature_size += ch
                chans.append(ch)
            self.input_blocks.append(TimestepEmbedSequential(block(ch, ch, down=True)))
            self.input_blocks.append(TimestepEmbedSequential(WaveletDownsample(in_ch=pyramid_ch, out_ch=ch)))
            pyramid_ch = ch
            chans.append(ch)
            self._feature_size += ch
        self.input_block_chans_bk = chans[:]

        # ---- bottleneck
        self.middle_block = TimestepEmbedSequential(block(ch), block(ch))
        self._feature_size += ch

        # ---- decoder: per level [ResBlock] x num_res_blocks, then [the last of those AGAIN, ResBlock(up)]:
        # the reference appends the up block to the list that still holds the previous ResBlock, so that
        # block is shared (same Parameters) between two consecutive output_blocks entries.
        self.output_blocks = nn.ModuleList([])
        for mult in reversed(self.channel_mult):
            width = mult * model_channels
            last = None
            for _ in range(num_res_blocks):
                last = block(ch, width)
                self.output_blocks.append(TimestepEmbedSequential(last))
                ch = width
                self._feature_size += ch
            if last is None:
                self.output_blocks.append(TimestepEmbedSequential(block(ch, ch, up=True)))
            else:
                self.output_blocks.append(TimestepEmbedSequential(last, block(ch, ch, up=True)))
            self._feature_size += ch

        self.out_res = nn.ModuleList([TimestepEmbedSequential(block(ch, ch)) for _ in range(num_res_blocks)])
        self.out = nn.Sequential(normalization(ch, num_groups), nn.SiLU(),
                                 conv_nd(dims, model_channels, out_channels, 3, padding=1))
        self._engine = None
        self._train_engine = None

    # the reference overrides .to() to support a (broken for WavUNet) 2-device split and to return None
    # (wunet.py:707-732).  Here: a 1-element list/tuple is unwrapped, >1 devices is refused, and the module is
    # returned as nn.Module.to does (callers that ignore the return value are unaffected).
    def to(self, *args, **kwargs):
        if args and isinstance(args[0], (list, tuple)):
            if len(args[0]) > 1:
                raise NotImplementedError("splitting WavUNetModel across two devices is not supported (the reference's "
                                          "own split never moves activations for this model, SURVEY.md 2.1)")
            args = (args[0][0],) + tuple(args[1:])
        out = super().to(*args, **kwargs)
        p = next(self.parameters())
        self.devices = [p.device, p.device]
        return out

    def engine(self):
        if self._engine is None:
            from fcwdm.engine import WavUNetEngine
            object.__setattr__(self, "_engine", WavUNetEngine(self))
        return self._engine

    def train_engine(self):
        if getattr(self, "_train_engine", None) is None:
            from fcwdm.train_engine import WavUNetTrainEngine
            object.__setattr__(self, "_train_engine", WavUNetTrainEngine(self))
        return self._train_engine

    def forward(self, x, timesteps):
        """x: [N, C, D, H, W] fp32, timesteps: [N] -> [N, out_channels, D, H, W].

        Under ``torch.no_grad()`` (sampling) this is the fused inference plan.  With autograd enabled and trainable
        parameters (scripts/train.py) the forward records a tape and the returned tensor carries ONE autograd node whose
        backward runs the explicit fcwdm backward kernels (fcwdm/train_engine.py) and hands every parameter its fp32
        gradient; the gradient w.r.t. ``x`` is not produced (training_losses never asks for it)."""
        self.hs_shapes = []
        if th.is_grad_enabled():
            params = [p for p in self.parameters()]
            if any(p.requires_grad for p in params):
                if not all(p.requires_grad for p in params):
                    raise NotImplementedError("partially frozen WavUNetModel: the fcwdm backward produces gradients for "
                                              "all parameters or none")
                from fcwdm.train_engine import WavUNetFunction
                return WavUNetFunction.apply(self.train_engine(), x, timesteps, *params)
            if x.requires_grad:
                raise NotImplementedError("gradient w.r.t. the denoiser input is not implemented by the fcwdm backward")
        return self.engine().forward(x, timesteps)
