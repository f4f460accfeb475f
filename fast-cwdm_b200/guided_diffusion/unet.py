"""Drop-in for the reference's ``guided_diffusion.unet.UNetModel``: the plain (non-wavelet-resampling) 3-D U-Net that
``run.sh`` instantiates with ``--use_freq=False`` (run.sh:59-66,109-133) and that fast-cwdm's released checkpoints
(``brats_<mod>_BEST_sampled_10.pt``) load into -- SURVEY.md section 8f, row 1.

``UNetModel`` keeps the reference constructor signature (guided_diffusion/unet.py:482-507), builds the same module
tree -- hence the same ``state_dict`` keys and default initialisation -- and keeps ``forward(x, timesteps)``.  As for
``WavUNetModel`` the sub-modules are parameter containers: the forward pass is a fixed sequence of sm_100a kernels run
by ``fcwdm.unet_engine.UNetEngine`` on channels-last bf16 activations: tcgen05 implicit-GEMM conv3d (bias, timestep
embedding and residual in the epilogue), fused GroupNorm+SiLU, average-pool / nearest-neighbour resampling, and
skip concatenation (unet.py:796) without a copy -- producers write straight into column slices of the concat buffer.

Supported flag set = what run.sh ships: dims=3, resblock_updown=True, no attention (attention_resolutions="",
bottleneck_attention=False), additive_skips=False, class_cond=False; resample_2d either way; use_scale_shift_norm
either way for sampling (False for training).  Anything else raises NotImplementedError at construction.  Sampling and training (autograd) are both served.
"""
import torch as th
import torch.nn as nn

from .nn import conv_nd, linear, normalization, zero_module
from .wunet import TimestepBlock, TimestepEmbedSequential   # same helper classes as the reference's unet.py:13-37

_FUSED_MSG = "this block runs inside UNetModel's fused fcwdm plan; call UNetModel.forward"


class Upsample(nn.Module):
    """Nearest-neighbour x2 (reference unet.py:40-70); use_conv=False is the only form ResBlock builds."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, resample_2d=True):
        super().__init__()
        if use_conv:
            raise NotImplementedError("Upsample(use_conv=True) is only built with resblock_updown=False")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.resample_2d = resample_2d

    def forward(self, x):
        raise NotImplementedError(_FUSED_MSG)


class Downsample(nn.Module):
    """Average pool, kernel = stride = 2 or (1,2,2) (reference unet.py:73-100)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, resample_2d=True):
        super().__init__()
        if use_conv:
            raise NotImplementedError("Downsample(use_conv=True) (strided conv) is only built with resblock_updown=False")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        self.resample_2d = resample_2d
        self.op = nn.Identity()        # parameter-free placeholder for avg_pool_nd (keeps the module tree shape)

    def forward(self, x):
        raise NotImplementedError(_FUSED_MSG)


class ResBlock(TimestepBlock):
    """Residual block with optional average-pool / nearest up-sampling (reference unet.py:185-311)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False, num_groups=32, resample_2d=True):
        super().__init__()
        if dims != 3:
            raise NotImplementedError("only dims=3 is implemented")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.num_groups = num_groups
        self.up = up
        self.down = down
        self.resample_2d = resample_2d

        self.in_layers = nn.Sequential(
            normalization(channels, self.num_groups),
            nn.SiLU(),
            conv_nd(dims, channels, self.out_channels, 3, padding=1),
        )
        self.updown = up or down
        if up:
            self.h_upd = Upsample(channels, False, dims, resample_2d=resample_2d)
            self.x_upd = Upsample(channels, False, dims, resample_2d=resample_2d)
        elif down:
            self.h_upd = Downsample(channels, False, dims, resample_2d=resample_2d)
            self.x_upd = Downsample(channels, False, dims, resample_2d=resample_2d)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        # use_scale_shift_norm: emb_out = (scale, shift), h = out_norm(h) * (1 + scale) + shift (reference unet.py:301-305)
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, (2 if use_scale_shift_norm else 1) * self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels, self.num_groups),
            nn.SiLU(),
            nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)),
        )
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward(self, x, emb):
        raise NotImplementedError(_FUSED_MSG)


class UNetModel(nn.Module):
    """The full U-Net with timestep embedding (reference unet.py:451-800)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False, num_groups=32,
                 bottleneck_attention=True, resample_2d=True, additive_skips=False, decoder_device_thresh=0):
        super().__init__()
        unsupported = []
        if dims != 3:
            unsupported.append("dims != 3")
        if not resblock_updown:
            unsupported.append("resblock_updown=False (strided-conv Downsample / conv Upsample layers)")
        if additive_skips:
            unsupported.append("additive_skips=True")
        if tuple(attention_resolutions) or bottleneck_attention:
            unsupported.append("attention blocks")
        if num_classes is not None:
            unsupported.append("class conditioning")
        if unsupported:
            raise NotImplementedError("UNetModel (fcwdm B200 path) does not support: " + "; ".join(unsupported))

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = tuple(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads if num_heads_upsample == -1 else num_heads_upsample
        self.num_groups = num_groups
        self.bottleneck_attention = bottleneck_attention
        self.devices = None
        self.decoder_device_thresh = decoder_device_thresh
        self.additive_skips = additive_skips
        self.resample_2d = resample_2d

        emb_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, emb_dim), nn.SiLU(), linear(emb_dim, emb_dim))

        def block(cin, cout=None, **kw):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, num_groups=num_groups, resample_2d=resample_2d, **kw)

        # ---- encoder: stem conv, then per level num_res_blocks ResBlocks and (except at the last level) a down block
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        self._feature_size = model_channels
        chans = [model_channels]
        ch = model_channels
        levels = len(self.channel_mult)
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                self.input_blocks.append(TimestepEmbedSequential(block(ch, mult * model_channels)))
                ch = mult * model_channels
                self._feature_size += ch
                chans.append(ch)
            if level != levels - 1:
                self.input_blocks.append(TimestepEmbedSequential(block(ch, ch, down=True)))
                chans.append(ch)
                self._feature_size += ch
        self.input_block_chans_bk = chans[:]

        # ---- bottleneck
        self.middle_block = TimestepEmbedSequential(block(ch), block(ch))
        self._feature_size += ch

        # ---- decoder: per level num_res_blocks + 1 ResBlocks on cat(h, skip); the last of a level (except level 0)
        # is followed by an up block in the same Sequential
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [block(ch + ich, model_channels * mult)]
                ch = model_channels * mult
                if level and i == num_res_blocks:
                    layers.append(block(ch, ch, up=True))
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch

        self.out = nn.Sequential(normalization(ch, num_groups), nn.SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))
        self._engine = None
        self._train_engine = None

    # The reference's .to() supports a 2-device split and returns None (unet.py:727-752).  Here a 1-element list is
    # unwrapped, a real split is refused, and the module is returned as nn.Module.to does.
    def to(self, *args, **kwargs):
        if args and isinstance(args[0], (list, tuple)):
            if len(args[0]) > 1 and args[0][0] != args[0][1]:
                raise NotImplementedError("splitting UNetModel across two devices is not supported on the fcwdm path "
                                          "(one B200 holds the whole model and its activations)")
            args = (args[0][0],) + tuple(args[1:])
        out = super().to(*args, **kwargs)
        p = next(self.parameters())
        self.devices = [p.device, p.device]
        return out

    def engine(self):
        if self._engine is None:
            from fcwdm.unet_engine import UNetEngine
            object.__setattr__(self, "_engine", UNetEngine(self))
        return self._engine

    def train_engine(self):
        if getattr(self, "_train_engine", None) is None:
            from fcwdm.train_engine import UNetTrainEngine
            object.__setattr__(self, "_train_engine", UNetTrainEngine(self))
        return self._train_engine

    def forward(self, x, timesteps, y=None):
        """x: [N, C, D, H, W] fp32, timesteps: [N] -> [N, out_channels, D, H, W] (reference unet.py:754-800).  Under
        autograd with trainable parameters the forward is taped and the returned tensor carries one autograd node whose
        backward runs the fcwdm backward kernels (as WavUNetModel does)."""
        assert y is None, "must specify y if and only if the model is class-conditional"
        self.hs_shapes = []
        if th.is_grad_enabled():
            params = [p for p in self.parameters()]
            if any(p.requires_grad for p in params):
                if not all(p.requires_grad for p in params):
                    raise NotImplementedError("partially frozen UNetModel: the fcwdm backward produces gradients for all "
                                              "parameters or none")
                from fcwdm.train_engine import WavUNetFunction
                return WavUNetFunction.apply(self.train_engine(), x, timesteps, *params)
            if x.requires_grad:
                raise NotImplementedError("gradient w.r.t. the denoiser input is not implemented by the fcwdm backward")
        return self.engine().forward(x, timesteps)
