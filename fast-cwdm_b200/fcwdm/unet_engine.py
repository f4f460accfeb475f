"""Execution engine of the plain UNetModel (guided_diffusion/unet.py of the reference, run.sh's use_freq=False model)
on fcwdm kernels.  Same building blocks and launch-plan style as ``WavUNetEngine`` (tcgen05 conv3d with bias /
timestep embedding / residual in the epilogue, fused GroupNorm+SiLU, CTA-pair conv with GroupNorm in the operand path
for the 64-channel full-resolution layers); what differs from the wavelet U-Net:

* ResBlock up/down-sampling is average-pool / nearest-neighbour (reference unet.py:285-311: GroupNorm+SiLU, THEN
  resample h and x, THEN the 3x3x3 conv) -> ``fcwdm_avgpool2_cl`` / ``fcwdm_upsample2_cl``;
* skip connections are channel concatenations ``th.cat([h, hs.pop()], dim=1)`` (unet.py:796).  No copy is made: the
  encoder output that will be concatenated later is written by its producing conv straight into the right-hand column
  slice of the future concat buffer (and read from there, with the buffer's row stride, by the next encoder block);
  the decoder tensor it meets is written into the left-hand slice by ITS producing conv.  Every kernel takes row strides,
  so a column slice is an ordinary operand.
"""
import torch

from . import ops
from .engine import WavUNetEngine, _ld
from .native import FcwdmError


class UNetEngine(WavUNetEngine):
    def _resample(self, x, N, dims, C, up, depth):
        """avg-pool (up=False) or nearest x2 (up=True) of a cl buffer; depth: also along D (resample_2d=False)."""
        fd = 2 if depth else 1
        if up:
            d2 = (dims[0] * fd, dims[1] * 2, dims[2] * 2)
        else:
            d2 = (dims[0] // fd, dims[1] // 2, dims[2] // 2)
        y = self._buf(N * d2[0] * d2[1] * d2[2], C, x.device)
        if up:
            ops.upsample2_cl(x, (N,) + tuple(dims), C, y, up_depth=depth)
        else:
            ops.avgpool2_cl(x, (N,) + tuple(dims), C, y, pool_depth=depth)
        return y, d2

    def _resblock_u(self, blk, x, emb, N, dims, out=None):
        """unet.ResBlock._forward (reference unet.py:285-311).  Returns (out, dims_out)."""
        S = dims[0] * dims[1] * dims[2]
        cin = blk.channels
        emb_out = self._emb_out(blk, emb)
        ssn = getattr(blk, "use_scale_shift_norm", False)
        emb_add = None if ssn else emb_out              # scale-shift norm: emb_out modulates the second GroupNorm instead
        gn1, conv1 = blk.in_layers[0], blk.in_layers[2]
        g2 = blk.out_layers[0].num_groups
        if blk.updown:
            depth = not blk.resample_2d
            a = self._gn_silu(gn1, x, N, S)                                    # in_rest(x)                (:288)
            a, d2 = self._resample(a, N, dims, cin, blk.up, depth)             # h = h_upd(h)              (:289)
            x, _ = self._resample(x, N, dims, cin, blk.up, depth)              # x = x_upd(x)              (:290)
            dims = d2
            h = self._conv3d(conv1, a, N, dims, chan_bias=emb_add, stats_groups=g2)   # in_conv(h) + emb_out (:291,:308)
        else:
            h = self._gn_silu_conv(gn1, x, conv1, N, dims, chan_bias=emb_add, stats_groups=g2)
        if isinstance(blk.skip_connection, torch.nn.Conv3d):
            x = self._conv3d(blk.skip_connection, x, N, dims)
        if ssn:                                                                # out_norm(h) * (1 + scale) + shift (:301-305)
            a2 = self._gn_silu_ssn(blk.out_layers[0], h, emb_out, N, dims[0] * dims[1] * dims[2])
            res = self._conv3d(blk.out_layers[3], a2, N, dims, residual=x, stats_groups=self.model.num_groups, out=out)
        else:
            res = self._gn_silu_conv(blk.out_layers[0], h, blk.out_layers[3], N, dims, residual=x,   # skip(x) + h (:311)
                                     stats_groups=self.model.num_groups, out=out)
        return res, dims

    def _gn_silu_conv(self, gn, x, mod, N, dims, **kw):
        # the CTA-pair kernel's fused operand-path GroupNorm reads 64 channels per voxel: the operand must not be a
        # narrower column slice whose neighbours belong to another tensor
        pk = self._conv[id(mod)]
        if pk.pair and self.fuse_gn_in and x.stride(0) != _ld(pk.cin) and gn.num_channels < 64:
            S = dims[0] * dims[1] * dims[2]
            return self._conv3d(mod, self._gn_silu(gn, x, N, S), N, dims, **kw)
        return super()._gn_silu_conv(gn, x, mod, N, dims, **kw)

    def forward_cl(self, x_cl, t, N, dims, out_ld=None):
        """x_cl: (N*S, >= round_up(in_channels, 64)) bf16 channels-last; t: (N,) int64 CUDA.  Mirrors UNetModel.forward
        (reference unet.py:754-800)."""
        m = self.model
        dev = x_cl.device
        self.prepare(dev)
        n_down = len(m.channel_mult) - 1
        for i, dim in enumerate(dims):
            if (i > 0 or not m.resample_2d) and dim % (2 ** n_down):
                raise FcwdmError(f"spatial size {tuple(dims)} is not divisible by 2^{n_down} (one halving per level; the "
                                 f"reference's th.cat of mismatched skip shapes fails the same way)")
        self._stats.clear()
        self._arena = torch.zeros(1 << 18, dtype=torch.float64, device=dev)
        self._arena_pos = 0
        emb = self._emb_all(self.time_embedding(t))

        # ---- static concat plan: encoder output j meets the decoder in output_blocks[n-1-j], whose first ResBlock
        # has channels = ch_dec + ich; the skip goes to columns [ch_dec, ch_dec + ich) of that block's input buffer
        n_in = len(m.input_blocks)
        enc_out_ch = []
        for module in m.input_blocks:
            last = module[-1]
            enc_out_ch.append(last.out_channels)           # Conv3d and ResBlock both expose out_channels
        plan = []                                           # per encoder index j: (width, offset)
        for j in range(n_in):
            width = m.output_blocks[n_in - 1 - j][0].channels
            plan.append((width, width - enc_out_ch[j]))

        def cat_buf(rows, j):
            # operands narrower than 64 channels are read 64 wide by the TMA box (the extra columns meet zero weights):
            # inside a concat buffer those columns belong to the neighbour slice / the next row, which must then hold
            # finite values from the start -> zero-initialise unless every slice is a multiple of 64 channels
            width, off = plan[j]
            if off % 64 or (width - off) % 64:
                return torch.zeros((rows + 1, _ld(width)), dtype=torch.bfloat16, device=dev)[:rows]   # +1: the last row's overhang
            return self._buf(rows, width, dev)

        cats = [None] * n_in
        enc_views = [None] * n_in                              # the encoder outputs as written (right-hand column slices)
        h, hdims = x_cl, tuple(dims)
        hs_dims = []
        for j, module in enumerate(m.input_blocks):
            rows_in = N * hdims[0] * hdims[1] * hdims[2]
            first = module[0]
            width, off = plan[j]
            if isinstance(first, torch.nn.Conv3d):
                cat = cat_buf(rows_in, j)
                view = cat[:, off:off + first.out_channels]
                h = self._conv3d(first, h, N, hdims, stats_groups=m.num_groups, out=view)
            else:
                for li, layer in enumerate(module):
                    if not hasattr(layer, "in_layers"):
                        raise NotImplementedError(f"unsupported layer in input_blocks: {type(layer).__name__}")
                    if li == len(module) - 1:
                        od = hdims
                        if layer.updown:
                            fd = 1 if layer.resample_2d else 2
                            od = (hdims[0] // fd, hdims[1] // 2, hdims[2] // 2)
                        cat = cat_buf(N * od[0] * od[1] * od[2], j)
                        view = cat[:, off:off + layer.out_channels]
                        h, hdims = self._resblock_u(layer, h, emb, N, hdims, out=view)
                    else:
                        h, hdims = self._resblock_u(layer, h, emb, N, hdims)
            cats[j] = cat
            enc_views[j] = h
            hs_dims.append(hdims)

        # ---- bottleneck: the second block writes into the left slice of the first decoder input
        k_cat = n_in - 1
        lefts = [cats[j][:, :plan[j][1]] for j in range(n_in)]      # decoder-side column slices (one view object each)
        h, hdims = self._resblock_u(m.middle_block[0], h, emb, N, hdims)
        h, hdims = self._resblock_u(m.middle_block[1], h, emb, N, hdims, out=lefts[k_cat])

        # ---- decoder
        for k, module in enumerate(m.output_blocks):
            j = n_in - 1 - k
            if hs_dims[j] != hdims:
                raise FcwdmError(f"skip connection {j} has spatial size {hs_dims[j]}, decoder has {hdims}")
            h = cats[j]                                          # == th.cat([h, hs.pop()], dim=1), already in place
            self._on_concat(h, lefts[j], enc_views[j], plan[j], N * hdims[0] * hdims[1] * hdims[2])
            nxt = lefts[j - 1] if j > 0 else None
            for li, layer in enumerate(module):
                last = li == len(module) - 1
                h, hdims = self._resblock_u(layer, h, emb, N, hdims, out=nxt if last else None)
        return self._gn_silu_conv(m.out[0], h, m.out[2], N, hdims,
                                  out_ld=out_ld or max(8, (m.out_channels + 7) // 8 * 8))

    def _on_concat(self, cat, left, right, plan_j, rows):
        """Hook: `cat` (the decoder block's input) is the in-place concatenation of `left` (decoder tensor) and `right`
        (encoder output); the training engine records the gradient split here."""
