"""Execution engine of the wavelet U-Net denoiser on fcwdm kernels.

Walks a ``guided_diffusion.wunet.WavUNetModel`` module tree (the parameter container with the reference's
346 state-dict keys) and runs its forward as a fixed sequence of C-ABI launches on channels-last bf16
activation buffers:

    conv3d (tcgen05 implicit GEMM, bias / timestep-embedding / residual fused in the epilogue)
    GroupNorm + SiLU (two HBM passes)           Haar DWT / IDWT (one pass, /3, x3, embedding add fused)

The launch sequence contains no host synchronisation and no allocation-dependent control flow, so it can be
captured into a CUDA graph (see fcwdm.sampler).  Weights are repacked to bf16 [tap][C_out][C_in] lazily and
again whenever a parameter is modified in place (load_state_dict, optimizer step).
"""
import math

import torch

from . import native, ops
from .native import FcwdmError


def _ld(c):
    return (c + 63) // 64 * 64


def _timesteps(t):
    """(N,) timesteps as the embedding kernel takes them: int64, or float32 when they are fractional (the model sees
    t * 1000 / T with rescale_timesteps=True, gaussian_diffusion.py:417-420 / respace.py:128-132)."""
    return t.float().contiguous() if t.is_floating_point() else t.to(torch.int64).contiguous()


class _Packed:
    __slots__ = ("wp", "bias", "cin", "cout", "k", "pair")


class WavUNetEngine:
    def __init__(self, model):
        self.model = model
        self._sig = None
        self._conv = {}
        self._f32 = {}
        self._device = None
        self._stats = {}
        self._arena = None
        self._arena_pos = 0
        self._emb_w = self._emb_b = None
        self._ptrs = None
        # generation: bumped whenever the packed operand copies are refreshed (parameters changed in place, by version or
        # through invalidate()); buffers_id: bumped only when the persistent device buffers themselves are re-allocated.
        # A captured CUDA graph (fcwdm.sampler) stays valid across generations -- the re-pack writes into the same
        # buffers -- and must be re-captured only when buffers_id changes.
        self.generation = 0
        self.buffers_id = 0
        import os
        # fused GroupNorm statistics in the conv epilogue: measured break-even against the separate (HBM-roofline)
        # statistics pass in round 1, so opt-in
        self.fuse_stats = os.environ.get("FCWDM_FUSED_STATS", "0") == "1"          # single-CTA kernel: opt-in
        self.fuse_stats_pair = os.environ.get("FCWDM_NO_FUSED_STATS_PAIR", "0") != "1"  # pair kernel: register sums, free
        self.fuse_stats_rows = int(os.environ.get("FCWDM_FUSED_STATS_ROWS", "20000"))   # <= 28x28x20 voxels per batch
        self.use_pair = os.environ.get("FCWDM_NO_PAIR", "0") != "1"
        self.fuse_gn_in = os.environ.get("FCWDM_NO_FUSED_GN_IN", "0") != "1"
        self.fuse_gn_in_general = os.environ.get("FCWDM_NO_FUSED_GN_IN_GENERAL", "0") != "1"   # single-CTA kernel too
        # runs of consecutive low-resolution convs (<= chain_rows voxels per launch) go out as ONE persistent launch
        # (csrc/conv3d_chain.cu); only the plain inference engine defers launches, the subclasses interleave other kernels
        self.use_chain = type(self) is WavUNetEngine and os.environ.get("FCWDM_NO_CHAIN", "0") != "1"
        self.chain_rows = int(os.environ.get("FCWDM_CHAIN_ROWS", "20000"))
        self._chain = []
        self.chain_launches = 0

    # ------------------------------------------------------------------ weights
    def _signature(self):
        return tuple((id(p), p._version, p.data_ptr()) for p in self.model.parameters())

    want_dgrad = False      # the training engine also keeps the data-gradient form of every conv weight

    def prepare(self, device):
        sig = self._signature()
        if sig == self._sig and self._device == device:
            return
        self._f32.clear()
        ptrs = tuple(s[2] for s in sig)
        if ptrs != self._ptrs:            # parameter storage moved (.to(), FusedAdamW re-laying): bias / GroupNorm aliases are new
            self._ptrs = ptrs
            self.buffers_id += 1
        convs = [m for m in self.model.modules() if isinstance(m, torch.nn.Conv3d)]
        layout = (str(device), self.want_dgrad, self.use_pair) + tuple(
            (id(m), m.weight.data_ptr(), tuple(m.weight.shape), m.weight.dtype) for m in convs)
        if layout != getattr(self, "_layout", None):
            # (re)build the persistent packed-weight buffers and the one-launch re-pack job table
            self._conv.clear()
            self._conv_t = {}
            jobs, keep = [], []
            for mod in convs:
                k = mod.kernel_size[0]
                if mod.kernel_size != (k, k, k) or k not in (1, 3) or mod.stride != (1, 1, 1) or \
                        mod.padding != (k // 2,) * 3 or mod.groups != 1 or mod.dilation != (1, 1, 1):
                    raise NotImplementedError(f"conv3d configuration not supported by the fcwdm kernel: {mod}")
                if not mod.weight.is_cuda:
                    raise FcwdmError("model parameters are on the CPU; call model.to(cuda_device) first "
                                     "(the fcwdm denoiser has no CPU path)")
                if mod.weight.dtype != torch.float32 or not mod.weight.is_contiguous():
                    raise NotImplementedError("conv weights must be contiguous float32 master weights (use_fp16 / .half() "
                                              "models are not supported: the fcwdm path keeps its own bf16 operand copies)")
                forms = [(self._conv, mod.out_channels, mod.in_channels, 0)]
                if self.want_dgrad and mod.in_channels % 8 == 0:        # the stem conv's input gradient is never needed
                    forms.append((self._conv_t, mod.in_channels, mod.out_channels, 1))
                for table, O, I, transposed in forms:
                    pk = _Packed()
                    # C_in, C_out <= 64 (the full-resolution layers): kd-fused CTA-pair kernel with resident weights
                    pk.pair = self.use_pair and ops.conv3d_pair_supported(I, O, k)
                    n = native.load().fcwdm_conv3d_pair_packed_elems(O, I) if pk.pair else ops.conv3d_packed_elems(O, I, k)
                    pk.wp = torch.empty(n, dtype=torch.bfloat16, device=device)
                    pk.cout, pk.cin, pk.k = O, I, k
                    pk.bias = None
                    table[id(mod)] = pk
                    jobs.append([mod.weight.data_ptr(), pk.wp.data_ptr(), O, I, k ** 3, int(pk.pair), transposed, n])
            self._jobs = torch.tensor(jobs, dtype=torch.int64).to(device)
            self._jobs_max = max(j[7] for j in jobs)
            self._layout = layout
            self._emb_w = self._emb_b = None
            self.buffers_id += 1
        with ops._on(device) as st:
            native.call("fcwdm_conv3d_pack_all", ops._ptr(self._jobs), self._jobs.shape[0], self._jobs_max, st)
        for mod in convs:
            self._conv[id(mod)].bias = mod.bias.detach().float().contiguous() if mod.bias is not None else None
        # all per-ResBlock timestep projections Linear(SiLU(emb)) (wunet.py:203-206,250) as ONE dense layer:
        # rows of W_cat are the concatenated emb_layers[1] weights; a block reads its column slice of the result
        self._emb_off, ws, bs, off = {}, [], [], 0
        for mod in self.model.modules():
            # every ResBlock of either U-Net (wunet.ResBlock / unet.ResBlock): has in_layers + emb_layers
            if hasattr(mod, "emb_layers") and hasattr(mod, "in_layers") and id(mod) not in self._emb_off:
                lin = mod.emb_layers[1]
                self._emb_off[id(mod)] = (off, lin.out_features)
                ws.append(lin.weight.detach().float())
                bs.append(lin.bias.detach().float())
                off += (lin.out_features + 3) // 4 * 4      # keep every slice 16-byte aligned
                if off != self._emb_off[id(mod)][0] + lin.out_features:
                    pad = off - self._emb_off[id(mod)][0] - lin.out_features
                    ws.append(torch.zeros((pad, lin.in_features), device=device))
                    bs.append(torch.zeros((pad,), device=device))
        # persistent buffers, refreshed in place: a captured graph holds their addresses
        emb_w, emb_b = torch.cat(ws, dim=0), torch.cat(bs, dim=0)
        if self._emb_w is None or self._emb_w.shape != emb_w.shape or self._emb_w.device != emb_w.device:
            self._emb_w, self._emb_b = emb_w.contiguous(), emb_b.contiguous()
            self.buffers_id += 1
        else:
            self._emb_w.copy_(emb_w)
            self._emb_b.copy_(emb_b)
        self._sig = sig
        self._device = device
        self.generation += 1

    def invalidate(self):
        """The parameters were updated through raw pointers (e.g. fcwdm.optim.FusedAdamW): re-pack on the next call."""
        self._sig = None
        self.generation += 1

    def _p32(self, p):
        """fp32 contiguous view of a (GroupNorm / Linear) parameter."""
        t = self._f32.get(id(p))
        if t is None:
            t = p.detach().float().contiguous()
            self._f32[id(p)] = t
        return t

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _buf(rows, c, device, ld=None):
        ld = ld or _ld(c)
        if ld == c:
            return torch.empty((rows, ld), dtype=torch.bfloat16, device=device)
        return torch.zeros((rows, ld), dtype=torch.bfloat16, device=device)   # pad channels feed zero weights

    def _conv3d(self, mod, x, N, dims, chan_bias=None, residual=None, out_ld=None, stats_groups=0, gn_in=None, out=None):
        """stats_groups > 0: the conv epilogue also produces the GroupNorm statistics of its output (kept in
        self._stats under the output buffer's id until the consuming GroupNorm picks them up).
        out: write into this (rows, >= C_out) buffer / column-slice view instead of allocating (zero-copy concat)."""
        pk = self._conv[id(mod)]
        rows = N * dims[0] * dims[1] * dims[2]
        y = out if out is not None else self._buf(rows, pk.cout, x.device, out_ld)
        stats = None
        cpg = pk.cout // stats_groups if stats_groups and pk.cout % stats_groups == 0 else 0
        ok_pair = pk.pair and self.fuse_stats_pair and cpg and cpg % 2 == 0
        # single-CTA kernel: break-even against the separate (HBM-roofline) statistics pass on large tensors, but on the
        # low-resolution layers the statistics pass is a latency-bound launch of its own -> fuse there
        chain = self._chainable(pk, rows, x, gn_in)
        ok_single = (not pk.pair) and (self.fuse_stats or rows <= self.fuse_stats_rows or chain) and \
            (cpg in (1, 2, 4) or (cpg and cpg % 8 == 0))
        if stats_groups and stats_groups <= 32 and (ok_pair or ok_single):
            stats = self._stats_slot(N, stats_groups, x.device)
            self._stats[id(y)] = (stats, stats_groups, y)     # holding y keeps its id unique until consumed
        if chain:
            # deferred: becomes one layer of the next fcwdm_conv3d_chain launch (flushed before any other kernel runs)
            self._chain.append(ops.conv3d_chain_layer(x, pk.wp, pk.bias, y, (N,) + tuple(dims), pk.cin, pk.cout,
                                                      chan_bias=chan_bias, residual=residual, gn_stats=stats,
                                                      gn_groups=stats_groups if stats is not None else 0, gn_in=gn_in))
            return y
        self._flush_chain()
        if pk.pair:
            ops.conv3d_pair_cl(x, pk.wp, pk.bias, y, (N,) + tuple(dims), pk.cin, pk.cout, chan_bias=chan_bias,
                               residual=residual, gn_stats=stats, gn_groups=stats_groups if stats is not None else 0,
                               gn_in=gn_in)
        else:
            ops.conv3d_cl(x, pk.wp, pk.bias, y, (N,) + tuple(dims), pk.cin, pk.cout, pk.k, chan_bias=chan_bias,
                          residual=residual, gn_stats=stats, gn_groups=stats_groups if stats is not None else 0,
                          gn_in=gn_in)
        return y

    def _chainable(self, pk, rows, x, gn_in):
        if not self.use_chain or pk.pair or pk.k != 3 or rows > self.chain_rows:
            return False
        if not ops.conv3d_chain_supported(pk.cin, pk.cout, 3) or x.stride(0) < pk.cin:
            return False
        return gn_in is None or pk.cin <= 256

    def _aux_ok(self, rows, *channels):
        """A wavelet re-sampling op can ride in the pending chain launch: low-resolution tensor, supported widths."""
        return self.use_chain and rows <= self.chain_rows and all(ops.chain_aux_supported(c) for c in channels)

    def _aux_stats(self, y, N, gn):
        """Statistics slot for the output y of an in-chain DWT / IDWT whose consumer is GroupNorm `gn`."""
        if gn is None or gn.num_groups > 32 or gn.num_channels % gn.num_groups:
            return None, 0
        stats = self._stats_slot(N, gn.num_groups, y.device)
        self._stats[id(y)] = (stats, gn.num_groups, y)
        return stats, gn.num_groups

    def _flush_chain(self):
        """Launch the deferred run of low-resolution convs (if any) as persistent chain launches."""
        if not self._chain:
            return
        pending, self._chain = self._chain, []
        cap = native.load().fcwdm_conv3d_chain_max_layers()
        for i in range(0, len(pending), cap):
            part = pending[i:i + cap]
            if self._arena is None or self._arena_pos + 1 > self._arena.numel():
                self._arena = torch.zeros(1 << 18, dtype=torch.float64, device=part[0][1][0].device)
                self._arena_pos = 0
            counter = self._arena[self._arena_pos:self._arena_pos + 1]      # 8 zero bytes: the grid-barrier counter
            self._arena_pos += 1
            ops.conv3d_chain([layer for layer, _ in part], counter)
            self.chain_launches += 1

    def _stats_slot(self, N, G, device):
        """Slice of the per-forward statistics arena (zeroed by ONE memset at the start of forward_cl)."""
        n = N * ops.GN_STAT_REPLICAS * G * 2
        if self._arena is None or self._arena_pos + n > self._arena.numel():
            self._arena = torch.zeros(max(n, 1 << 18), dtype=torch.float64, device=device)   # overflow: fresh arena
            self._arena_pos = 0
        out = self._arena[self._arena_pos:self._arena_pos + n].view(N, ops.GN_STAT_REPLICAS, G, 2)
        self._arena_pos += n
        return out

    def _take_stats(self, gn, x, N):
        """(stats buffer, already_filled): statistics left by the producing conv's epilogue, or an empty buffer."""
        pre = self._stats.pop(id(x), None)
        if pre is not None and pre[1] == gn.num_groups and pre[2] is x:
            return pre[0], True
        return torch.empty((N, ops.GN_STAT_REPLICAS, gn.num_groups, 2), dtype=torch.float64, device=x.device), False

    def _gn_silu(self, gn, x, N, S, silu=True):
        C = gn.num_channels
        self._flush_chain()
        y = self._buf(N * S, C, x.device)
        stats, have = self._take_stats(gn, x, N)
        ops.groupnorm_silu(x, y, stats, self._p32(gn.weight), self._p32(gn.bias), N, S, C, gn.num_groups, gn.eps, silu,
                           have_stats=have)
        return y

    def _gn_silu_conv(self, gn, x, mod, N, dims, **kw):
        """conv(SiLU(GroupNorm(x))).  For the CTA-pair kernel the normalisation + activation run inside the conv's
        operand producers (no GroupNorm-apply pass, no intermediate tensor); otherwise apply, then convolve."""
        pk = self._conv[id(mod)]
        S = dims[0] * dims[1] * dims[2]
        fusable = pk.pair or (self.fuse_gn_in_general and pk.k == 3 and pk.cout >= 64 and pk.cin % 64 == 0 and pk.cin <= 256)
        if fusable and self.fuse_gn_in and gn.num_channels == pk.cin and pk.cin % gn.num_groups == 0:
            stats, have = self._take_stats(gn, x, N)
            if not have:
                self._flush_chain()
                ops.groupnorm_stats(x, stats, N, S, pk.cin, gn.num_groups)
            return self._conv3d(mod, x, N, dims, gn_in=(stats, self._p32(gn.weight), self._p32(gn.bias),
                                                         gn.num_groups, gn.eps), **kw)
        return self._conv3d(mod, self._gn_silu(gn, x, N, S), N, dims, **kw)

    def _emb_all(self, emb):
        out = torch.empty((emb.shape[0], self._emb_w.shape[0]), dtype=torch.float32, device=emb.device)
        ops.linear(emb, self._emb_w, self._emb_b, out, act_in=1, act_out=0)                      # Linear(SiLU(emb))
        return out

    def _emb_out(self, blk, emb_all):
        off, n = self._emb_off[id(blk)]
        return emb_all[:, off:off + n]          # strided view: the kernels take the row stride (bias_ld / cb_ld)

    def _resblock(self, blk, x, skip, emb, N, dims):
        """ResBlock.forward (reference wunet.py:223-269).  Returns (out, skip_out, dims_out)."""
        S = dims[0] * dims[1] * dims[2]
        dev = x.device
        cin, cout = blk.channels, blk.out_channels
        emb_out = self._emb_out(blk, emb)
        ssn = getattr(blk, "use_scale_shift_norm", False)
        emb_add = None if ssn else emb_out              # scale-shift norm: emb_out modulates the second GroupNorm instead
        gn1, conv1 = blk.in_layers[0], blk.in_layers[2]
        skip_out = skip
        if blk.down:
            h_full = self._gn_silu_conv(gn1, x, conv1, N, dims)
            d2 = (dims[0] // 2, dims[1] // 2, dims[2] // 2)
            s2 = d2[0] * d2[1] * d2[2]
            h = self._buf(N * s2, cout, dev)
            hi = torch.empty((7, N * s2, _ld(cout)), dtype=torch.bfloat16, device=dev) if _ld(cout) == cout else \
                torch.zeros((7, N * s2, _ld(cout)), dtype=torch.bfloat16, device=dev)
            xs = self._buf(N * s2, cin, dev)
            # h, hSkip = Downsample(h): LLL/3 (+ emb, :262) and the 7 high bands (:118-121, :240); x_upd: LLL/3 only (:241)
            if self._aux_ok(N * S, cout, cin):
                st, g = self._aux_stats(h, N, None if ssn else blk.out_layers[0])
                self._chain.append(ops.chain_dwt_op(h_full, (N,) + tuple(dims), cout, h, hi, lll_bias=emb_add,
                                                    lll_scale=1.0 / 3.0, gn_stats=st, gn_groups=g))
                self._chain.append(ops.chain_dwt_op(x, (N,) + tuple(dims), cin, xs, None, lll_scale=1.0 / 3.0))
            else:
                self._flush_chain()
                ops.dwt3d_cl(h_full, (N,) + tuple(dims), cout, h, hi, lll_bias=emb_add, lll_scale=1.0 / 3.0)
                ops.dwt3d_cl(x, (N,) + tuple(dims), cin, xs, None, lll_scale=1.0 / 3.0)
            x, dims, S, skip_out = xs, d2, s2, hi
        elif blk.up:
            if skip is None:
                raise FcwdmError("up-sampling ResBlock reached without stored high-frequency sub-bands")
            h_low = self._gn_silu_conv(gn1, x, conv1, N, dims)
            d2 = (dims[0] * 2, dims[1] * 2, dims[2] * 2)
            s2 = d2[0] * d2[1] * d2[2]
            h = self._buf(N * s2, cout, dev)
            xu = self._buf(N * s2, cin, dev)
            # h = IDWT(3h, skip) + emb; x = IDWT(3x, skip)
            if self._aux_ok(N * s2, cout, cin) and skip.stride(-2) >= max(cin, cout):
                st, g = self._aux_stats(h, N, None if ssn else blk.out_layers[0])
                self._chain.append(ops.chain_idwt_op(h_low, skip, (N,) + d2, cout, h, bias=emb_add, lll_scale=3.0,
                                                     gn_stats=st, gn_groups=g))
                self._chain.append(ops.chain_idwt_op(x, skip, (N,) + d2, cin, xu, bias=None, lll_scale=3.0))
            else:
                self._flush_chain()
                ops.idwt3d_cl(h_low, skip, (N,) + d2, cout, h, bias=emb_add, lll_scale=3.0)
                ops.idwt3d_cl(x, skip, (N,) + d2, cin, xu, bias=None, lll_scale=3.0)
            x, dims, S, skip_out = xu, d2, s2, None
        else:
            h = self._gn_silu_conv(gn1, x, conv1, N, dims, chan_bias=emb_add,             # conv + emb (:262)
                                   stats_groups=blk.out_layers[0].num_groups)
        if isinstance(blk.skip_connection, torch.nn.Conv3d):
            x = self._conv3d(blk.skip_connection, x, N, dims)
        if ssn:
            a = self._gn_silu_ssn(blk.out_layers[0], h, emb_out, N, S)
            out = self._conv3d(blk.out_layers[3], a, N, dims, residual=x, stats_groups=self.model.num_groups)
        else:
            out = self._gn_silu_conv(blk.out_layers[0], h, blk.out_layers[3], N, dims, residual=x,   # skip(x) + h (:266)
                                     stats_groups=self.model.num_groups)
        return out, skip_out, dims

    def _gn_silu_ssn(self, gn, x, emb_out, N, S):
        """SiLU(GroupNorm(x) * (1 + scale) + shift) with (scale, shift) = chunk(emb_out, 2) per sample (reference
        wunet.py:256-260): the per-sample modulation folds into the affine parameters, gamma' = gamma * (1 + scale),
        beta' = beta * (1 + scale) + shift, so each sample is one fused GroupNorm+SiLU launch with its own (gamma', beta')."""
        C = gn.num_channels
        if emb_out.shape[1] != 2 * C:
            raise FcwdmError(f"scale-shift norm: emb_layers must produce 2 x {C} values, got {emb_out.shape[1]}")
        self._flush_chain()
        y = self._buf(N * S, C, x.device)
        stats, have = self._take_stats(gn, x, N)
        if not have:
            ops.groupnorm_stats(x, stats, N, S, C, gn.num_groups)
        one_plus = 1.0 + emb_out[:, :C].float()
        gamma = (self._p32(gn.weight)[None] * one_plus).contiguous()
        beta = (self._p32(gn.bias)[None] * one_plus + emb_out[:, C:].float()).contiguous()
        for n in range(N):
            ops.groupnorm_silu(x[n * S:(n + 1) * S], y[n * S:(n + 1) * S], stats[n:n + 1], gamma[n], beta[n], 1, S, C,
                               gn.num_groups, gn.eps, True, have_stats=True)
        return y

    # ------------------------------------------------------------------ whole network
    def time_embedding(self, t):
        m = self.model
        te = torch.empty((t.shape[0], m.model_channels), dtype=torch.float32, device=t.device)
        ops.timestep_embedding(t, te, m.model_channels)
        l0, l2 = m.time_embed[0], m.time_embed[2]
        e1 = torch.empty((t.shape[0], l0.out_features), dtype=torch.float32, device=t.device)
        ops.linear(te, self._p32(l0.weight), self._p32(l0.bias), e1, act_in=0, act_out=1)
        e2 = torch.empty((t.shape[0], l2.out_features), dtype=torch.float32, device=t.device)
        ops.linear(e1, self._p32(l2.weight), self._p32(l2.bias), e2, act_in=0, act_out=0)
        return e2

    def forward_cl(self, x_cl, t, N, dims, out_ld=None):
        """x_cl: (N*S, >= round_up(in_channels, 64)) bf16 channels-last; t: (N,) int64 (or float32) CUDA.
        Returns the (N*S, out_ld) bf16 channels-last model output.  Mirrors WavUNetModel.forward
        (reference wunet.py:734-795)."""
        from guided_diffusion.wunet import ResBlock, WaveletDownsample
        m = self.model
        self.prepare(x_cl.device)
        levels = len(m.channel_mult)
        for i, dim in enumerate(dims):
            if dim % (2 ** levels):
                raise FcwdmError(f"spatial size {tuple(dims)} is not divisible by 2^{levels} (one Haar level per "
                                 f"channel_mult entry; the reference fails the same way, SURVEY.md fact 3)")
        self._stats.clear()
        self._chain = []
        self._arena = torch.zeros(1 << 18, dtype=torch.float64, device=x_cl.device)   # one memset per forward
        self._arena_pos = 0
        emb = self._emb_all(self.time_embedding(t))
        hs = []
        pyramid, pyr_dims, pyr_c = x_cl, tuple(dims), m.in_channels
        h, hdims = x_cl, tuple(dims)
        for module in m.input_blocks:
            first = module[0]
            if isinstance(first, WaveletDownsample):
                # input_pyramid = conv(cat(DWT(pyramid)) / 3) + h   (:142-145, :758-760)
                d2 = (pyr_dims[0] // 2, pyr_dims[1] // 2, pyr_dims[2] // 2)
                s2 = d2[0] * d2[1] * d2[2]
                cat = self._buf(N * s2, 8 * pyr_c, x_cl.device)
                if self._aux_ok(N * pyr_dims[0] * pyr_dims[1] * pyr_dims[2], pyr_c):
                    self._chain.append(ops.chain_dwt_op(pyramid, (N,) + pyr_dims, pyr_c, cat[:, :pyr_c], cat[:, pyr_c:],
                                                        lll_scale=1.0 / 3.0, hi_scale=1.0 / 3.0, hi_sb=pyr_c))
                else:
                    self._flush_chain()
                    ops.dwt3d_cl(pyramid, (N,) + pyr_dims, pyr_c, cat[:, :pyr_c], cat[:, pyr_c:], lll_scale=1.0 / 3.0,
                                 hi_scale=1.0 / 3.0, hi_sb=pyr_c)
                pyramid = self._conv3d(first.conv, cat, N, d2, residual=h, stats_groups=m.num_groups)
                pyr_dims, pyr_c = d2, first.out_ch
                h = pyramid
                continue
            skip = None
            if isinstance(first, torch.nn.Conv3d):
                h = self._conv3d(first, h, N, hdims, stats_groups=m.num_groups)
            else:
                for layer in module:
                    if not isinstance(layer, ResBlock):
                        raise NotImplementedError(f"unsupported layer in input_blocks: {type(layer).__name__}")
                    h, skip, hdims = self._resblock(layer, h, None, emb, N, hdims)
            hs.append(skip)
        skip = None
        for layer in m.middle_block:
            h, skip, hdims = self._resblock(layer, h, None, emb, N, hdims)
            skip = None                                                   # `h, skip = h` overwrites with None (:765)
        for module in m.output_blocks:
            new_hs = hs.pop()
            if new_hs is not None:
                skip = new_hs
            cur = skip
            for layer in module:
                h, cur, hdims = self._resblock(layer, h, cur, emb, N, hdims)
        for module in m.out_res:
            for layer in module:
                h, _, hdims = self._resblock(layer, h, None, emb, N, hdims)
        out = self._gn_silu_conv(m.out[0], h, m.out[2], N, hdims,
                                 out_ld=out_ld or max(8, (m.out_channels + 7) // 8 * 8))
        self._flush_chain()
        return out

    def forward(self, x, timesteps):
        """Planar fp32 API of the reference: x (N, C, D, H, W), timesteps (N,) -> (N, out_channels, D, H, W)."""
        m = self.model
        if not x.is_cuda:
            raise FcwdmError("WavUNetModel.forward: input is on the CPU; the fcwdm denoiser has no CPU path")
        if x.dim() != 5 or x.shape[1] != m.in_channels:
            raise ValueError(f"expected input of shape (N, {m.in_channels}, D, H, W), got {tuple(x.shape)}")
        N, C, D, H, W = x.shape
        S = D * H * W
        with torch.cuda.device(x.device):
            x_cl = torch.zeros((N * S, _ld(C)), dtype=torch.bfloat16, device=x.device) if _ld(C) != C else \
                torch.empty((N * S, C), dtype=torch.bfloat16, device=x.device)
            ops.planar_to_cl(x.float(), x_cl, C)
            out_cl = self.forward_cl(x_cl, _timesteps(timesteps), N, (D, H, W))
            out = torch.empty((N, m.out_channels, D, H, W), dtype=torch.float32, device=x.device)
            ops.cl_to_planar(out_cl, out, m.out_channels)
        return out.to(x.dtype) if x.dtype != torch.float32 else out
