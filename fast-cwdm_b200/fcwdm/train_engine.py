"""Training execution of the wavelet U-Net denoiser on fcwdm kernels: forward with a tape, explicit backward.

The reference trains ``WavUNetModel`` through autograd (scripts/train.py -> TrainLoop.forward_backward,
guided_diffusion/train_util.py:396-460): cuDNN dgrad/wgrad for every nn.Conv3d, native GroupNorm / SiLU backward,
the matmul chains of DWT_IDWT_Functions.py:139-156,184-208 for the wavelet up/down-sampling.  Here the forward is
the same launch sequence as ``WavUNetEngine`` (GroupNorm-apply materialised, because its output is the wgrad
operand) and records one closure per launch; the backward replays the closures in reverse, each a C-ABI launch:

    conv3d  ->  dX: the forward tcgen05 kernel on dY with transposed, tap-reversed weights (gradient fan-in fused as
                its residual add);  dW: fcwdm_conv3d_wgrad (tcgen05, MN-major operands);  db / d(timestep
                embedding): fcwdm_colsum_cl
    GroupNorm+SiLU -> fcwdm_groupnorm_bwd;   Haar DWT / IDWT -> their adjoints;   Linear -> fcwdm_linear_bwd

Activation gradients are channels-last bf16; parameter gradients are fp32 views into ONE flat buffer laid out in
``model.parameters()`` order (what the bucketed NCCL all-reduce of fcwdm.ddp and the fused AdamW of fcwdm.optim
operate on).  Weight-tied ResBlocks accumulate both uses into the same slice.
"""
import torch

from . import ops
from .engine import WavUNetEngine, _ld, _Packed, _timesteps
from .native import FcwdmError


class WavUNetTrainEngine(WavUNetEngine):
    def __init__(self, model):
        super().__init__(model)
        self._tsig = None
        self._conv_t = {}
        self._gflat = None
        self._gview = {}
        self._tape = []
        self._grads = {}
        self._keep = []
        self._uses = {}
        self._wants_cs = {}              # id(conv output) -> C_out: its gradient's column sums are wanted (bias / embedding)
        self._cs = {}                    # id(dx) -> (column sums taken by the GroupNorm backward that wrote dx, dx)
        self._cs_arena = None
        self._cs_pos = 0
        import os
        self.fuse_colsum = os.environ.get("FCWDM_NO_FUSED_COLSUM", "0") != "1"
        self.grad_ready_hook = None      # callable(lo, hi): flat-gradient range [lo, hi) is final (fcwdm.ddp)
        self.grad_sync = None            # fcwdm.ddp.GradSync: bucketed all-reduce overlapped with the backward

    # ------------------------------------------------------------------ parameters / gradients
    def flat_grad(self, device):
        params = list(self.model.parameters())
        total = sum((p.numel() + 3) // 4 * 4 for p in params)
        if self._gflat is None or self._gflat.device != device or self._gflat.numel() != total:
            self._gflat = torch.zeros(total, dtype=torch.float32, device=device)
            self._gview, self._goff, off = {}, {}, 0
            for p in params:
                self._gview[id(p)] = self._gflat[off:off + p.numel()].view(p.shape)
                self._goff[id(p)] = (off, off + p.numel())
                off += (p.numel() + 3) // 4 * 4          # 16-byte aligned slices
        return self._gflat

    def _gp(self, p):
        return self._gview[id(p)]

    want_dgrad = True       # WavUNetEngine.prepare then also packs w'[ci][co][tap] = w[co][ci][26-tap] into self._conv_t

    def prepare_train(self, device):
        self.prepare(device)

    # ------------------------------------------------------------------ gradient bookkeeping
    def _take(self, t):
        return self._grads.pop(id(t), None)

    def _partial(self, t):
        return self._grads.get(id(t))

    def _set(self, t, g):
        self._grads[id(t)] = g

    def _pass(self, t, g, rows, C):
        """Identity edge: d(t) += g."""
        cur = self._grads.get(id(t))
        if cur is None:
            self._grads[id(t)] = g
        else:
            out = self._buf(rows, C, g.device)
            ops.add_cl(cur, g, out, rows, C)
            self._grads[id(t)] = out

    def _cs_slot(self, N, C, device):
        """(N, C) zeroed fp32 slice of the per-backward column-sum arena (one memset per backward)."""
        n = N * C
        if self._cs_arena is None or self._cs_arena.device != device or self._cs_pos + n > self._cs_arena.numel():
            self._cs_arena = torch.zeros(max(n, 1 << 17), dtype=torch.float32, device=device)
            self._cs_pos = 0
        out = self._cs_arena[self._cs_pos:self._cs_pos + n].view(N, C)
        self._cs_pos += n
        return out

    @staticmethod
    def _buf7(rows, c, device):
        """The 7 high-frequency bands (7, rows, ld); pad channels (C < ld) must be zero: they meet zero weights in a conv."""
        ld = _ld(c)
        alloc = torch.empty if ld == c else torch.zeros
        return alloc((7, rows, ld), dtype=torch.bfloat16, device=device)

    def _zeros_like_grad(self, rows, C, device):
        return torch.zeros((rows, _ld(C)), dtype=torch.bfloat16, device=device)

    # ------------------------------------------------------------------ taped building blocks
    def _conv3d_t(self, mod, x, N, dims, emb=None, residual=None, out_ld=None, stats_groups=0, need_dx=True, out=None):
        """emb = (d_emb_all, off, n, emb_slice): the timestep-embedding add fused in the conv epilogue."""
        pk = self._conv[id(mod)]
        rows = N * dims[0] * dims[1] * dims[2]
        y = WavUNetEngine._conv3d(self, mod, x, N, dims, chan_bias=emb[3] if emb else None, residual=residual,
                                  out_ld=out_ld, stats_groups=stats_groups, out=out)
        self._keep.append((x, y, residual))
        self._count(mod.weight, mod.bias)
        dims4 = (N,) + tuple(dims)
        if (emb is not None or mod.bias is not None) and pk.cout % 8 == 0:
            self._wants_cs[id(y)] = pk.cout      # a GroupNorm backward that writes d(y) can sum its columns on the way

        def bwd():
            dy = self._take(y)
            if dy is None:
                return
            if residual is not None:
                self._pass(residual, dy, rows, pk.cout)
            # bias gradient (+ per-sample timestep-embedding gradient) from one pass over dY
            cs = _ld(pk.cout) if pk.cout % 8 else pk.cout
            db = self._gp(mod.bias) if mod.bias is not None else None
            if emb is not None or db is not None:
                if pk.cout % 8:
                    raise FcwdmError("conv3d backward: C_out must be a multiple of 8")
                out_sample = emb[0][:, emb[1]:emb[1] + emb[2]] if emb else None
                pre = self._cs.pop(id(dy), None)
                if pre is not None and pre[1] is dy:         # dy is exactly what that GroupNorm backward stored
                    ops.colsum_scatter(pre[0], N, pk.cout, out_sample=out_sample, out_total=db)
                else:
                    ops.colsum_cl(dy, N, rows // N, cs, out_sample=out_sample, out_total=db)
            ops.conv3d_wgrad(x, dy, self._gp(mod.weight), dims4, pk.cin, pk.cout, pk.k, accumulate=True)
            if need_dx:
                pt = self._conv_t[id(mod)]
                dx = self._buf(rows, pk.cin, dy.device)
                acc = self._partial(x)
                if pt.pair:
                    ops.conv3d_pair_cl(dy, pt.wp, None, dx, dims4, pt.cin, pt.cout, residual=acc)
                else:
                    ops.conv3d_cl(dy, pt.wp, None, dx, dims4, pt.cin, pt.cout, pt.k, residual=acc)
                self._set(x, dx)
            self._param_done(mod.weight, mod.bias)

        self._tape.append(bwd)
        return y

    def _gn_silu_t(self, gn, x, N, S, silu=True):
        C = gn.num_channels
        y = self._buf(N * S, C, x.device)
        stats, have = self._take_stats(gn, x, N)
        gamma, beta = self._p32(gn.weight), self._p32(gn.bias)
        ops.groupnorm_silu(x, y, stats, gamma, beta, N, S, C, gn.num_groups, gn.eps, silu, have_stats=have)
        self._keep.append((x, y))
        self._count(gn.weight, gn.bias)

        def bwd():
            dy = self._take(y)
            if dy is None:
                return
            dx = self._buf(N * S, C, x.device)
            cs = None
            if self.fuse_colsum and self._wants_cs.get(id(x)) == C:
                # x is a conv output whose bias / embedding gradient is the column sum of d(x): taken in this pass.  The
                # conv's backward uses it only if it receives THIS tensor (no later fan-in was added to it)
                cs = self._cs_slot(N, C, x.device)
                self._cs[id(dx)] = (cs, dx)
            ops.groupnorm_bwd(x, dy, stats, gamma, beta, dx, self._gp(gn.weight), self._gp(gn.bias), N, S, C,
                              gn.num_groups, gn.eps, silu, acc=self._partial(x), colsum=cs)
            self._set(x, dx)
            self._param_done(gn.weight, gn.bias)

        self._tape.append(bwd)
        return y

    def _gn_silu_conv(self, gn, x, mod, N, dims, **kw):       # training: GroupNorm output materialised (wgrad operand)
        S = dims[0] * dims[1] * dims[2]
        return self._conv3d_t(mod, self._gn_silu_t(gn, x, N, S), N, dims, **kw)

    def _count(self, *params):
        """One more tape entry contributes to these parameters (weight-tied blocks are taped twice)."""
        for p in params:
            if p is not None:
                self._uses[id(p)] = self._uses.get(id(p), 0) + 1

    def _param_done(self, *params):
        if self.grad_ready_hook is not None:
            for p in params:
                if p is not None:
                    self._uses_left[id(p)] -= 1
                    if self._uses_left[id(p)] == 0:
                        self.grad_ready_hook(*self._goff[id(p)])

    def _emb_slice(self, blk, emb_all, d_emb_all):
        off, n = self._emb_off[id(blk)]
        return (d_emb_all, off, n, emb_all[:, off:off + n])

    def _dwt_t(self, x, dims4, C, lll, hi, emb=None, lll_scale=1.0 / 3.0, hi_scale=1.0, hi_sb=None, need_dx=True,
               cat=None):
        N, D, H, W = dims4
        ops.dwt3d_cl(x, dims4, C, lll, hi, lll_bias=emb[3] if emb else None, lll_scale=lll_scale, hi_scale=hi_scale,
                     hi_sb=hi_sb)
        self._keep.append((x, lll, hi))
        rows2 = N * (D // 2) * (H // 2) * (W // 2)
        # WaveletDownsample: lll and hi are column slices of ONE tensor `cat` (the conv's operand): its gradient is keyed
        # on `cat`

        def bwd():
            if cat is not None:
                dcat = self._take(cat)
                if dcat is None:
                    return
                dl, dh = dcat[:, :C], dcat[:, C:]
            else:
                dl = self._take(lll)
                dh = self._take(hi) if hi is not None else None
                if dl is None and dh is None:
                    return
                if dl is None:
                    dl = self._zeros_like_grad(rows2, C, x.device)
            if emb is not None:
                ops.colsum_cl(dl, N, rows2 // N, C, out_sample=emb[0][:, emb[1]:emb[1] + emb[2]])
            if not need_dx:
                return
            dx = self._buf(N * D * H * W, C, x.device)
            ops.dwt3d_cl_bwd(dl, dh, dims4, C, dx, acc=self._partial(x), lll_scale=lll_scale, hi_scale=hi_scale,
                             hi_sb=hi_sb)
            self._set(x, dx)

        self._tape.append(bwd)

    def _idwt_t(self, lll, hi, dims4, C, y, emb=None, lll_scale=3.0):
        N, D, H, W = dims4
        ops.idwt3d_cl(lll, hi, dims4, C, y, bias=emb[3] if emb else None, lll_scale=lll_scale)
        self._keep.append((lll, hi, y))
        rows2 = N * (D // 2) * (H // 2) * (W // 2)

        def bwd():
            dy = self._take(y)
            if dy is None:
                return
            if emb is not None:
                ops.colsum_cl(dy, N, D * H * W, C, out_sample=emb[0][:, emb[1]:emb[1] + emb[2]])
            dl = self._buf(rows2, C, dy.device)
            dh = self._partial(hi)
            accumulate = dh is not None
            if dh is None:
                dh = self._buf7(rows2, C, dy.device)
            ops.idwt3d_cl_bwd(dy, dims4, C, dl, dh, lll_acc=self._partial(lll), hi_accumulate=accumulate,
                              lll_scale=lll_scale)
            self._set(lll, dl)
            self._set(hi, dh)

        self._tape.append(bwd)

    def _resblock_t(self, blk, x, skip, emb_all, d_emb_all, N, dims):
        """ResBlock.forward (reference wunet.py:223-269) with the backward taped."""
        dev = x.device
        cin, cout = blk.channels, blk.out_channels
        if blk.dropout:
            raise NotImplementedError("dropout > 0 in training is not implemented (run.sh ships dropout=0)")
        ssn = getattr(blk, "use_scale_shift_norm", False)
        emb_full = self._emb_slice(blk, emb_all, d_emb_all)
        emb = None if ssn else emb_full                 # scale-shift norm: emb_out modulates the second GroupNorm instead
        gn1, conv1 = blk.in_layers[0], blk.in_layers[2]
        skip_out = skip
        d4 = (N,) + tuple(dims)
        if blk.down:
            h_full = self._gn_silu_conv(gn1, x, conv1, N, dims)
            d2 = (dims[0] // 2, dims[1] // 2, dims[2] // 2)
            s2 = d2[0] * d2[1] * d2[2]
            h = self._buf(N * s2, cout, dev)
            hi = self._buf7(N * s2, cout, dev)
            self._dwt_t(h_full, d4, cout, h, hi, emb=emb, lll_scale=1.0 / 3.0)
            xs = self._buf(N * s2, cin, dev)
            self._dwt_t(x, d4, cin, xs, None, lll_scale=1.0 / 3.0)
            x, dims, skip_out = xs, d2, hi
        elif blk.up:
            if skip is None:
                raise FcwdmError("up-sampling ResBlock reached without stored high-frequency sub-bands")
            h_low = self._gn_silu_conv(gn1, x, conv1, N, dims)
            d2 = (dims[0] * 2, dims[1] * 2, dims[2] * 2)
            s2 = d2[0] * d2[1] * d2[2]
            h = self._buf(N * s2, cout, dev)
            self._idwt_t(h_low, skip, (N,) + d2, cout, h, emb=emb, lll_scale=3.0)
            xu = self._buf(N * s2, cin, dev)
            self._idwt_t(x, skip, (N,) + d2, cin, xu, lll_scale=3.0)
            x, dims, skip_out = xu, d2, None
        else:
            h = self._gn_silu_conv(gn1, x, conv1, N, dims, emb=emb, stats_groups=blk.out_layers[0].num_groups)
        if isinstance(blk.skip_connection, torch.nn.Conv3d):
            x = self._conv3d_t(blk.skip_connection, x, N, dims)
        if ssn:
            a = self._gn_silu_ssn_t(blk.out_layers[0], h, emb_full, N, dims[0] * dims[1] * dims[2])
            out = self._conv3d_t(blk.out_layers[3], a, N, dims, residual=x, stats_groups=self.model.num_groups)
        else:
            out = self._gn_silu_conv(blk.out_layers[0], h, blk.out_layers[3], N, dims, residual=x,
                                     stats_groups=self.model.num_groups)
        return out, skip_out, dims

    def _gn_silu_ssn_t(self, gn, x, emb, N, S):
        """Taped SiLU(GroupNorm(x) * (1 + scale) + shift), (scale, shift) = chunk(emb_out, 2) (reference wunet.py:256-260).
        Forward: per sample, the fused GroupNorm+SiLU with gamma'_n = gamma (1 + scale_n), beta'_n = beta (1 + scale_n)
        + shift_n.  Backward: the ordinary GroupNorm backward per sample gives dx and (dgamma'_n, dbeta'_n); then
        dgamma = sum_n dgamma'_n (1 + scale_n), dbeta = sum_n dbeta'_n (1 + scale_n),
        dscale_n = dgamma'_n gamma + dbeta'_n beta, dshift_n = dbeta'_n (into the timestep-embedding gradient)."""
        C = gn.num_channels
        d_emb_all, off, width, emb_out = emb
        if width != 2 * C:
            raise FcwdmError(f"scale-shift norm: emb_layers must produce 2 x {C} values, got {width}")
        y = self._buf(N * S, C, x.device)
        stats, have = self._take_stats(gn, x, N)
        if not have:
            ops.groupnorm_stats(x, stats, N, S, C, gn.num_groups)
        gam, bet = self._p32(gn.weight), self._p32(gn.bias)
        one_plus = (1.0 + emb_out[:, :C].float()).contiguous()
        gamma_n = (gam[None] * one_plus).contiguous()
        beta_n = (bet[None] * one_plus + emb_out[:, C:].float()).contiguous()
        for n in range(N):
            ops.groupnorm_silu(x[n * S:(n + 1) * S], y[n * S:(n + 1) * S], stats[n:n + 1], gamma_n[n], beta_n[n], 1, S, C,
                               gn.num_groups, gn.eps, True, have_stats=True)
        self._keep.append((x, y))
        self._count(gn.weight, gn.bias)

        def bwd():
            dy = self._take(y)
            if dy is None:
                return
            dx = self._buf(N * S, C, x.device)
            acc = self._partial(x)
            dgp = torch.zeros((N, C), dtype=torch.float32, device=x.device)
            dbp = torch.zeros((N, C), dtype=torch.float32, device=x.device)
            for n in range(N):
                rows = slice(n * S, (n + 1) * S)
                ops.groupnorm_bwd(x[rows], dy[rows], stats[n:n + 1], gamma_n[n], beta_n[n], dx[rows], dgp[n], dbp[n], 1, S, C,
                                  gn.num_groups, gn.eps, True, acc=acc[rows] if acc is not None else None)
            self._gp(gn.weight).add_((dgp * one_plus).sum(dim=0))
            self._gp(gn.bias).add_((dbp * one_plus).sum(dim=0))
            d_emb_all[:, off:off + C] += dgp * gam[None] + dbp * bet[None]
            d_emb_all[:, off + C:off + 2 * C] += dbp
            self._set(x, dx)
            self._param_done(gn.weight, gn.bias)

        self._tape.append(bwd)
        return y

    def _time_path_t(self, t, N, dev):
        """Timestep path (wunet.py:472-475,736 / unet.py:777; ResBlock emb_layers) with pre-activations kept and the
        backward taped.  Returns (emb_all, d_emb_all): all per-block projections and their gradient accumulator."""
        m = self.model
        l0, l2 = m.time_embed[0], m.time_embed[2]
        te = torch.empty((N, m.model_channels), dtype=torch.float32, device=dev)
        ops.timestep_embedding(t, te, m.model_channels)
        z1 = torch.empty((N, l0.out_features), dtype=torch.float32, device=dev)
        ops.linear(te, self._p32(l0.weight), self._p32(l0.bias), z1, act_in=0, act_out=0)
        e2 = torch.empty((N, l2.out_features), dtype=torch.float32, device=dev)
        ops.linear(z1, self._p32(l2.weight), self._p32(l2.bias), e2, act_in=1, act_out=0)
        emb_all = WavUNetEngine._emb_all(self, e2)
        d_emb_all = torch.zeros_like(emb_all)
        blocks, seen = [], set()
        for mod in m.modules():
            if hasattr(mod, "emb_layers") and hasattr(mod, "in_layers") and id(mod) not in seen:
                seen.add(id(mod))
                blocks.append(mod)
                self._count(mod.emb_layers[1].weight, mod.emb_layers[1].bias)
        self._count(l2.weight, l2.bias, l0.weight, l0.bias)

        def emb_bwd():
            for mod in blocks:
                lin = mod.emb_layers[1]
                off, n = self._emb_off[id(mod)]
                ops.linear_bwd(e2, None, d_emb_all[:, off:off + n], dW=self._gp(lin.weight), db=self._gp(lin.bias), act_in=1)
                self._param_done(lin.weight, lin.bias)
            d_e2 = torch.empty_like(e2)
            ops.linear_bwd(e2, self._emb_w, d_emb_all, dx=d_e2, act_in=1)
            d_z1 = torch.empty_like(z1)
            ops.linear_bwd(z1, self._p32(l2.weight), d_e2, dx=d_z1, dW=self._gp(l2.weight), db=self._gp(l2.bias), act_in=1)
            ops.linear_bwd(te, None, d_z1, dW=self._gp(l0.weight), db=self._gp(l0.bias), act_in=0)
            self._param_done(l2.weight, l2.bias, l0.weight, l0.bias)

        self._tape.append(emb_bwd)
        return emb_all, d_emb_all

    # ------------------------------------------------------------------ whole network
    def forward_train(self, x, timesteps):
        """Planar fp32 (N, C, D, H, W) -> (N, out_channels, D, H, W), recording the backward tape."""
        from guided_diffusion.wunet import ResBlock, WaveletDownsample
        m = self.model
        if not x.is_cuda:
            raise FcwdmError("WavUNetModel.forward: input is on the CPU; the fcwdm denoiser has no CPU path")
        if x.dim() != 5 or x.shape[1] != m.in_channels:
            raise ValueError(f"expected input of shape (N, {m.in_channels}, D, H, W), got {tuple(x.shape)}")
        N, C, D, H, W = x.shape
        dims = (D, H, W)
        levels = len(m.channel_mult)
        for dim in dims:
            if dim % (2 ** levels):
                raise FcwdmError(f"spatial size {dims} is not divisible by 2^{levels}")
        dev = x.device
        with torch.cuda.device(dev):
            self.prepare_train(dev)
            self.flat_grad(dev)
            self._tape, self._grads, self._keep, self._uses = [], {}, [], {}
            self._wants_cs = {}
            self._stats.clear()
            self._arena = torch.zeros(1 << 18, dtype=torch.float64, device=dev)
            self._arena_pos = 0
            S = D * H * W
            x_cl = torch.zeros((N * S, _ld(C)), dtype=torch.bfloat16, device=dev)
            ops.planar_to_cl(x.detach().float(), x_cl, C)
            t = _timesteps(timesteps)

            emb_all, d_emb_all = self._time_path_t(t, N, dev)

            # ---- U-Net (mirrors WavUNetEngine.forward_cl / reference wunet.py:734-795)
            hs = []
            pyramid, pyr_dims, pyr_c = x_cl, dims, m.in_channels
            h, hdims = x_cl, dims
            first_pyramid = True
            for module in m.input_blocks:
                first = module[0]
                if isinstance(first, WaveletDownsample):
                    d2 = (pyr_dims[0] // 2, pyr_dims[1] // 2, pyr_dims[2] // 2)
                    s2 = d2[0] * d2[1] * d2[2]
                    cat = self._buf(N * s2, 8 * pyr_c, dev)
                    self._dwt_t(pyramid, (N,) + pyr_dims, pyr_c, cat[:, :pyr_c], cat[:, pyr_c:], lll_scale=1.0 / 3.0,
                                hi_scale=1.0 / 3.0, hi_sb=pyr_c, need_dx=not first_pyramid, cat=cat)
                    pyramid = self._conv3d_t(first.conv, cat, N, d2, residual=h, stats_groups=m.num_groups,
                                             need_dx=not first_pyramid)     # level 0: cat derives from the input only
                    first_pyramid = False
                    pyr_dims, pyr_c = d2, first.out_ch
                    h = pyramid
                    continue
                skip = None
                if isinstance(first, torch.nn.Conv3d):
                    h = self._conv3d_t(first, h, N, hdims, stats_groups=m.num_groups, need_dx=False)
                else:
                    for layer in module:
                        if not isinstance(layer, ResBlock):
                            raise NotImplementedError(f"unsupported layer in input_blocks: {type(layer).__name__}")
                        h, skip, hdims = self._resblock_t(layer, h, None, emb_all, d_emb_all, N, hdims)
                hs.append(skip)
            for layer in m.middle_block:
                h, _, hdims = self._resblock_t(layer, h, None, emb_all, d_emb_all, N, hdims)
            skip = None
            for module in m.output_blocks:
                new_hs = hs.pop()
                if new_hs is not None:
                    skip = new_hs
                cur = skip
                for layer in module:
                    h, cur, hdims = self._resblock_t(layer, h, cur, emb_all, d_emb_all, N, hdims)
            for module in m.out_res:
                for layer in module:
                    h, _, hdims = self._resblock_t(layer, h, None, emb_all, d_emb_all, N, hdims)
            out_cl = self._gn_silu_conv(m.out[0], h, m.out[2], N, hdims, out_ld=max(8, (m.out_channels + 7) // 8 * 8))
            self._out_cl = out_cl
            self._shape = (N, D, H, W)
            out = torch.empty((N, m.out_channels, D, H, W), dtype=torch.float32, device=dev)
            ops.cl_to_planar(out_cl, out, m.out_channels)
        return out.to(x.dtype) if x.dtype != torch.float32 else out

    def backward(self, dout):
        """dout: planar (N, out_channels, D, H, W).  Runs the tape in reverse; returns the flat fp32 gradient (views
        per parameter through .grad_views())."""
        m = self.model
        N, D, H, W = self._shape
        dev = dout.device
        with torch.cuda.device(dev):
            self._gflat.zero_()
            self._cs, self._cs_pos = {}, 0
            if self._cs_arena is not None:
                self._cs_arena.zero_()
            if self.grad_ready_hook is not None:
                self._uses_left = dict(self._uses)
            if self.grad_sync is not None:
                self.grad_sync.begin()
            dy = torch.zeros((N * D * H * W, _ld(m.out_channels)), dtype=torch.bfloat16, device=dev)
            ops.planar_to_cl(dout.detach().float().contiguous(), dy, m.out_channels)
            self._set(self._out_cl, dy)
            for fn in reversed(self._tape):
                fn()
            if self.grad_sync is not None:
                self.grad_sync.finish()
            self._tape, self._grads, self._keep, self._out_cl = [], {}, [], None
            self._cs, self._wants_cs = {}, {}
            self._stats.clear()
        return self._gflat

    def grad_views(self):
        """Per-parameter views of a COPY of the flat gradient (autograd may adopt them as .grad and later accumulate
        into them in place; the engine's own buffer is rewritten by the next backward).  The copy is one contiguous
        buffer (self.last_flat) so fcwdm.optim.FusedAdamW can step on it with one launch."""
        self.last_flat = self._gflat.clone()
        return [self.last_flat[lo:hi].view(p.shape) for p in self.model.parameters()
                for lo, hi in (self._goff[id(p)],)]

    def flat_offsets(self):
        return [self._goff[id(p)] for p in self.model.parameters()]


class WavUNetFunction(torch.autograd.Function):
    """Autograd boundary: one node for the whole denoiser.  Inputs: the engine, x, timesteps, then every parameter
    (so autograd routes their gradients to .grad / DDP hooks).  The gradient w.r.t. x is not computed (the training
    loss, gaussian_diffusion.py:1084-1166, never needs it)."""

    @staticmethod
    def forward(ctx, engine, x, timesteps, *params):
        ctx.engine = engine
        ctx.n_params = len(params)
        return engine.forward_train(x, timesteps)

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.engine
        eng.backward(dout)
        return (None, None, None) + tuple(eng.grad_views())


class UNetTrainEngine(WavUNetTrainEngine):
    """Training execution of the plain UNetModel (run.sh's use_freq=False model, the one scripts/train.py actually
    trains): UNetEngine's launch plan with every primitive taped.  The skip concatenations stay zero-copy in the
    backward as well: the gradient of a concat buffer is split into two column-slice views."""

    def __init__(self, model):
        super().__init__(model)
        self._emb_map = {}
        self._x_in = None
        self._d_emb_all = None

    # UNetEngine's forward plan, with this class's taped primitives underneath
    from .unet_engine import UNetEngine as _U
    forward_cl = _U.forward_cl
    _resblock_u = _U._resblock_u
    del _U

    def time_embedding(self, t):
        return t                                            # the taped time path runs in _emb_all

    def _emb_all(self, t):
        emb_all, self._d_emb_all = self._time_path_t(t, t.shape[0], t.device)
        return emb_all

    def _emb_out(self, blk, emb_all):
        off, n = self._emb_off[id(blk)]
        v = emb_all[:, off:off + n]
        self._emb_map[id(v)] = (self._d_emb_all, off, n, v)     # holding v keeps its id unique for this forward
        return v

    def _conv3d(self, mod, x, N, dims, chan_bias=None, residual=None, out_ld=None, stats_groups=0, gn_in=None, out=None):
        emb = self._emb_map[id(chan_bias)] if chan_bias is not None else None
        return self._conv3d_t(mod, x, N, dims, emb=emb, residual=residual, out_ld=out_ld, stats_groups=stats_groups,
                              need_dx=x is not self._x_in, out=out)

    def _gn_silu(self, gn, x, N, S, silu=True):
        return self._gn_silu_t(gn, x, N, S, silu)

    def _gn_silu_ssn(self, gn, x, emb_out, N, S):
        """use_scale_shift_norm=True (unet.py:297-309; the default of the reference's model_and_diffusion_defaults): the
        taped per-sample GroupNorm of the wavelet U-Net's training engine, keyed by the embedding slice view."""
        return self._gn_silu_ssn_t(gn, x, self._emb_map[id(emb_out)], N, S)

    def _gn_silu_conv(self, gn, x, mod, N, dims, **kw):
        S = dims[0] * dims[1] * dims[2]
        return self._conv3d(mod, self._gn_silu(gn, x, N, S), N, dims, **kw)

    def _resample(self, x, N, dims, C, up, depth):
        fd = 2 if depth else 1
        d2 = (dims[0] * fd, dims[1] * 2, dims[2] * 2) if up else (dims[0] // fd, dims[1] // 2, dims[2] // 2)
        y = self._buf(N * d2[0] * d2[1] * d2[2], C, x.device)
        if up:
            ops.upsample2_cl(x, (N,) + tuple(dims), C, y, up_depth=depth)
        else:
            ops.avgpool2_cl(x, (N,) + tuple(dims), C, y, pool_depth=depth)
        self._keep.append((x, y))
        rows_x = N * dims[0] * dims[1] * dims[2]

        def bwd():
            dy = self._take(y)
            if dy is None:
                return
            dx = self._buf(rows_x, C, x.device)
            if up:
                ops.upsample2_cl_bwd(dy, (N,) + d2, C, dx, acc=self._partial(x), up_depth=depth)
            else:
                ops.avgpool2_cl_bwd(dy, (N,) + d2, C, dx, acc=self._partial(x), pool_depth=depth)
            self._set(x, dx)

        self._tape.append(bwd)
        return y, d2

    def _on_concat(self, cat, left, right, plan_j, rows):
        width, off = plan_j
        self._keep.append((cat, left, right))

        def bwd():
            dcat = self._take(cat)
            if dcat is None:
                return
            self._pass(left, dcat[:, :off], rows, off)
            self._pass(right, dcat[:, off:width], rows, width - off)

        self._tape.append(bwd)

    def forward_train(self, x, timesteps):
        m = self.model
        if not x.is_cuda:
            raise FcwdmError("UNetModel.forward: input is on the CPU; the fcwdm denoiser has no CPU path")
        if x.dim() != 5 or x.shape[1] != m.in_channels:
            raise ValueError(f"expected input of shape (N, {m.in_channels}, D, H, W), got {tuple(x.shape)}")
        for mod in m.modules():
            if getattr(mod, "dropout", 0) and hasattr(mod, "in_layers"):
                raise NotImplementedError("dropout > 0 in training is not implemented (run.sh ships dropout=0)")
        N, C, D, H, W = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            self.prepare_train(dev)
            self.flat_grad(dev)
            self._tape, self._grads, self._keep, self._uses, self._emb_map = [], {}, [], {}, {}
            self._wants_cs = {}
            S = D * H * W
            x_cl = torch.zeros((N * S, _ld(C)), dtype=torch.bfloat16, device=dev)
            ops.planar_to_cl(x.detach().float(), x_cl, C)
            self._x_in = x_cl
            out_cl = self.forward_cl(x_cl, _timesteps(timesteps), N, (D, H, W))
            self._out_cl = out_cl
            self._shape = (N, D, H, W)
            out = torch.empty((N, m.out_channels, D, H, W), dtype=torch.float32, device=dev)
            ops.cl_to_planar(out_cl, out, m.out_channels)
        return out.to(x.dtype) if x.dtype != torch.float32 else out
