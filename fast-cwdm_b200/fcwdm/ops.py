"""Torch-tensor wrappers over the C-ABI (fcwdm/native.py).  PyTorch is plumbing here: it owns device memory
and streams; every computation below is a hand-written sm_100a kernel in libfcwdm.so.  CPU tensors are
rejected -- there is no fallback path."""
import ctypes

import torch

from . import native
from .native import FcwdmError

_VP = ctypes.c_void_p
GN_STAT_REPLICAS = 16   # FCWDM_GN_STAT_REPLICAS in include/fcwdm.h


def _ptr(t):
    return _VP(t.data_ptr()) if t is not None else _VP(None)


def _stream(dev):
    return _VP(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise FcwdmError(f"{what}: tensor is on {t.device}; the fcwdm operators run on CUDA (sm_100a) only and "
                         "have no CPU fallback")


class _on:
    """Make `dev` current for the launch (the reference's .cuda() used the *current* device,
    DWT_IDWT_layer.py:505-511; here the input's device decides) and make sure the library is initialised.
    Fast path (every launch of a training / sampling step goes through here): when `dev` already is the current
    device and the library is initialised for it, no guard object is created."""

    __slots__ = ("dev", "guard")

    def __init__(self, dev):
        self.dev = dev
        self.guard = None

    def __enter__(self):
        idx = self.dev.index
        cur = torch.cuda.current_device()
        if idx is None:
            idx = cur
        if idx != cur:
            self.guard = torch.cuda.device(self.dev)
            self.guard.__enter__()
        if idx not in native._inited_devices:
            native.init(idx)
        return _VP(torch.cuda.current_stream(self.dev).cuda_stream)

    def __exit__(self, *a):
        if self.guard is not None:
            return self.guard.__exit__(*a)
        return False


def _dtype_code(t):
    if t.dtype == torch.float32:
        return native.FCWDM_F32
    if t.dtype == torch.bfloat16:
        return native.FCWDM_BF16
    # the reference's matmul against fp32 band matrices raises "expected scalar type Float" for anything else
    raise TypeError(f"DWT/IDWT: unsupported dtype {t.dtype} (float32 and bfloat16 are implemented)")


def _spatial_contig(x):
    """Return x with (D,H,W) contiguous, keeping arbitrary N/C strides when possible."""
    _, _, D, H, W = x.shape
    if x.stride(4) == 1 and x.stride(3) == W and x.stride(2) == H * W:
        return x
    return x.contiguous()


# ----------------------------------------------------------------------------------------------------
# planar DWT / IDWT
# ----------------------------------------------------------------------------------------------------
def dwt3d_planar(x, lll_scale=1.0, concat=False):
    """x (N,C,D,H,W) -> 8 bands.  concat=False: tensor (8,N,C,d,h,w); concat=True: (N,8*C,d,h,w) laid out as
    th.cat([LLL*lll_scale, LLH, ...], dim=1)."""
    _need_cuda(x, "DWT_3D")
    if x.dim() != 5:
        raise AssertionError("DWT_3D expects a 5-D (N, C, D, H, W) input")   # DWT_IDWT_layer.py:525
    x = _spatial_contig(x)
    N, C, D, H, W = x.shape
    if D % 2 or H % 2 or W % 2:
        raise FcwdmError(f"DWT_3D: D, H, W must be even, got {(D, H, W)}")
    d, h, w = D // 2, H // 2, W // 2
    s = d * h * w
    if concat:
        out = torch.empty((N, 8 * C, d, h, w), dtype=x.dtype, device=x.device)
        o_sn, o_sc, o_sb = 8 * C * s, s, C * s
    else:
        out = torch.empty((8, N, C, d, h, w), dtype=x.dtype, device=x.device)
        o_sn, o_sc, o_sb = C * s, s, N * C * s
    with _on(x.device) as st:
        native.call("fcwdm_dwt3d_fwd", _ptr(x), _ptr(out), _dtype_code(x), N, C, D, H, W, x.stride(0), x.stride(1),
                    o_sn, o_sc, o_sb, float(lll_scale), st)
    return out


def idwt3d_planar(bands, lll_scale=1.0, concat=False):
    """Inverse of dwt3d_planar.  bands: (8,N,C,d,h,w) or, with concat=True, (N,8*C,d,h,w)."""
    _need_cuda(bands, "IDWT_3D")
    bands = bands.contiguous()
    if concat:
        N, C8, d, h, w = bands.shape
        C = C8 // 8
        s = d * h * w
        b_sn, b_sc, b_sb = 8 * C * s, s, C * s
    else:
        _, N, C, d, h, w = bands.shape
        s = d * h * w
        b_sn, b_sc, b_sb = C * s, s, N * C * s
    y = torch.empty((N, C, 2 * d, 2 * h, 2 * w), dtype=bands.dtype, device=bands.device)
    with _on(bands.device) as st:
        native.call("fcwdm_idwt3d_fwd", _ptr(bands), _ptr(y), _dtype_code(bands), N, C, 2 * d, 2 * h, 2 * w, b_sn, b_sc,
                    b_sb, y.stride(0), y.stride(1), float(lll_scale), st)
    return y


class DWT3DFunction(torch.autograd.Function):
    """Autograd wrapper; backward of the analysis is the synthesis of the 8 upstream gradients
    (DWT_IDWT_Functions.py:139-156)."""

    @staticmethod
    def forward(ctx, x):
        out = dwt3d_planar(x)
        return tuple(out[i] for i in range(8))

    @staticmethod
    def backward(ctx, *grads):
        g = torch.stack([gi.contiguous() for gi in grads], dim=0)
        return idwt3d_planar(g)


class IDWT3DFunction(torch.autograd.Function):
    """Backward of the synthesis is the analysis of the upstream gradient (DWT_IDWT_Functions.py:184-208)."""

    @staticmethod
    def forward(ctx, *bands):
        shapes = {tuple(b.shape) for b in bands}
        if len(shapes) != 1:
            raise FcwdmError(f"IDWT_3D: the 8 sub-bands must share one shape, got {sorted(shapes)}")
        return idwt3d_planar(torch.stack([b.contiguous() for b in bands], dim=0))

    @staticmethod
    def backward(ctx, grad):
        out = dwt3d_planar(grad.contiguous())
        return tuple(out[i] for i in range(8))


# ----------------------------------------------------------------------------------------------------
# channels-last (bf16) building blocks used by the denoiser
# ----------------------------------------------------------------------------------------------------
def dwt3d_cl(x, dims, C, lll, hi, lll_bias=None, lll_scale=1.0 / 3.0, hi_scale=1.0, hi_sb=None):
    """x: cl buffer (voxels, ld) for dims=(N,D,H,W).  lll / hi: destination cl buffers (hi may be None)."""
    N, D, H, W = dims
    with _on(x.device) as st:
        native.call("fcwdm_dwt3d_cl", _ptr(x), x.stride(0), _ptr(lll), lll.stride(0), _ptr(hi),
                    hi.stride(-2) if hi is not None else 0,
                    (hi_sb if hi_sb is not None else (hi.stride(0) if hi is not None else 0)), _ptr(lll_bias),
                    lll_bias.stride(0) if lll_bias is not None else 0, N, D, H, W, C, float(lll_scale), float(hi_scale), st)


def idwt3d_cl(lll, hi, dims_out, C, y, bias=None, lll_scale=3.0):
    N, D, H, W = dims_out
    with _on(y.device) as st:
        native.call("fcwdm_idwt3d_cl", _ptr(lll), lll.stride(0), _ptr(hi), hi.stride(-2), hi.stride(0), _ptr(y),
                    y.stride(0), _ptr(bias), bias.stride(0) if bias is not None else 0, N, D, H, W, C, float(lll_scale), st)


def planar_to_cl(src, dst, C):
    """src (N,C,...) f32 planar -> dst cl bf16 (N*S, ld) channels [0, C)."""
    N = src.shape[0]
    S = src[0, 0].numel()
    src = src.contiguous()
    with _on(src.device) as st:
        native.call("fcwdm_planar_to_cl", _ptr(src), _ptr(dst), dst.stride(0), N, C, S, st)


def cl_to_planar(src, dst, C):
    N = dst.shape[0]
    S = dst[0, 0].numel()
    with _on(dst.device) as st:
        native.call("fcwdm_cl_to_planar", _ptr(src), src.stride(0), _ptr(dst), N, C, S, st)


def groupnorm_silu(x, y, stats, gamma, beta, N, S, C, G, eps=1e-5, silu=True, have_stats=False):
    """have_stats=True: `stats` was already filled by the producing conv's epilogue (fused statistics)."""
    with _on(x.device) as st:
        if not have_stats:
            native.call("fcwdm_groupnorm_stats", _ptr(x), x.stride(0), _ptr(stats), N, S, C, G, st)
        native.call("fcwdm_groupnorm_apply", _ptr(x), x.stride(0), _ptr(y), y.stride(0), _ptr(stats), _ptr(gamma),
                    _ptr(beta), N, S, C, G, float(eps), 1 if silu else 0, st)


def groupnorm_stats(x, stats, N, S, C, G):
    with _on(x.device) as st:
        native.call("fcwdm_groupnorm_stats", _ptr(x), x.stride(0), _ptr(stats), N, S, C, G, st)


def timestep_embedding(t, out, dim, max_period=10000.0):
    """t: (N,) int64, or float32 for fractional timesteps (rescale_timesteps=True)."""
    if t.dtype == torch.int64:
        name = "fcwdm_timestep_embedding"
    elif t.dtype == torch.float32:
        name = "fcwdm_timestep_embedding_f32"
    else:
        raise TypeError(f"timesteps must be int64 or float32, got {t.dtype}")
    t = t.contiguous()
    with _on(t.device) as st:
        native.call(name, _ptr(t), _ptr(out), t.shape[0], dim, float(max_period), st)


def linear(x, W, b, y, act_in=0, act_out=0):
    N, K = x.shape
    M = W.shape[0]
    with _on(x.device) as st:
        native.call("fcwdm_linear", _ptr(x), _ptr(W), _ptr(b), _ptr(y), N, K, M, act_in, act_out, st)


def conv3d_packed_elems(cout, cin, k):
    return native.load().fcwdm_conv3d_packed_elems(cout, cin, k)


def conv3d_pack_weights(w):
    """w: (Cout, Cin, k, k, k) float32 CUDA -> packed bf16 [k^3][Cout_p][Cin_p]."""
    _need_cuda(w, "conv3d_pack_weights")
    cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
    wp = torch.empty(conv3d_packed_elems(cout, cin, k), dtype=torch.bfloat16, device=w.device)
    w = w.detach().float().contiguous()
    with _on(w.device) as st:
        native.call("fcwdm_conv3d_pack_weights", _ptr(w), _ptr(wp), cout, cin, k, st)
    return wp


def conv3d_cl(x, wp, bias, y, dims, cin, cout, k, chan_bias=None, residual=None, gn_stats=None, gn_groups=0, gn_in=None):
    """x, y, residual: cl bf16 buffers (voxels, ld).  dims = (N, D, H, W).  gn_stats: optional PRE-ZEROED
    (N, GN_STAT_REPLICAS, gn_groups, 2) float64 buffer that receives the GroupNorm statistics of y.
    gn_in = (stats, gamma, beta, groups, eps): convolve SiLU(GroupNorm(x)), normalised in the kernel's operand path."""
    N, D, H, W = dims
    if gn_in is not None:
        assert k == 3
        with _on(x.device) as st:
            native.call("fcwdm_conv3d_gn_fwd", _ptr(x), x.stride(0), _ptr(wp), _ptr(bias), _ptr(chan_bias),
                        chan_bias.stride(0) if chan_bias is not None else 0, _ptr(residual),
                        residual.stride(0) if residual is not None else 0, _ptr(y), y.stride(0), _ptr(gn_stats), gn_groups,
                        _ptr(gn_in[0]), _ptr(gn_in[1]), _ptr(gn_in[2]), gn_in[3], float(gn_in[4]), N, D, H, W, cin, cout, st)
        return
    with _on(x.device) as st:
        native.call("fcwdm_conv3d_fwd", _ptr(x), x.stride(0), _ptr(wp), _ptr(bias), _ptr(chan_bias),
                    chan_bias.stride(0) if chan_bias is not None else 0, _ptr(residual),
                    residual.stride(0) if residual is not None else 0, _ptr(y), y.stride(0), _ptr(gn_stats), gn_groups,
                    N, D, H, W, cin, cout, k, st)


def conv3d_chain_supported(cin, cout, k):
    return bool(native.load().fcwdm_conv3d_chain_supported(cin, cout, k))


def conv3d_chain_layer(x, wp, bias, y, dims, cin, cout, chan_bias=None, residual=None, gn_stats=None, gn_groups=0, gn_in=None):
    """Describe one layer of a conv3d_chain launch (same arguments as conv3d_cl, k = 3).  Returns (struct, keep-alive)."""
    N, D, H, W = dims
    L = native.ChainLayer()
    L.x, L.x_ld, L.wp = x.data_ptr(), x.stride(0), wp.data_ptr()
    L.bias = bias.data_ptr() if bias is not None else None
    L.chan_bias = chan_bias.data_ptr() if chan_bias is not None else None
    L.cb_ld = chan_bias.stride(0) if chan_bias is not None else 0
    L.residual = residual.data_ptr() if residual is not None else None
    L.res_ld = residual.stride(0) if residual is not None else 0
    L.y, L.y_ld = y.data_ptr(), y.stride(0)
    L.gn_stats = gn_stats.data_ptr() if gn_stats is not None else None
    L.gn_groups = gn_groups if gn_stats is not None else 0
    if gn_in is not None:
        L.gn_in_stats, L.gn_in_gamma, L.gn_in_beta = gn_in[0].data_ptr(), gn_in[1].data_ptr(), gn_in[2].data_ptr()
        L.gn_in_groups, L.gn_in_eps = gn_in[3], float(gn_in[4])
    L.N, L.D, L.H, L.W, L.Cin, L.Cout = N, D, H, W, cin, cout
    return L, (x, wp, bias, y, chan_bias, residual, gn_stats, gn_in)


CHAIN_CONV, CHAIN_DWT, CHAIN_IDWT = 0, 1, 2


def chain_aux_supported(C):
    return C in (64, 128, 256)


def chain_dwt_op(x, dims, C, lll, hi, lll_bias=None, lll_scale=1.0 / 3.0, hi_scale=1.0, hi_sb=None, gn_stats=None, gn_groups=0):
    """dwt3d_cl as an op of a conv3d_chain launch (same arguments); gn_stats: statistics of the LLL output."""
    N, D, H, W = dims
    L = native.ChainLayer()
    L.kind = CHAIN_DWT
    L.x, L.x_ld, L.y, L.y_ld = x.data_ptr(), x.stride(0), lll.data_ptr(), lll.stride(0)
    if hi is not None:
        L.aux, L.aux_ld = hi.data_ptr(), hi.stride(-2)
        L.aux_sb = hi_sb if hi_sb is not None else hi.stride(0)
    if lll_bias is not None:
        L.chan_bias, L.cb_ld = lll_bias.data_ptr(), lll_bias.stride(0)
    L.lll_scale, L.hi_scale = float(lll_scale), float(hi_scale)
    if gn_stats is not None:
        L.gn_stats, L.gn_groups = gn_stats.data_ptr(), gn_groups
    L.N, L.D, L.H, L.W, L.Cin, L.Cout = N, D, H, W, C, C
    return L, (x, lll, hi, lll_bias, gn_stats)


def chain_idwt_op(lll, hi, dims_out, C, y, bias=None, lll_scale=3.0, gn_stats=None, gn_groups=0):
    """idwt3d_cl as an op of a conv3d_chain launch (same arguments); gn_stats: statistics of the output y."""
    N, D, H, W = dims_out
    L = native.ChainLayer()
    L.kind = CHAIN_IDWT
    L.x, L.x_ld, L.y, L.y_ld = lll.data_ptr(), lll.stride(0), y.data_ptr(), y.stride(0)
    L.aux, L.aux_ld, L.aux_sb = hi.data_ptr(), hi.stride(-2), hi.stride(0)
    if bias is not None:
        L.chan_bias, L.cb_ld = bias.data_ptr(), bias.stride(0)
    L.lll_scale, L.hi_scale = float(lll_scale), 1.0
    if gn_stats is not None:
        L.gn_stats, L.gn_groups = gn_stats.data_ptr(), gn_groups
    L.N, L.D, L.H, L.W, L.Cin, L.Cout = N, D, H, W, C, C
    return L, (lll, hi, y, bias, gn_stats)


def conv3d_chain(layers, sync_counter):
    """Run a list of conv3d_chain_layer structs as ONE persistent launch (csrc/conv3d_chain.cu).  sync_counter: a device
    tensor whose first 4 bytes are zero (the grid-barrier counter)."""
    n = len(layers)
    arr = (native.ChainLayer * n)(*layers)
    with _on(sync_counter.device) as st:
        native.call("fcwdm_conv3d_chain", arr, n, _ptr(sync_counter), st)


def conv3d_pair_supported(cin, cout, k):
    return bool(native.load().fcwdm_conv3d_pair_supported(cin, cout, k))


def conv3d_pair_pack_weights(w):
    """w: (Cout<=64, Cin<=64, 3, 3, 3) float32 CUDA -> packed bf16 [kh*3+kw][kd][Cout_p][64] for the CTA-pair kernel."""
    _need_cuda(w, "conv3d_pair_pack_weights")
    cout, cin = w.shape[0], w.shape[1]
    wp = torch.empty(native.load().fcwdm_conv3d_pair_packed_elems(cout, cin), dtype=torch.bfloat16, device=w.device)
    w = w.detach().float().contiguous()
    with _on(w.device) as st:
        native.call("fcwdm_conv3d_pair_pack_weights", _ptr(w), _ptr(wp), cout, cin, st)
    return wp


def conv3d_pair_cl(x, wp, bias, y, dims, cin, cout, chan_bias=None, residual=None, gn_stats=None, gn_groups=0,
                   gn_in=None):
    """kd-fused two-CTA conv (3x3x3, C_in <= 64, C_out <= 64); same arguments as conv3d_cl.
    gn_in = (stats, gamma, beta, groups, eps): convolve SiLU(GroupNorm(x)) with the normalisation fused into the
    operand producers (x is then the raw tensor)."""
    N, D, H, W = dims
    with _on(x.device) as st:
        native.call("fcwdm_conv3d_pair_fwd", _ptr(x), x.stride(0), _ptr(wp), _ptr(bias), _ptr(chan_bias),
                    chan_bias.stride(0) if chan_bias is not None else 0, _ptr(residual),
                    residual.stride(0) if residual is not None else 0, _ptr(y), y.stride(0), _ptr(gn_stats), gn_groups,
                    _ptr(gn_in[0]) if gn_in else _VP(None), _ptr(gn_in[1]) if gn_in else _VP(None),
                    _ptr(gn_in[2]) if gn_in else _VP(None), gn_in[3] if gn_in else 0, float(gn_in[4]) if gn_in else 0.0,
                    N, D, H, W, cin, cout, st)


# ----------------------------------------------------------------------------------------------------
# diffusion step
# ----------------------------------------------------------------------------------------------------
def p_sample_step(model_out, x_t, noise, coef, t, clip_denoised=True, predict_xstart=True, want_pred=True,
                  model_out_cl_ld=0, x_prev_cl=None):
    """One fused reverse step.  model_out: planar f32 (N,8,d,h,w) or (model_out_cl_ld > 0) a cl bf16 buffer.
    Returns (x_prev, pred_xstart or None)."""
    _need_cuda(x_t, "p_sample")
    N, C, d, h, w = x_t.shape
    assert C == 8, "the wavelet-domain sample has 8 sub-band channels"
    x_t = x_t.contiguous()
    noise = noise.contiguous()
    if model_out_cl_ld == 0:
        model_out = model_out.contiguous()
    x_prev = torch.empty_like(x_t)
    pred = torch.empty_like(x_t) if want_pred else None
    with _on(x_t.device) as st:
        native.call("fcwdm_p_sample_step", _ptr(model_out), model_out_cl_ld, _ptr(x_t), _ptr(noise), _ptr(x_prev),
                    _ptr(pred), _ptr(x_prev_cl), x_prev_cl.stride(0) if x_prev_cl is not None else 0, _ptr(coef), _ptr(t),
                    coef.shape[0], N, d, h, w, 1 if clip_denoised else 0, 1 if predict_xstart else 0, st)
    return x_prev, pred


def q_sample(x_start, noise, coef, t):
    _need_cuda(x_start, "q_sample")
    x_start = x_start.contiguous()
    noise = noise.contiguous()
    out = torch.empty_like(x_start)
    N = x_start.shape[0]
    with _on(x_start.device) as st:
        native.call("fcwdm_q_sample", _ptr(x_start), _ptr(noise), _ptr(out), _ptr(coef), _ptr(t), coef.shape[0], N,
                    x_start[0].numel() if N else 0, st)
    return out


def sample_to_image(sample, cond_1=None):
    """(N,8,d,h,w) wavelet sample -> (N,1,2d,2h,2w) image, clamped to [0,1] and masked where cond_1 == 0
    (scripts/sample.py:113-125)."""
    _need_cuda(sample, "sample_to_image")
    sample = sample.contiguous()
    N, C, d, h, w = sample.shape
    assert C == 8
    img = torch.empty((N, 1, 2 * d, 2 * h, 2 * w), dtype=torch.float32, device=sample.device)
    if cond_1 is not None:
        cond_1 = cond_1.contiguous()
        assert cond_1.numel() == img.numel()
    with _on(sample.device) as st:
        native.call("fcwdm_sample_to_image", _ptr(sample), _ptr(cond_1), _ptr(img), N, d, h, w, st)
    return img


# ----------------------------------------------------------------------------------------------------
# training path: backward kernels (csrc/conv3d_wgrad.cu, csrc/train.cu)
# ----------------------------------------------------------------------------------------------------
_wgrad_ws = {}


def _wgrad_workspace(nbytes, device):
    """One grow-only scratch buffer per device for the wgrad split partial sums (stream-ordered reuse)."""
    ws = _wgrad_ws.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _wgrad_ws[device] = ws
    return ws


def conv3d_wgrad(x, dy, dw, dims, cin, cout, k, accumulate=True):
    """dw (Cout, Cin, k, k, k) f32 (+)= sum_v dy[v] (x) x[v + tap].  x: the conv's input (voxels, x_ld) cl bf16,
    dy: gradient of its output (voxels, dy_ld >= round_up(Cout, 64)) cl bf16."""
    N, D, H, W = dims
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == cout * cin * k ** 3
    nbytes = native.load().fcwdm_conv3d_wgrad_workspace_bytes(N, D, H, W, cin, cout, k)
    if nbytes < 0:
        raise FcwdmError("conv3d_wgrad: bad shape")
    ws = _wgrad_workspace(nbytes, x.device)
    with _on(x.device) as st:
        native.call("fcwdm_conv3d_wgrad", _ptr(x), x.stride(0), _ptr(dy), dy.stride(0), _ptr(dw), _ptr(ws), ws.numel(),
                    1 if accumulate else 0, N, D, H, W, cin, cout, k, st)


def conv3d_transpose_flip_weights(w):
    """(Cout, Cin, k, k, k) f32 -> (Cin, Cout, k, k, k) f32 with reversed taps: the data-gradient conv's weights."""
    _need_cuda(w, "conv3d_transpose_flip_weights")
    cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
    w = w.detach().float().contiguous()
    wt = torch.empty((cin, cout, k, k, k), dtype=torch.float32, device=w.device)
    with _on(w.device) as st:
        native.call("fcwdm_conv3d_transpose_flip_weights", _ptr(w), _ptr(wt), cout, cin, k, st)
    return wt


def groupnorm_bwd(x, dy, stats, gamma, beta, dx, dgamma, dbeta, N, S, C, G, eps=1e-5, silu=True, acc=None, colsum=None):
    """colsum: optional zero-initialised (N, >= C) fp32 buffer that receives the per-sample column sums of dx (the bias /
    timestep-embedding gradient of the conv that produced x) from the same pass."""
    sums = torch.empty((N, GN_STAT_REPLICAS, C, 2), dtype=torch.float64, device=x.device)
    with _on(x.device) as st:
        native.call("fcwdm_groupnorm_bwd_colsum", _ptr(x), x.stride(0), _ptr(dy), dy.stride(0), _ptr(stats), _ptr(gamma),
                    _ptr(beta), _ptr(sums), _ptr(acc), acc.stride(0) if acc is not None else 0, _ptr(dx), dx.stride(0),
                    _ptr(dgamma), _ptr(dbeta), _ptr(colsum), colsum.stride(0) if colsum is not None else 0, N, S, C, G,
                    float(eps), 1 if silu else 0, st)


def colsum_scatter(part, N, C, out_sample=None, out_total=None):
    """out_sample[n, c] += part[n, c]; out_total[c] += sum_n part[n, c] (part: the colsum buffer of groupnorm_bwd)."""
    with _on(part.device) as st:
        native.call("fcwdm_colsum_scatter", _ptr(part), part.stride(0), _ptr(out_sample),
                    out_sample.stride(0) if out_sample is not None else 0, _ptr(out_total), N, C, st)


def colsum_cl(x, N, S, C, out_sample=None, out_total=None):
    with _on(x.device) as st:
        native.call("fcwdm_colsum_cl", _ptr(x), x.stride(0), _ptr(out_sample),
                    out_sample.stride(0) if out_sample is not None else 0, _ptr(out_total), N, S, C, st)


def dwt3d_cl_bwd(dlll, dhi, dims, C, dx, acc=None, lll_scale=1.0 / 3.0, hi_scale=1.0, hi_sb=None):
    """Adjoint of dwt3d_cl.  dims = (N, D, H, W) of dx (the DWT's input)."""
    N, D, H, W = dims
    with _on(dx.device) as st:
        native.call("fcwdm_dwt3d_cl_bwd", _ptr(dlll), dlll.stride(0), _ptr(dhi), dhi.stride(-2) if dhi is not None else 0,
                    (hi_sb if hi_sb is not None else (dhi.stride(0) if dhi is not None else 0)), _ptr(acc),
                    acc.stride(0) if acc is not None else 0, _ptr(dx), dx.stride(0), N, D, H, W, C, float(lll_scale),
                    float(hi_scale), st)


def idwt3d_cl_bwd(dy, dims, C, dlll, dhi, lll_acc=None, hi_accumulate=False, lll_scale=3.0):
    """Adjoint of idwt3d_cl.  dims = (N, D, H, W) of dy (the IDWT's output)."""
    N, D, H, W = dims
    with _on(dy.device) as st:
        native.call("fcwdm_idwt3d_cl_bwd", _ptr(dy), dy.stride(0), _ptr(lll_acc),
                    lll_acc.stride(0) if lll_acc is not None else 0, _ptr(dlll), dlll.stride(0) if dlll is not None else 0,
                    _ptr(dhi), dhi.stride(-2) if dhi is not None else 0, dhi.stride(0) if dhi is not None else 0,
                    1 if hi_accumulate else 0, N, D, H, W, C, float(lll_scale), st)


def add_cl(a, b, y, rows, C):
    with _on(a.device) as st:
        native.call("fcwdm_add_cl", _ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(y), y.stride(0), rows, C, st)


def linear_bwd(x, W, dy, dx=None, dW=None, db=None, act_in=0, accumulate_dx=False):
    N, K = x.shape
    M = dy.shape[1]
    with _on(x.device) as st:
        native.call("fcwdm_linear_bwd", _ptr(x), _ptr(W), _ptr(dy), dy.stride(0), _ptr(dx), _ptr(dW), _ptr(db), N, K, M,
                    act_in, 1 if accumulate_dx else 0, st)


def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """One fused AdamW step over flat fp32 tensors (in place on p, m, v)."""
    _need_cuda(p, "adamw")
    with _on(p.device) as st:
        native.call("fcwdm_adamw", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(beta1), float(beta2),
                    float(eps), float(weight_decay), int(step), float(grad_scale), st)


# ----------------------------------------------------------------------------------------------------
# plain UNetModel resampling (csrc/resample.cu)
# ----------------------------------------------------------------------------------------------------
def avgpool2_cl(x, dims, C, y, pool_depth=True):
    """x (N,D,H,W,C) cl bf16 -> y pooled by 2 in H, W (and D when pool_depth)."""
    N, D, H, W = dims
    with _on(x.device) as st:
        native.call("fcwdm_avgpool2_cl", _ptr(x), x.stride(0), _ptr(y), y.stride(0), N, D, H, W, C, 1 if pool_depth else 0, st)


def upsample2_cl(x, dims, C, y, up_depth=True):
    """nearest-neighbour x2; dims = (N, D, H, W) of the INPUT."""
    N, D, H, W = dims
    with _on(x.device) as st:
        native.call("fcwdm_upsample2_cl", _ptr(x), x.stride(0), _ptr(y), y.stride(0), N, D, H, W, C, 1 if up_depth else 0, st)


def avgpool2_cl_bwd(dy, dims, C, dx, acc=None, pool_depth=True):
    """Adjoint of avgpool2_cl; dims = (N, D, H, W) of dy (the pooled tensor)."""
    N, D, H, W = dims
    with _on(dy.device) as st:
        native.call("fcwdm_avgpool2_cl_bwd", _ptr(dy), dy.stride(0), _ptr(acc), acc.stride(0) if acc is not None else 0,
                    _ptr(dx), dx.stride(0), N, D, H, W, C, 1 if pool_depth else 0, st)


def upsample2_cl_bwd(dy, dims, C, dx, acc=None, up_depth=True):
    """Adjoint of upsample2_cl; dims = (N, D, H, W) of dy (the up-sampled tensor)."""
    N, D, H, W = dims
    with _on(dy.device) as st:
        native.call("fcwdm_upsample2_cl_bwd", _ptr(dy), dy.stride(0), _ptr(acc), acc.stride(0) if acc is not None else 0,
                    _ptr(dx), dx.stride(0), N, D, H, W, C, 1 if up_depth else 0, st)
