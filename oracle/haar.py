"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's 3-D Haar DWT / IDWT.

Follows the *order of operations* of the reference so that fp32 rounding matches as closely as a
non-BLAS restatement can:

* analysis  = DWTFunction_3D.forward  (DWT_IDWT/DWT_IDWT_Functions.py:117-136): filter along H (dim -2),
  then along W (dim -1), then along D (dim -3); the band matrices built by DWT_3D.get_matrix
  (DWT_IDWT/DWT_IDWT_layer.py:459-518) hold the pywt 'haar' reconstruction taps rec_lo=[s,s],
  rec_hi=[s,-s] on consecutive pairs, i.e. low[i] = s*x[2i] + s*x[2i+1], high[i] = s*x[2i] - s*x[2i+1].
* synthesis = IDWTFunction_3D.forward (DWT_IDWT_Functions.py:161-181): D first, then W, then H, with the
  transposed matrices (IDWT_3D uses reversed dec_lo/dec_hi, DWT_IDWT_layer.py:554-557, which for Haar
  gives the same matrices): x[2i] = s*lo[i] + s*hi[i], x[2i+1] = s*lo[i] - s*hi[i].

Band naming: letters are the filters on (D, H, W) in that order, L=low, H=high; the tuple order is
(LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH)  (DWT_IDWT_Functions.py:128-136).
"""
import numpy as np

S32 = np.float32(0.7071067811865476)  # torch.Tensor(matrix) casts the float64 taps to fp32 (layer.py:506-518)

BAND_NAMES = ("LLL", "LLH", "LHL", "LHH", "HLL", "HLH", "HHL", "HHH")


def _analysis(x, axis):
    x = np.moveaxis(x, axis, -1)
    a, b = x[..., 0::2], x[..., 1::2]
    s = x.dtype.type(S32)
    lo = s * a + s * b
    hi = s * a - s * b
    return np.moveaxis(lo, -1, axis), np.moveaxis(hi, -1, axis)


def _synthesis(lo, hi, axis):
    lo = np.moveaxis(lo, axis, -1)
    hi = np.moveaxis(hi, axis, -1)
    s = lo.dtype.type(S32)
    out = np.empty(lo.shape[:-1] + (2 * lo.shape[-1],), dtype=lo.dtype)
    out[..., 0::2] = s * lo + s * hi
    out[..., 1::2] = s * lo - s * hi
    return np.moveaxis(out, -1, axis)


def dwt3d(x):
    """x: (N, C, D, H, W) array with even D, H, W -> tuple of 8 arrays (N, C, D/2, H/2, W/2).

    The reference additionally requires D <= max(H, W) (layer.py:465 sizes the matrices from H and W
    only); the restatement, like the product kernel, accepts any even dims.
    """
    x = np.asarray(x)
    assert x.ndim == 5, "DWT_3D asserts a 5-D input (DWT_IDWT_layer.py:525)"
    assert x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and x.shape[4] % 2 == 0
    L, H = _analysis(x, 3)                      # Functions.py:122-123 (matrix_Low_0 / High_0 act on H)
    LL, LH = _analysis(L, 4)                    # :124-125 (W)
    HL, HH = _analysis(H, 4)                    # :126-127
    LLL, HLL = _analysis(LL, 2)                 # :128,:132 (D): first letter = D filter
    LLH, HLH = _analysis(LH, 2)                 # :129,:133
    LHL, HHL = _analysis(HL, 2)                 # :130,:134
    LHH, HHH = _analysis(HH, 2)                 # :131,:135
    return LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH


def idwt3d(LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH):
    """Inverse of :func:`dwt3d`; 8 arrays (N, C, d, h, w) -> (N, C, 2d, 2h, 2w)."""
    bands = [np.asarray(b) for b in (LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH)]
    assert all(b.ndim == 5 for b in bands), "IDWT_3D asserts 5-D inputs (DWT_IDWT_layer.py:636-639)"
    LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH = bands
    LL = _synthesis(LLL, HLL, 2)                # Functions.py:167-168 (D)
    LH = _synthesis(LLH, HLH, 2)                # :169-170
    HL = _synthesis(LHL, HHL, 2)                # :171-172
    HH = _synthesis(LHH, HHH, 2)                # :173-174
    L = _synthesis(LL, LH, 4)                   # :175-176 (W)
    H = _synthesis(HL, HH, 4)                   # :177-178
    return _synthesis(L, H, 3)                  # :179-180 (H)


def dwt3d_butterfly_f64(x):
    """Closed form used as a second, independent check:

    band[fd,fh,fw][d,h,w] = 1/(2*sqrt(2)) * sum_{i,j,k in {0,1}} (-1)^(fd*i+fh*j+fw*k) x[2d+i,2h+j,2w+k]
    evaluated in float64.
    """
    x = np.asarray(x, dtype=np.float64)
    out = []
    c = 1.0 / (2.0 * np.sqrt(2.0))
    for fd in (0, 1):
        for fh in (0, 1):
            for fw in (0, 1):
                acc = np.zeros(x.shape[:2] + (x.shape[2] // 2, x.shape[3] // 2, x.shape[4] // 2))
                for i in (0, 1):
                    for j in (0, 1):
                        for k in (0, 1):
                            sign = (-1.0) ** (fd * i + fh * j + fw * k)
                            acc += sign * x[:, :, i::2, j::2, k::2]
                out.append(c * acc)
    return tuple(out)
