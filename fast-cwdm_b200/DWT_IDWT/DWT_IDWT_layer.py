"""Drop-in for the reference's ``DWT_IDWT.DWT_IDWT_layer`` (3-D operators only; the 1-D/2-D classes of the
reference are imported by nothing on the hot path).

Same constructor and call signatures as the reference modules (DWT_IDWT/DWT_IDWT_layer.py:432-531 and
:534-646): ``DWT_3D(wavename)(x) -> (LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH)`` and
``IDWT_3D(wavename)(LLL, ..., HHH) -> x``; no parameters, no buffers, differentiable.  The work is one
hand-written sm_100a butterfly kernel per call (fcwdm_dwt3d_fwd / fcwdm_idwt3d_fwd) instead of 14 matmuls
against band matrices that the reference rebuilds and uploads on every forward (:459-518).

Differences, all supersets of the reference's behaviour:
  * any even D, H, W (the reference sizes its matrices from max(H, W) and fails when D is larger, :465);
  * runs on the input's device (the reference uses the current device, :505-511);
  * bf16 inputs are accepted (fp32 math); the reference raises on anything but fp32.
CPU tensors raise ``FcwdmError``: there is no CPU fallback.
"""
from torch.nn import Module

from fcwdm.ops import DWT3DFunction, IDWT3DFunction

__all__ = ['DWT_3D', 'IDWT_3D']

_SUPPORTED = ("haar", "db1")   # pywt: 'db1' is the same filter bank as 'haar'


def _check_wavelet(wavename):
    if wavename not in _SUPPORTED:
        raise NotImplementedError(
            f"wavelet {wavename!r}: only the Haar filter bank is implemented (every call site of the reference "
            "uses 'haar': wunet.py:59,106,140; gaussian_diffusion.py:26-27; sample.py:48-49; train_util.py:91-92)")


class DWT_3D(Module):
    """input (N, C, D, H, W) -> 8 tensors (N, C, D/2, H/2, W/2); band letters are the (D, H, W) filters."""

    def __init__(self, wavename):
        super(DWT_3D, self).__init__()
        _check_wavelet(wavename)
        self.wavename = wavename
        # attributes the reference exposes (pywt 'haar' rec_lo / rec_hi, DWT_IDWT_layer.py:451-457)
        s = 0.7071067811865476
        self.band_low = [s, s]
        self.band_high = [s, -s]
        self.band_length = 2
        self.band_length_half = 1

    def forward(self, input):
        assert len(input.size()) == 5
        self.input_depth = input.size()[-3]
        self.input_height = input.size()[-2]
        self.input_width = input.size()[-1]
        return DWT3DFunction.apply(input)


class IDWT_3D(Module):
    """8 tensors (N, C, d, h, w) -> (N, C, 2d, 2h, 2w)."""

    def __init__(self, wavename):
        super(IDWT_3D, self).__init__()
        _check_wavelet(wavename)
        self.wavename = wavename
        s = 0.7071067811865476
        self.band_low = [s, s]      # reversed dec_lo  (DWT_IDWT_layer.py:554-557)
        self.band_high = [s, -s]    # reversed dec_hi
        self.band_length = 2
        self.band_length_half = 1

    def forward(self, LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH):
        assert len(LLL.size()) == len(LLH.size()) == len(LHL.size()) == len(LHH.size()) == 5
        assert len(HLL.size()) == len(HLH.size()) == len(HHL.size()) == len(HHH.size()) == 5
        self.input_depth = LLL.size()[-3] + HHH.size()[-3]
        self.input_height = LLL.size()[-2] + HHH.size()[-2]
        self.input_width = LLL.size()[-1] + HHH.size()[-1]
        return IDWT3DFunction.apply(LLL, LLH, LHL, LHH, HLL, HLH, HHL, HHH)
