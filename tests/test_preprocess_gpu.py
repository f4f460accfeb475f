"""GPU parity of fcwdm_clip_normalize (through the C-ABI) with the reference loader's clip_and_normalize + pad + crop
(guided_diffusion/bratsloader.py:44-50,107-111).  Quantiles are exact order statistics (radix select) interpolated in
float64 like np.quantile; the normalisation runs in fp32: |err| <= 2e-7 on values in [0, 1]."""
import numpy as np
import pytest
import torch

from oracle import preprocess as op

pytestmark = pytest.mark.gpu


def test_fixtures(golden):
    from fcwdm import preprocess
    g = golden("preprocess")
    for k, crop, pad in (("vol", 4, 32), ("neg", 2, 9)):
        raw = torch.from_numpy(g[k]).cuda()
        out, q = preprocess.clip_and_normalize(raw[None], crop=crop, pad_to=pad, return_quantiles=True)
        torch.cuda.synchronize()
        np.testing.assert_allclose(q.cpu().numpy()[0], g[k + "_q"], rtol=1e-7)
        X, Y, Z = g[k].shape
        ref = op.pad_crop(g[k + "_out"], crop, pad)
        assert tuple(out.shape) == (1,) + ref.shape
        np.testing.assert_allclose(out.cpu().numpy()[0], ref, rtol=0, atol=2e-7)
        assert float(out[..., Z:].abs().max()) == 0.0 if pad > Z else True


def test_full_size_case_against_oracle():
    """Four raw BraTS-shaped modalities (240 x 240 x 155), each normalised on its own -> 4 x (1, 224, 224, 160)."""
    from fcwdm import preprocess
    rng = np.random.default_rng(3)
    raw = rng.gamma(2.0, 150.0, size=(4, 240, 240, 155)).astype(np.float32)
    raw[:, :20] = 0
    raw[:, :, -30:] = 0
    raw[1] *= 0.01
    raw[2] = np.round(raw[2])                                        # integer-valued intensities: many ties
    case = preprocess.preprocess_case(*[torch.from_numpy(raw[i]).cuda() for i in range(4)])
    torch.cuda.synchronize()
    assert case["missing"] == "none"
    for i, k in enumerate(("t1n", "t1c", "t2w", "t2f")):
        ref = op.preprocess_volume(raw[i])
        got = case[k].cpu().numpy()
        assert got.shape == (1, 224, 224, 160) and got.dtype == np.float32
        assert float(np.abs(got - ref).max()) <= 2e-7
        assert 0.0 <= got.min() and got.max() <= 1.0
    one = preprocess.preprocess_case(t1c=torch.from_numpy(raw[1]).cuda(), t2w=torch.from_numpy(raw[2]).cuda(),
                                     t2f=torch.from_numpy(raw[3]).cuda())
    assert one["missing"] == "t1n" and one["t1n"].shape == (1,)
    assert torch.equal(one["t2f"], case["t2f"])


def test_errors():
    from fcwdm import FcwdmError, preprocess
    with pytest.raises(FcwdmError):
        preprocess.clip_and_normalize(torch.zeros(1, 8, 8, 8))          # CPU tensor: no fallback
    with pytest.raises(ValueError):
        preprocess.clip_and_normalize(torch.zeros(8, 8, 8, device="cuda"))
