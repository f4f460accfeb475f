"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the fast-cwdm hot path: 3-D Haar DWT/IDWT, the WavUNetModel denoiser forward,
and the GaussianDiffusion sampling / training arithmetic with respace.py timestep spacing.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this package, and only as the checker or the timed CPU baseline -- never as part of the product path
(`fast-cwdm_b200/`), which must fail loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the *reference itself*, executed in the build container through ``oracle/ref_shims.py``;
``oracle/make_golden.py`` is the committed generator and ``tests/golden/*.npz`` the frozen fixtures.
"""
