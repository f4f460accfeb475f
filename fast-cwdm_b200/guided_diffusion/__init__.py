"""Drop-in for the reference's ``guided_diffusion`` package, hot path only.

Own modules: ``nn``, ``wunet``, ``gaussian_diffusion``, ``respace``, ``script_util`` (the factories and flag
plumbing the entry scripts import).  Everything else the reference's scripts import from this package
(``dist_util``, ``logger``, ``bratsloader``, ``train_util``, ``resample``, ``losses``, ``unet`` -- orchestration,
I/O and the sibling model, out of scope per SURVEY.md section 8) is resolved from an unmodified checkout of the
reference when one is available: set ``FCWDM_REFERENCE_ROOT`` (default ``/root/reference``) and those files are
found through this package's ``__path__`` *after* the modules here, so ``scripts/sample.py`` / ``scripts/train.py``
run unchanged with ``PYTHONPATH=<repo>/fast-cwdm_b200`` while the hot path goes through the B200 kernels.
"""
import os as _os

_ref = _os.path.join(_os.environ.get("FCWDM_REFERENCE_ROOT", "/root/reference"), "guided_diffusion")
if _os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)
