"""Drop-in for the reference's ``guided_diffusion`` package, hot path only.

Own modules: ``nn``, ``wunet``, ``unet``, ``gaussian_diffusion``, ``respace``, ``script_util`` (the models, the
diffusion and the factories the entry scripts import) and, for the rows SURVEY.md section 8f ranks next, the training
driver and its helpers: ``train_util``, ``dist_util``, ``resample``, ``logger``, ``bratsloader``.  Anything else the
reference's scripts import from this package (``losses``, ``lidcloader``, ...) is resolved from an unmodified checkout
of the reference when one is available: set ``FCWDM_REFERENCE_ROOT`` (default ``/root/reference``) and those files are
found through this package's ``__path__`` *after* the modules here, so ``scripts/sample.py`` / ``scripts/train.py``
run unchanged with ``PYTHONPATH=<repo>/fast-cwdm_b200`` while the hot path goes through the B200 kernels.
"""
import os as _os

_ref = _os.path.join(_os.environ.get("FCWDM_REFERENCE_ROOT", "/root/reference"), "guided_diffusion")
if _os.path.isdir(_ref) and _ref not in __path__:
    __path__.append(_ref)
