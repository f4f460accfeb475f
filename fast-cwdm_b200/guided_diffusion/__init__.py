"""Drop-in for the reference's ``guided_diffusion`` package, hot path only.

Own modules: ``nn``, ``wunet``, ``unet``, ``gaussian_diffusion``, ``respace``, ``script_util`` (the models, the
diffusion and the factories the entry scripts import) and, for the rows SURVEY.md section 8f ranks next, the training
driver and its helpers: ``train_util``, ``dist_util``, ``resample``, ``logger``, ``bratsloader``.  Anything else the
reference's scripts import from this package (``losses``, ``lidcloader``, ...) is NOT part of this package.  Opt-in
only: when ``FCWDM_REFERENCE_ROOT`` is set to an unmodified checkout of the reference, its ``guided_diffusion``
directory is appended to this package's ``__path__`` so those out-of-scope modules are found *after* the modules here
(``scripts/sample.py`` / ``scripts/train.py`` themselves need none of them).  Without the variable nothing outside this
directory is ever imported under the ``guided_diffusion`` name.
"""
import os as _os

_root = _os.environ.get("FCWDM_REFERENCE_ROOT")
if _root:
    _ref = _os.path.join(_root, "guided_diffusion")
    if _os.path.isdir(_ref) and _ref not in __path__:
        __path__.append(_ref)
