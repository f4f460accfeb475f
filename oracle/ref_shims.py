"""TEST INFRASTRUCTURE ONLY -- import harness for the *unmodified* reference.

The reference (tsereda/fast-cwdm, mounted read-only at /root/reference in the build container) is pure
Python but imports four third-party modules that are not installed here: ``pywt`` (only the Haar filter
taps are used, DWT_IDWT/DWT_IDWT_layer.py:451-453,553-557), ``blobfile`` (dist_util.py:9),
``matplotlib.pyplot`` (gaussian_diffusion.py:21) and ``nibabel`` (bratsloader.py:7).  This module injects
stub modules for those names, then puts the reference root first on ``sys.path`` so that
``import DWT_IDWT.DWT_IDWT_layer`` / ``import guided_diffusion.*`` resolve to the reference's own files.

It is used by ``oracle/make_golden.py`` (fixture generation) and by CPU tests that cross-check the oracle
restatement against the real reference when ``/root/reference`` exists.  It never runs on the GPU box
(the reference does not travel) and nothing in the product path may import it.
"""
import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FCWDM_REFERENCE_ROOT", "/root/reference")

# pywavelets 1.4.1 (environment.yml:11) values for pywt.Wavelet('haar')
_S = 0.7071067811865476
_HAAR = dict(dec_lo=[_S, _S], dec_hi=[-_S, _S], rec_lo=[_S, _S], rec_hi=[_S, -_S])


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "guided_diffusion"))


def _install_stubs():
    if "pywt" not in sys.modules:
        pywt = types.ModuleType("pywt")

        class Wavelet:  # only what DWT_IDWT_layer.py touches
            def __init__(self, name):
                if name != "haar":
                    raise ValueError("shim only provides the 'haar' wavelet")
                for k, v in _HAAR.items():
                    setattr(self, k, list(v))

        pywt.Wavelet = Wavelet
        sys.modules["pywt"] = pywt
    for name in ("blobfile", "nibabel"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    try:
        importlib.import_module("matplotlib.pyplot")
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


_REF_PACKAGES = ("DWT_IDWT", "guided_diffusion")


@contextlib.contextmanager
def reference_modules():
    """Context manager: inside it, ``DWT_IDWT`` and ``guided_diffusion`` are the REFERENCE's packages.

    On exit the reference modules are removed from ``sys.modules`` again (and whatever was there before
    is restored) so the product's same-named drop-in packages can be imported in the same process.
    """
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_PACKAGES}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    os.environ.setdefault("WANDB_MODE", "disabled")
    import warnings
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = types.SimpleNamespace()
            ns.layer = importlib.import_module("DWT_IDWT.DWT_IDWT_layer")
            ns.wunet = importlib.import_module("guided_diffusion.wunet")
            ns.gd = importlib.import_module("guided_diffusion.gaussian_diffusion")
            ns.respace = importlib.import_module("guided_diffusion.respace")
            ns.script_util = importlib.import_module("guided_diffusion.script_util")
            ns.nn = importlib.import_module("guided_diffusion.nn")
        yield ns
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_PACKAGES]:
            del sys.modules[k]
        sys.modules.update(saved)
