"""TEST INFRASTRUCTURE ONLY -- import harness for the *unmodified* reference.

The reference (tsereda/fast-cwdm, mounted read-only at /root/reference in the build container) is pure
Python but imports four third-party modules that are not installed here: ``pywt`` (only the Haar filter
taps are used, DWT_IDWT/DWT_IDWT_layer.py:451-453,553-557), ``blobfile`` (dist_util.py:9),
``matplotlib.pyplot`` (gaussian_diffusion.py:21) and ``nibabel`` (bratsloader.py:7).  This module injects
stub modules for those names, then puts the reference root first on ``sys.path`` so that
``import DWT_IDWT.DWT_IDWT_layer`` / ``import guided_diffusion.*`` resolve to the reference's own files.

It is used by ``oracle/make_golden*.py`` (fixture generation), by CPU tests that cross-check the oracle
restatement against the real reference, and by ``bench.py --impl reference`` (the CPU arm), which on the GPU box
finds the byte-for-byte staged copy ``oracle/_ref/`` made by ``stage_reference()`` at build time.  Nothing in the
product path may import it.
"""
import contextlib
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")          # git-ignored copy made by stage_reference() (travels to the GPU box)
_PACKAGES = ("DWT_IDWT", "guided_diffusion")


def _pick_root():
    """$FCWDM_REFERENCE_ROOT, else the read-only mount of the build container, else the staged copy."""
    env = os.environ.get("FCWDM_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", STAGED_ROOT):
        if os.path.isdir(os.path.join(root, "guided_diffusion")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def stage_reference(src="/root/reference"):
    """Copy the two Python packages of the UNMODIFIED reference that the hot path lives in into oracle/_ref/ (git-ignored,
    not gpurun-ignored), so that bench.py --impl reference can time the reference's own code on the GPU box's host
    cores.  Called by __graft_entry__.build() in the build container; a no-op where the reference is not mounted.
    Nothing is edited: files are byte-for-byte copies, verified by size + sha256 in _ref/MANIFEST.json."""
    import hashlib
    import json
    import shutil
    if not os.path.isdir(os.path.join(src, "guided_diffusion")):
        return None
    manifest = {}
    for pkg in _PACKAGES:
        dst_dir = os.path.join(STAGED_ROOT, pkg)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(os.path.join(src, pkg))):
            if not name.endswith(".py"):
                continue
            s, d = os.path.join(src, pkg, name), os.path.join(dst_dir, name)
            data = open(s, "rb").read()
            if not os.path.exists(d) or open(d, "rb").read() != data:
                shutil.copyfile(s, d)
            manifest[f"{pkg}/{name}"] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    # the two entry scripts the drop-in claims to serve unchanged (tests/test_reference_scripts_gpu.py executes them)
    os.makedirs(os.path.join(STAGED_ROOT, "scripts"), exist_ok=True)
    for name in ("sample.py", "train.py"):
        s = os.path.join(src, "scripts", name)
        if os.path.exists(s):
            data = open(s, "rb").read()
            d = os.path.join(STAGED_ROOT, "scripts", name)
            if not os.path.exists(d) or open(d, "rb").read() != data:
                shutil.copyfile(s, d)
            manifest[f"scripts/{name}"] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    with open(os.path.join(STAGED_ROOT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    return STAGED_ROOT

# pywavelets 1.4.1 (environment.yml:11) values for pywt.Wavelet('haar')
_S = 0.7071067811865476
_HAAR = dict(dec_lo=[_S, _S], dec_hi=[-_S, _S], rec_lo=[_S, _S], rec_hi=[_S, -_S])


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "guided_diffusion"))


def reference_script(name):
    """Path of the reference's unmodified scripts/<name> (mount or staged copy), or None."""
    for root in (REFERENCE_ROOT, STAGED_ROOT):
        p = os.path.join(root, "scripts", name)
        if os.path.exists(p):
            return p
    return None


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
