"""Drop-in boundary: every public call signature equals the reference's.

tests/golden/signatures.json is ``inspect.signature`` of each boundary symbol of the UNMODIFIED reference
(oracle/make_golden_signatures.py).  The drop-in's signature must be identical, or start with the reference's
parameters (same names, order, defaults) and only ADD keyword parameters with defaults; the deliberate differences
are whitelisted below with the reason."""
import inspect
import json
import os

import pytest

from oracle import make_golden_signatures as mgs
from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "signatures.json")

# symbol -> (what differs, why)
WHITELIST = {
    "guided_diffusion.gaussian_diffusion:GaussianDiffusion.p_sample_loop_progressive":
        ("time=None instead of time=1000",
         "None = num_timesteps: identical for T = 1000, and p_sample_loop (which never passes `time`) no longer raises "
         "IndexError for every other T (SURVEY.md fact 4)"),
    "guided_diffusion.resample:UniformSampler.__init__":
        ("maxt=None instead of a required maxt", "every reference call site still binds it; None = diffusion.num_timesteps"),
    "guided_diffusion.train_util:TrainLoop.run_step":
        ("info=None instead of the mutable default info={}", "same behaviour without sharing one dict between calls"),
}


def _params(sig_text):
    """'(self, a, b=1, *, c)' -> list of 'name[=default]' tokens with kind markers kept."""
    body = sig_text.strip()
    body = body[body.index("(") + 1:body.rindex(")")]
    out, depth, cur = [], 0, ""
    for ch in body:
        if ch in "([{":
            depth += 1
        if ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _compatible(ref, ours):
    """Identical, or the reference's parameter list is a prefix and every extra parameter has a default."""
    if ref == ours:
        return True
    r, o = _params(ref), _params(ours)
    if o[:len(r)] != r:
        return False
    return all("=" in extra or extra.startswith("**") for extra in o[len(r):])


def test_drop_in_signatures_match_the_reference_fixture():
    ref = json.load(open(GOLDEN))
    ours = mgs.signatures()
    assert set(ours) == set(ref)
    bad = []
    for name, rsig in sorted(ref.items()):
        osig = ours[name]
        if _compatible(rsig, osig):
            continue
        if name in WHITELIST:
            continue
        bad.append(f"{name}\n   reference: {rsig}\n   drop-in:   {osig}")
    assert not bad, "signature differences not whitelisted:\n" + "\n".join(bad)
    for name in WHITELIST:                      # a whitelist entry that no longer differs must be removed
        assert not _compatible(ref[name], ours[name]), f"{name} is whitelisted but no longer differs"


def test_positional_order_of_create_model():
    """positional callers of create_model must bind num_groups before dims (script_util.py:207-208)."""
    from guided_diffusion.script_util import create_model
    names = list(inspect.signature(create_model).parameters)
    assert names.index("num_groups") + 1 == names.index("dims")


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not mounted (GPU box)")
def test_fixture_is_the_live_reference():
    assert mgs.reference_signatures() == json.load(open(GOLDEN))


def test_guided_diffusion_path_is_closed_by_default(monkeypatch):
    """The product package resolves nothing from a reference checkout unless FCWDM_REFERENCE_ROOT is set."""
    import importlib
    import guided_diffusion
    monkeypatch.delenv("FCWDM_REFERENCE_ROOT", raising=False)
    mod = importlib.reload(guided_diffusion)
    assert len(mod.__path__) == 1 and mod.__path__[0].endswith(os.path.join("fast-cwdm_b200", "guided_diffusion"))
