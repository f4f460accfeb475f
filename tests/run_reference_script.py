"""TEST INFRASTRUCTURE: run one of the reference's UNMODIFIED entry scripts (scripts/sample.py) in this process with
  * `nibabel` provided by tests/nibabel_shim.py (not installed in this image),
  * the drop-in packages first on sys.path (what PYTHONPATH=<repo>/fast-cwdm_b200 does for a user),
  * the fused sampler's per-step noise drawn from per-(case, step) seeded generators and every case's x_T recorded, so
    the caller can replay the exact chain through the oracle.

    python tests/run_reference_script.py <script.py> <record_dir> [script args ...]"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fast-cwdm_b200"), os.path.join(ROOT, "tests")]

import torch  # noqa: E402

import nibabel_shim  # noqa: E402

script, record = sys.argv[1], sys.argv[2]
sys.modules["nibabel"] = nibabel_shim.as_module()

from fcwdm.sampler import FusedSampler  # noqa: E402

NOISE_SEED = 7000
case = [-1]
_begin = FusedSampler.begin


def begin(self, noise, cond, want_pred=True):
    case[0] += 1
    torch.save(noise.detach().cpu(), os.path.join(record, f"x_T_{case[0]}.pt"))
    return _begin(self, noise, cond, want_pred=want_pred)


def hook(buf, i):
    g = torch.Generator(device=buf.device).manual_seed(NOISE_SEED + 100 * case[0] + int(i))
    buf.normal_(generator=g)


FusedSampler.begin = begin
FusedSampler.noise_hook = staticmethod(hook)
sys.argv = [script] + sys.argv[3:]
os.chdir(os.path.dirname(os.path.dirname(script)))       # the scripts do sys.path.append(".") from the checkout root
runpy.run_path(script, run_name="__main__")
print(f"__SCRIPT_OK__ cases={case[0] + 1}")
