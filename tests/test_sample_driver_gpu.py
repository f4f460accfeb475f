"""fcwdm.sample_driver.SamplingDriver end to end on the GPU: NIfTI cases on disk -> reader threads
-> VolumeStream(raw=True) -> writer threads -> NIfTI results, for the two output conventions of the reference
(scripts/sample.py and scripts/sample_auto.py), sharded and unsharded, against the synchronous per-case call."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

AFFINE = np.array([[-1.0, 0, 0, 90.0], [0, -1.0, 0, 126.0], [0, 0, 1.0, -72.0], [0, 0, 0, 1.0]])


def make_cases(root, n, drop=None):
    from fcwdm import nifti
    g = np.random.default_rng(11)
    raws = []
    for i in range(n):
        subj = f"BraTS-GLI-{i:05d}-000"
        (root / subj).mkdir(parents=True)
        case = {}
        for m in ("t1n", "t1c", "t2w", "t2f"):
            v = (g.random((240, 240, 155)) * (500.0 + 100 * i)).astype(np.float32)
            v[:30] = 0                                       # background: exercises the cond_1 == 0 mask
            v[:, :25] = 0
            case[m] = v
            if m != drop:
                nifti.write(root / subj / f"{subj}-{m}.nii.gz", v, affine=AFFINE)
        raws.append(case)
    return raws


@pytest.fixture(scope="module")
def model_and_diffusion():
    import bench
    model, diffusion = bench.build_model(torch.device("cuda"))
    return model, diffusion


def direct(diffusion, model, raw_case, contr, index, seed, post="sample"):
    """The synchronous per-case path the driver pipelines: GPU preprocessing, synthesize."""
    from fcwdm import pipeline, preprocess
    from fcwdm.sample_driver import conditions_for
    conds = conditions_for(contr)
    stack = torch.stack([torch.from_numpy(raw_case[m]) for m in conds]).cuda()
    v = preprocess.clip_and_normalize(stack)                           # (3, 1, 224, 224, 160)
    noise = torch.randn((1, 8, 112, 112, 80), generator=torch.Generator().manual_seed(seed + index)).cuda()
    torch.cuda.manual_seed(seed + index)
    img = pipeline.synthesize(diffusion, model, v[0:1], v[1:2], v[2:3], noise, post=post)
    return img[0].cpu().numpy()


def close(a, b):
    # same kernels and inputs; only the GroupNorm statistics' atomic summation order differs run to run
    return float(np.abs(a - b).max()) <= 5e-2 and float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-6)) <= 1e-2


def test_sample_mode_sharded_and_unsharded(tmp_path, model_and_diffusion):
    from fcwdm import nifti, preprocess
    from fcwdm.sample_driver import SamplingDriver
    from guided_diffusion.bratsloader import BRATSVolumes
    model, diffusion = model_and_diffusion
    raws = make_cases(tmp_path / "validation", 3)
    ds = BRATSVolumes(str(tmp_path / "validation"), mode="eval", raw=True)
    assert len(ds) == 3
    out1 = tmp_path / "out_1rank"
    drv = SamplingDriver(diffusion, model, ds.database, output_dir=str(out1), mode="sample", contr="t1n", seed=3)
    stats = drv.run()
    drv.close()
    assert stats["cases"] == 3 and not stats["skipped"] and stats["bytes_written"] > 0
    print("driver stats", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in stats.items()})
    for i, raw in enumerate(raws):
        subj = f"BraTS-GLI-{i:05d}-000"
        got, h = nifti.read(out1 / subj / "sample.nii.gz", dtype=np.float32, return_header=True)
        assert got.shape == (224, 224, 155) and h.datatype == 16 and np.array_equal(h.affine, np.eye(4))
        assert got.min() >= 0.0 and got.max() <= 1.0
        assert float(np.abs(got[:22]).max()) == 0.0                     # cond_1 (t1c) background -> 0 (sample.py:125)
        want = direct(diffusion, model, raw, "t1n", i, seed=3)
        assert close(got, want), i
        tgt = nifti.read(out1 / subj / "target.nii.gz", dtype=np.float32)
        ref_t = preprocess.clip_and_normalize(torch.from_numpy(raw["t1n"])[None].cuda())[0, 0, :, :, :155].cpu().numpy()
        assert np.array_equal(tgt, ref_t)
    # two "ranks" one after the other on this GPU: disjoint shards, same files as the unsharded run
    out2 = tmp_path / "out_2rank"
    seen = 0
    for rank in range(2):
        d2 = SamplingDriver(diffusion, model, ds.database, output_dir=str(out2), mode="sample", contr="t1n", seed=3,
                            rank=rank, world_size=2, write_target=False)
        seen += d2.run()["cases"]
        d2.close()
    assert seen == 3
    for i in range(3):
        subj = f"BraTS-GLI-{i:05d}-000"
        a = nifti.read(out1 / subj / "sample.nii.gz", dtype=np.float32)
        b = nifti.read(out2 / subj / "sample.nii.gz", dtype=np.float32)
        assert close(b, a), i
        assert not (out2 / subj / "target.nii.gz").exists()


def test_auto_mode_fills_the_missing_modality(tmp_path, model_and_diffusion):
    from fcwdm import nifti
    from fcwdm.sample_driver import SamplingDriver
    from guided_diffusion.bratsloader import BRATSVolumes
    model, diffusion = model_and_diffusion
    raws = make_cases(tmp_path / "pseudo", 2, drop="t2w")
    ds = BRATSVolumes(str(tmp_path / "pseudo"), mode="auto", raw=True)
    assert ds[0]["missing"] == "t2w"
    # a case with two modalities missing is reported and skipped, not fatal
    broken = dict(ds.database[1])
    del broken["t2f"]
    drv = SamplingDriver(diffusion, {"t2w": model}, list(ds.database) + [broken], mode="auto", seed=9)
    stats = drv.run()
    drv.close()
    assert stats["cases"] == 2 and len(stats["skipped"]) == 1 and stats["skipped"][0][0] == 2
    for i, raw in enumerate(raws):
        subj = f"BraTS-GLI-{i:05d}-000"
        path = tmp_path / "pseudo" / subj / f"{subj}-t2w.nii.gz"        # next to the inputs (sample_auto.py:77)
        got, h = nifti.read(path, dtype=np.float32, return_header=True)
        assert got.shape == (240, 240, 155)
        assert np.array_equal(h.affine, AFFINE)                          # header of a present modality (t1n)
        assert float(np.abs(got[:8]).max()) == 0.0 and float(np.abs(got[:, -8:]).max()) == 0.0   # the padding
        inner = got[8:-8, 8:-8]
        assert not ((inner > 0) & (inner <= 0.04)).any()                 # sample_auto.py:137
        want = direct(diffusion, model, raw, "t2w", i, seed=9, post="auto")
        assert close(inner, want), i
    assert len(BRATSVolumes(str(tmp_path / "pseudo"), mode="auto")) == 2 and \
        "t2w" in BRATSVolumes(str(tmp_path / "pseudo"), mode="auto").database[0]
