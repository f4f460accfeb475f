"""Host side of the training driver: timestep samplers, checkpoint-name parsing, the logger and
the launcher-aware dist_util, checked on the CPU -- against the fixture the unmodified reference produced
(tests/golden/trainloop_host.json, oracle/make_golden_trainloop.py) and with a real 2-process gloo group."""
import csv
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trainloop_host.json")


class _Diff:
    def __init__(self, T):
        self.num_timesteps = T


def test_uniform_sampler_draws_the_reference_timesteps():
    from guided_diffusion.resample import UniformSampler, create_named_schedule_sampler
    with open(GOLDEN) as f:
        cases = json.load(f)["uniform"]
    assert cases
    for case in cases:
        np.random.seed(case["seed"])
        s = create_named_schedule_sampler("uniform", _Diff(case["T"]), case["T"])
        assert isinstance(s, UniformSampler)
        for call in case["calls"]:
            t, w = s.sample(case["B"], "cpu")
            assert t.dtype == torch.int64 and w.dtype == torch.float32
            assert t.tolist() == call["t"]
            assert w.tolist() == call["w"]


def test_unknown_sampler_name_is_refused():
    from guided_diffusion.resample import create_named_schedule_sampler
    with pytest.raises(NotImplementedError):
        create_named_schedule_sampler("nope", _Diff(10), 10)


def test_second_moment_resampler_weights():
    from guided_diffusion.resample import LossSecondMomentResampler
    T, H = 4, 3
    s = LossSecondMomentResampler(_Diff(T), history_per_term=H, uniform_prob=0.01)
    assert np.array_equal(s.weights(), np.ones(T))                       # not warmed up: uniform
    hist = {0: [1.0, 1.0, 1.0], 1: [2.0, 2.0, 2.0], 2: [0.0, 3.0, 0.0], 3: [1.0, 2.0, 3.0]}
    for k in range(H):
        s.update_with_all_losses(list(hist), [hist[t][k] for t in hist])
    rms = np.array([np.sqrt(np.mean(np.square(hist[t]))) for t in range(T)])
    want = rms / rms.sum() * 0.99 + 0.01 / T
    np.testing.assert_allclose(s.weights(), want, rtol=1e-12)
    s.update_with_all_losses([2], [6.0])                                  # full history: oldest value shifts out
    np.testing.assert_allclose(s._loss_history[2], [3.0, 0.0, 6.0])
    np.random.seed(1)
    t, w = s.sample(64, "cpu")
    p = s.weights() / s.weights().sum()
    np.testing.assert_allclose(w.numpy(), 1.0 / (T * p[t.numpy()]), rtol=1e-6)


def test_parse_resume_step_matches_reference():
    from guided_diffusion.train_util import parse_resume_step_from_filename
    with open(GOLDEN) as f:
        cases = json.load(f)["parse"]
    assert len(cases) >= 8
    for name, want in cases.items():
        assert parse_resume_step_from_filename(name) == want, name


def test_logger_mean_dump_and_files(tmp_path):
    from guided_diffusion import logger
    logger.configure(dir=str(tmp_path), format_strs=["log", "csv", "json"])
    try:
        logger.logkv("step", 1)
        logger.logkv_mean("loss", torch.tensor(2.0))
        logger.logkv_mean("loss", 4.0)
        assert logger.getkvs()["loss"] == 3.0
        first = logger.dumpkvs()
        assert first == {"step": 1, "loss": 3.0}
        logger.logkv("step", 2)
        logger.logkv("extra", 7)                                          # a new column widens the csv header
        logger.dumpkvs()
        logger.log("a text line")
        assert logger.get_dir() == str(tmp_path)
    finally:
        logger.reset()
    rows = list(csv.DictReader(open(tmp_path / "progress.csv")))
    assert [r["step"] for r in rows] == ["1", "2"]
    assert rows[0]["loss"] == "3.0" and rows[0]["extra"] == "" and rows[1]["extra"] == "7"
    lines = [json.loads(l) for l in open(tmp_path / "progress.json")]
    assert lines[0]["loss"] == 3.0 and lines[1]["extra"] == 7
    assert "a text line" in open(tmp_path / "log.txt").read()


def test_blob_logdir_override(tmp_path, monkeypatch):
    from guided_diffusion import train_util
    monkeypatch.setenv("FCWDM_CHECKPOINT_ROOT", str(tmp_path))
    assert train_util.get_blob_logdir() == str(tmp_path)
    assert train_util.find_resume_checkpoint() is None


def test_visualize():
    from guided_diffusion.train_util import visualize
    a = np.array([[1.0, 3.0], [2.0, 5.0]])
    np.testing.assert_allclose(visualize(a), (a - 1.0) / 4.0)
    assert np.array_equal(visualize(np.full((2, 2), 3.0)), np.zeros((2, 2)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), CUDA_VISIBLE_DEVICES="")
    from guided_diffusion import dist_util
    from guided_diffusion.resample import LossSecondMomentResampler
    dist_util.setup_dist(devices=[0])
    dist_util.setup_dist(devices=[0])                                     # idempotent
    assert dist.get_backend() == "gloo" and dist.get_world_size() == world and dist_util.get_rank() == rank
    assert dist_util.dev() == torch.device("cpu")
    # sync_params: every rank ends with rank 0's values
    params = [torch.full((5,), float(rank + 1)), torch.full((2, 3), float(10 * (rank + 1)))]
    dist_util.sync_params(params)
    assert torch.equal(params[0], torch.full((5,), 1.0)) and torch.equal(params[1], torch.full((2, 3), 10.0))
    # load_state_dict: only rank 0 can see the file; the others receive its bytes
    path = os.path.join(tmp, "ckpt_rank0.pt")
    if rank == 0:
        torch.save({"w": torch.arange(6.0)}, path)
    dist.barrier()
    sd = dist_util.load_state_dict(path if rank == 0 else "/nonexistent/for/this/rank.pt", map_location="cpu")
    assert torch.equal(sd["w"], torch.arange(6.0))
    # loss-aware sampler: ragged per-rank batches are gathered, all ranks apply the same update
    s = LossSecondMomentResampler(_Diff(4), history_per_term=2)
    ts = torch.tensor([0, 1] if rank == 0 else [2], dtype=torch.int64)
    ls = torch.tensor([1.0, 2.0] if rank == 0 else [3.0])
    s.update_with_local_losses(ts, ls)
    assert s._loss_counts.tolist() == [1, 1, 1, 0]
    assert s._loss_history[:, 0].tolist() == [1.0, 2.0, 3.0, 0.0]
    dist.barrier()
    dist.destroy_process_group()


def test_dist_util_two_rank_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)


def test_setup_dist_without_launcher_is_a_world_of_one(monkeypatch):
    from guided_diffusion import dist_util
    if dist.is_initialized():
        pytest.skip("a process group already exists in this interpreter")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "")
    try:
        dist_util.setup_dist(devices=(0,))
        assert dist.get_world_size() == 1 and dist.get_rank() == 0
        assert os.environ["MASTER_ADDR"] == "127.0.0.1"
        dist_util.sync_params([torch.zeros(3)])                            # no-op
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
        for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            os.environ.pop(k, None)
