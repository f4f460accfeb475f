"""Key/value + text logger with the call surface the reference's training code uses (guided_diffusion/logger.py:212-290,
442-472 of the reference: ``configure``, ``log``/``debug``/``info``/``warn``/``error``, ``logkv``, ``logkv_mean``,
``logkvs``, ``dumpkvs``, ``getkvs``, ``get_dir``, ``set_level``, ``reset``).

Written for one-process-per-GPU runs: only rank 0 (``RANK`` env, as torchrun sets it) writes files; other ranks keep
their key/values in memory (``getkvs`` works everywhere) and print nothing below WARN.  Values handed to ``logkv`` /
``logkv_mean`` may be CUDA tensors: they are kept as tensors and read back once, in ``dumpkvs`` -- logging a value does
not stall the launch queue the way the reference's per-step ``.item()`` calls do (train_util.py:215,371-375).

Outputs (rank 0): ``log.txt`` (human table + text lines), ``progress.csv``, ``progress.json`` (one JSON object per
dump) under ``get_dir()``; stdout gets the table too.  Format selection follows the reference's ``format_strs`` /
``OPENAI_LOG_FORMAT`` convention with the subset {stdout, log, csv, json}.
"""
import datetime
import json
import os
import sys
import tempfile
import time

DEBUG, INFO, WARN, ERROR, DISABLED = 10, 20, 30, 40, 50

_current = None


def _rank():
    for key in ("RANK", "PMI_RANK", "OMPI_COMM_WORLD_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


def _scalar(v):
    if hasattr(v, "detach"):                    # torch tensor (possibly on the GPU): one read-back, here
        v = v.detach()
        return float(v.float().mean().item()) if v.numel() != 1 else float(v.item())
    if hasattr(v, "item") and getattr(v, "size", 2) == 1:
        return v.item()
    return v


class _Logger:
    def __init__(self, directory, formats, rank):
        self.dir = directory
        self.rank = rank
        self.level = INFO
        self.kvs = {}                            # key -> value, or [sum-able list] for means
        self.means = {}
        self.formats = formats if rank == 0 else []
        self.files = {}
        self.csv_keys = []
        if directory and rank == 0:
            os.makedirs(directory, exist_ok=True)
            if "log" in self.formats:
                self.files["log"] = open(os.path.join(directory, "log.txt"), "a")
            if "json" in self.formats:
                self.files["json"] = open(os.path.join(directory, "progress.json"), "a")
            if "csv" in self.formats:
                self.files["csv"] = open(os.path.join(directory, "progress.csv"), "w+")

    # ---- key/values
    def logkv(self, key, val):
        self.kvs[key] = val
        self.means.pop(key, None)

    def logkv_mean(self, key, val):
        self.means.setdefault(key, []).append(val)

    def snapshot(self):
        out = {k: _scalar(v) for k, v in self.kvs.items()}
        for k, vals in self.means.items():
            vals = [_scalar(v) for v in vals]
            out[k] = sum(vals) / len(vals)
        return out

    def dumpkvs(self):
        out = self.snapshot()
        if self.level < DISABLED and out:
            table = _table(out)
            if "stdout" in self.formats:
                sys.stdout.write(table)
                sys.stdout.flush()
            if "log" in self.files:
                self.files["log"].write(table)
                self.files["log"].flush()
            if "json" in self.files:
                self.files["json"].write(json.dumps(out, default=float) + "\n")
                self.files["json"].flush()
            if "csv" in self.files:
                self._write_csv(out)
        self.kvs.clear()
        self.means.clear()
        return out

    def _write_csv(self, out):
        f = self.files["csv"]
        new = [k for k in sorted(out) if k not in self.csv_keys]
        if new:                                   # widen the header: rewrite the file with the extra columns
            f.seek(0)
            lines = f.read().splitlines()
            self.csv_keys.extend(new)
            f.seek(0)
            f.truncate()
            f.write(",".join(self.csv_keys) + "\n")
            for line in lines[1:]:
                f.write(line + "," * len(new) + "\n")
        f.write(",".join("" if out.get(k) is None else str(out.get(k)) for k in self.csv_keys) + "\n")
        f.flush()

    # ---- text
    def log(self, *args, level=INFO):
        if level < self.level or (self.rank != 0 and level < WARN):
            return
        line = " ".join(str(a) for a in args) + "\n"
        sys.stdout.write(line)
        sys.stdout.flush()
        if "log" in self.files:
            self.files["log"].write(line)
            self.files["log"].flush()

    def close(self):
        for f in self.files.values():
            f.close()
        self.files = {}


def _table(kvs):
    rows = []
    for k in sorted(kvs):
        v = kvs[k]
        rows.append((str(k)[:30], f"{v:<8.3g}" if isinstance(v, float) else str(v)[:30]))
    kw = max(len(r[0]) for r in rows)
    vw = max(len(r[1]) for r in rows)
    bar = "-" * (kw + vw + 7) + "\n"
    return bar + "".join(f"| {k:<{kw}} | {v:<{vw}} |\n" for k, v in rows) + bar


def configure(dir=None, format_strs=None, comm=None, log_suffix=""):
    """Start logging into `dir` (default: $OPENAI_LOGDIR, else a time-stamped directory under the system temp dir)."""
    global _current
    if dir is None:
        dir = os.getenv("OPENAI_LOGDIR")
    if dir is None:
        dir = os.path.join(tempfile.gettempdir(), datetime.datetime.now().strftime("fcwdm-%Y-%m-%d-%H-%M-%S-%f"))
    dir = os.path.expanduser(dir)
    if format_strs is None:
        format_strs = os.getenv("OPENAI_LOG_FORMAT", "stdout,log,csv").split(",")
    formats = [f for f in format_strs if f in ("stdout", "log", "csv", "json")]
    if _current is not None:
        _current.close()
    _current = _Logger(dir, formats, _rank())
    if formats:
        log(f"Logging to {dir}")
    return _current


def get_current():
    if _current is None:
        configure()
    return _current


def reset():
    global _current
    if _current is not None:
        _current.close()
    _current = None


def logkv(key, val):
    get_current().logkv(key, val)


def logkv_mean(key, val):
    get_current().logkv_mean(key, val)


def logkvs(d):
    for k, v in d.items():
        logkv(k, v)


def dumpkvs():
    return get_current().dumpkvs()


def getkvs():
    return get_current().snapshot()


def log(*args, level=INFO):
    get_current().log(*args, level=level)


def debug(*args):
    log(*args, level=DEBUG)


def info(*args):
    log(*args, level=INFO)


def warn(*args):
    log(*args, level=WARN)


def error(*args):
    log(*args, level=ERROR)


def set_level(level):
    get_current().level = level


def get_dir():
    return get_current().dir


class profile_kv:
    """``with logger.profile_kv("name"):`` accumulates wall time under ``wait_name`` (reference logger.py:294-301)."""

    def __init__(self, name):
        self.key = "wait_" + name

    def __enter__(self):
        self.t0 = time.time()

    def __exit__(self, *exc):
        cur = get_current()
        cur.kvs[self.key] = cur.kvs.get(self.key, 0.0) + time.time() - self.t0
