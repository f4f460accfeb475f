"""Sharded sampling driver: the loops of the reference's scripts/sample.py:56-149 and scripts/sample_auto.py:44-164
(and of complete_dataset.py's per-case synthesis) as a three-stage pipeline per GPU -- SURVEY.md section 8f row 3.

At ~20 volumes/s a B200 finishes a case in ~50 ms, while reading four gzipped NIfTI volumes, two host-side
``np.quantile`` sorts per modality and gzipping the result cost the reference's loop several hundred ms of CPU per
case, serialised with the GPU work.  Here, per rank:

    reader threads    NIfTI gunzip + parse of case i+k into pinned buffers          (zlib releases the GIL)
    GPU (this thread) VolumeStream(raw=True): H2D, clip/normalise/pad/crop, cond DWT, p_sample_loop, final IDWT,
                      D2H -- the uploads / downloads of neighbouring cases overlap the denoising
    writer threads    gzip level 1 + atomic write of the result(s) of case i-k

Cases are independent, so ranks just take disjoint slices of the case list (``pipeline.shard_indices``): no
collective.  The initial noise is drawn per case from ``torch.Generator().manual_seed(seed + case_index)`` on the host
and the device generator (the per-step noise of ``p_sample``) is re-seeded with the same number before each case, so a
case's result does not depend on how many ranks there are or on its position in the loop -- the reference seeds once
and draws sequentially (sample.py:51,100).

Two output conventions, as in the reference:

* ``mode='sample'`` (sample.py): fixed target contrast; ``<output_dir>/<subject>/sample.nii.gz`` (+ ``target.nii.gz``),
  array (224, 224, 155) float32 in [0, 1], background of the first condition zeroed, identity affine;
* ``mode='auto'`` (sample_auto.py / complete_dataset.py): the missing modality of each case is synthesised with the
  model registered for it; values <= 0.04 are zeroed, the volume is padded back to (240, 240, 155) and written as
  ``<case_dir>/<subject>-<missing>.nii.gz`` with the header of a present modality.
"""
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import nifti
from .pipeline import VolumeStream, shard_indices

MODALITIES = ("t1n", "t1c", "t2w", "t2f")
RAW_SHAPE = (240, 240, 155)


def conditions_for(contr):
    """The three conditioning modalities for a target contrast, in the reference's order (sample.py:64-89)."""
    if contr not in MODALITIES:
        raise ValueError(f"This contrast can't be synthesized: {contr!r}")
    return tuple(m for m in MODALITIES if m != contr)


def subject_of(path, marker=None):
    """Subject id of a case file.  With `marker` (e.g. 'validation/') the reference's rule: 19 characters after it
    (sample.py:61); otherwise the name of the directory that holds the file."""
    if marker and marker in path:
        return path.split(marker)[1][:19]
    return os.path.basename(os.path.dirname(path))


class _Case:
    __slots__ = ("index", "subject", "contr", "files", "volume", "noise", "header", "error")


class SamplingDriver:
    def __init__(self, diffusion, models, cases, output_dir=None, mode="sample", contr=None, rank=0, world_size=1,
                 seed=0, device=None, reader_threads=4, writer_threads=4, depth=3, write_target=True,
                 subject_marker=None, compresslevel=1):
        """models: one model (mode='sample') or {contrast: model} (mode='auto'); cases: a sequence of file dicts
        ``{'t1n': path, ...}`` (``BRATSVolumes(...).database``)."""
        if mode not in ("sample", "auto"):
            raise ValueError("mode must be 'sample' or 'auto'")
        if mode == "sample" and contr not in MODALITIES:
            raise ValueError("mode='sample' needs contr in " + str(MODALITIES))
        if mode == "sample" and output_dir is None:
            raise ValueError("mode='sample' needs output_dir")
        self.diffusion = diffusion
        self.models = models if isinstance(models, dict) else {contr: models}
        self.cases = list(cases)
        self.output_dir = output_dir
        self.mode, self.contr = mode, contr
        self.rank, self.world = rank, world_size
        self.seed = seed
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = max(2, depth, reader_threads)          # cases being read ahead of the GPU
        self.n_writers = writer_threads
        self.write_target = write_target and mode == "sample"
        self.marker = subject_marker
        self.compresslevel = compresslevel
        self.readers = ThreadPoolExecutor(max_workers=reader_threads, thread_name_prefix="fcwdm-read")
        # the 3-4 modality files of ONE case are gunzipped concurrently (zlib releases the GIL): the first case reaches the GPU
        # after one file's inflate time instead of four, and a short run is not dominated by that ramp
        self.file_readers = ThreadPoolExecutor(max_workers=max(4, reader_threads), thread_name_prefix="fcwdm-file")
        self.writers = ThreadPoolExecutor(max_workers=writer_threads, thread_name_prefix="fcwdm-write")
        self.my_indices = shard_indices(len(self.cases), rank, world_size)
        self.stats = {"cases": 0, "read_s": 0.0, "write_s": 0.0, "bytes_written": 0, "skipped": []}
        self._lock = threading.Lock()
        self._streams = {}

    # ------------------------------------------------------------------------------------------ stage 1: read
    def _target_of(self, files):
        if self.mode == "sample":
            return self.contr
        absent = [m for m in MODALITIES if m not in files]
        return absent[0] if len(absent) == 1 else None

    def _read_case(self, index, buffers):
        t0 = time.time()
        case = _Case()
        case.index, case.files, case.error, case.header = index, self.cases[index], None, None
        try:
            case.contr = self._target_of(case.files)
            if case.contr is None:
                raise ValueError("auto mode needs exactly one missing modality")
            if case.contr not in self.models:
                raise KeyError(f"no model registered for contrast {case.contr!r}")
            conds = conditions_for(case.contr)
            first = case.files[conds[0]]
            case.subject = subject_of(first, self.marker)
            vol, noise = buffers
            order = (case.contr,) + conds                   # channel 0 = target slot (zeros when absent), 1..3 = conditions

            def load(c, m):
                arr, hdr = nifti.read(case.files[m], dtype=np.float32, return_header=True)
                if tuple(arr.shape) != RAW_SHAPE:
                    raise ValueError(f"{case.files[m]}: shape {arr.shape}, expected {RAW_SHAPE}")
                vol[0, c].copy_(torch.from_numpy(arr.T))          # file order (Z, Y, X): a straight memcpy, no transpose
                return hdr

            loads = {c: self.file_readers.submit(load, c, m) for c, m in enumerate(order) if m in case.files}
            for c, m in enumerate(order):
                if m not in case.files:
                    vol[0, c].zero_()
            g = torch.Generator().manual_seed(self.seed + index)
            torch.randn(noise.shape, generator=g, out=noise)      # while the files inflate
            failed = None
            for c, fut in loads.items():                          # wait for ALL of them: a straggler must not write into a
                try:                                              # buffer that has already gone to the next case
                    hdr = fut.result()
                    if c == 1:
                        case.header = hdr
                except Exception as exc:
                    failed = failed or exc
            if failed is not None:
                raise failed
            case.volume, case.noise = vol, noise
        except Exception as exc:                            # a bad case is reported and skipped, the run goes on
            case.error = exc
        with self._lock:
            self.stats["read_s"] += time.time() - t0
        return case

    # ------------------------------------------------------------------------------------------ stage 3: write
    def _write_case(self, case, result, target, done_event, release_in, release_out):
        t0 = time.time()
        held_in = True
        try:
            done_event.synchronize()                        # this case's H2D and D2H are both done
            release_in()                                    # ... so its input buffers can take the next read already
            held_in = False
            n = 0
            if self.mode == "sample":
                folder = os.path.join(self.output_dir, case.subject)
                os.makedirs(folder, exist_ok=True)
                # buffers hold the volume in file order (Z, Y, X); .T is the (X, Y, Z) array, already Fortran-contiguous
                n += nifti.write(os.path.join(folder, "sample.nii.gz"), result[0].numpy().T, np.eye(4),
                                 compresslevel=self.compresslevel)
                if target is not None:
                    n += nifti.write(os.path.join(folder, "target.nii.gz"), target[0].numpy().T, np.eye(4),
                                     compresslevel=self.compresslevel)
            else:
                if self.output_dir is not None:
                    folder = os.path.join(self.output_dir, case.subject)
                else:                                       # next to the inputs, as sample_auto.py:77 does
                    folder = os.path.dirname(case.files[conditions_for(case.contr)[0]])
                os.makedirs(folder, exist_ok=True)
                full = np.zeros(RAW_SHAPE[::-1], dtype=np.float32)        # pad back to 240 x 240 (sample_auto.py:143)
                full[:, 8:-8, 8:-8] = result[0].numpy()
                n += nifti.write(os.path.join(folder, f"{case.subject}-{case.contr}.nii.gz"), full.T, like=case.header,
                                 compresslevel=self.compresslevel)
            with self._lock:
                self.stats["bytes_written"] += n
                self.stats["cases"] += 1
        finally:
            if held_in:
                release_in()
            release_out()
            with self._lock:
                self.stats["write_s"] += time.time() - t0

    # ------------------------------------------------------------------------------------------ stage 2: GPU
    def _stream_for(self, contr):
        if contr not in self._streams:
            model = self.models[contr]
            model.eval()
            self._streams[contr] = VolumeStream(self.diffusion, model, self.device, raw=True, crop=RAW_SHAPE[2],
                                                post=self.mode, file_order=True)
        return self._streams[contr]

    def run(self):
        """Process this rank's cases; returns the statistics dict (cases written, bytes, stage times, wall time)."""
        t_start = time.time()
        pin = torch.cuda.is_available()
        free_in = queue.Queue()                             # pinned staging: inputs are held from read to upload,
        free_out = queue.Queue()                            # outputs from download to the end of the file write
        for _ in range(self.depth + 2):
            free_in.put((torch.empty((1, 4) + RAW_SHAPE[::-1], dtype=torch.float32, pin_memory=pin),
                         torch.empty((1, 8, 112, 112, 80), dtype=torch.float32, pin_memory=pin)))
        for _ in range(self.n_writers + 2):
            free_out.put((torch.empty((1, RAW_SHAPE[2], 224, 224), dtype=torch.float32, pin_memory=pin),
                          torch.empty((1, RAW_SHAPE[2], 224, 224), dtype=torch.float32, pin_memory=pin)
                          if self.write_target else None))
        pending = []                                        # futures of cases being read, in order
        todo = list(self.my_indices)
        writes = []

        def top_up():
            while todo and len(pending) < self.depth:
                bufs = free_in.get()
                pending.append((self.readers.submit(self._read_case, todo.pop(0), bufs), bufs))

        top_up()
        with torch.no_grad():
            while pending:
                fut, bufs = pending.pop(0)
                case = fut.result()
                top_up()
                if case.error is not None:
                    self.stats["skipped"].append((case.index, repr(case.error)))
                    free_in.put(bufs)
                    continue
                stream = self._stream_for(case.contr)
                out_host, tgt_host = outs = free_out.get()
                # if the next case has already been read (and uses the same model), its upload is issued now, ahead of
                # this case's download on the copy stream, so it overlaps this case's denoising
                nxt = None
                if pending and pending[0][0].done():
                    peek = pending[0][0].result()
                    if peek.error is None and peek.contr == case.contr:
                        nxt = (peek.volume, peek.noise)
                torch.cuda.manual_seed(self.seed + case.index)       # the per-step noise of p_sample: per case, like `noise`
                stream.submit(case.volume, case.noise, out_host, next_case=nxt)
                target = None
                if self.write_target:
                    target = self._normalised_target(stream, case, tgt_host)
                done = torch.cuda.Event()
                with torch.cuda.stream(stream.copy_stream):
                    done.record(stream.copy_stream)
                # `done` is recorded on the copy stream after this case's H2D and D2H: the writer releases the input
                # buffers as soon as it fires and the output buffers after the files are written
                writes.append(self.writers.submit(self._write_case, case, out_host, target, done,
                                                  lambda b=bufs: free_in.put(b), lambda o=outs: free_out.put(o)))
            for s in self._streams.values():
                s.finish()
        for w in writes:
            w.result()
        self.stats["wall_s"] = time.time() - t_start
        self.stats["rank"], self.stats["world_size"] = self.rank, self.world
        return self.stats

    def _normalised_target(self, stream, case, tgt_host):
        """The ground-truth target as the loader would have normalised it, cropped like the sample (sample.py:133-136)."""
        from . import preprocess
        raw = case.volume[0, :1].to(self.device, non_blocking=True).permute(0, 3, 2, 1)      # file order -> (1, X, Y, Z)
        tgt = preprocess.clip_and_normalize(raw)[:, 0, :, :, :RAW_SHAPE[2]].permute(0, 3, 2, 1)
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(stream.copy_stream):
            stream.copy_stream.wait_event(ev)
            tgt.record_stream(stream.copy_stream)
            tgt_host.copy_(tgt, non_blocking=True)
        return tgt_host

    def close(self):
        self.readers.shutdown(wait=True)
        self.file_readers.shutdown(wait=True)
        self.writers.shutdown(wait=True)
