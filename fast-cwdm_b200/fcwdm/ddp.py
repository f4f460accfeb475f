"""Data-parallel gradient exchange for the fcwdm training path: one process per GPU, replicated weights, ONE
collective per step -- a bucketed NCCL all-reduce (mean) of the flat fp32 gradient over NVLink/NVSwitch, launched
bucket by bucket from inside the backward as soon as every parameter of a bucket has its final gradient, on a side
stream, so the exchange hides under the remaining backward kernels (the reference itself
trains single-process, train.py:26-29 -- nothing there does this).

The training engine (fcwdm/train_engine.py) owns the flat gradient and calls ``ready(lo, hi)`` per parameter;
``GradSync`` only sees a flat tensor and ranges, so the bucketing logic is testable on CPU with gloo.
"""
import os

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, flat, offsets, bucket_bytes=32 << 20, group=None):
        """flat: the flat gradient tensor; offsets: [(lo, hi)] per parameter in flat order (gaps are alignment pad)."""
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        self.buckets = []            # [lo, hi, n_params]
        cur_lo, n = None, 0
        limit = max(1, bucket_bytes // flat.element_size())
        for lo, hi in offsets:
            if cur_lo is None:
                cur_lo, n = lo, 0
            n += 1
            if hi - cur_lo >= limit:
                self.buckets.append([cur_lo, hi, n])
                cur_lo = None
        if cur_lo is not None:
            self.buckets.append([cur_lo, offsets[-1][1], n])
        self._starts = [b[0] for b in self.buckets]
        self._pending = [b[2] for b in self.buckets]
        self._works = []
        self.stream = torch.cuda.Stream(flat.device) if flat.is_cuda else None
        self.launched = 0
        self._sent = [False] * len(self.buckets)
        # FCWDM_DDP_OVERLAP=0: hold every bucket until finish() (one exchange after the backward instead of under it);
        # a measurement knob -- NCCL's CTAs compete with the persistent conv kernels for SMs while they overlap
        self.overlap = os.environ.get("FCWDM_DDP_OVERLAP", "1") != "0"
        # FCWDM_DDP_GRAD_DTYPE=bf16: exchange the gradient as bf16 (half the bytes on the wire: the bucket is converted on the
        # side stream, averaged, and converted back into the fp32 flat gradient -- torch DDP's bf16_compress_hook).  The
        # gradients come out of bf16 activations (relative error ~1e-2 per tensor), so the extra 2^-9 rounding is below
        # their own noise; fp32 (the default) keeps the sum exact.
        self.wire_dtype = torch.bfloat16 if os.environ.get("FCWDM_DDP_GRAD_DTYPE", "fp32").lower() in ("bf16", "bfloat16") \
            else flat.dtype
        self._wire = {}              # bucket -> persistent wire buffer (bf16 mode)

    def begin(self):
        self._pending = [b[2] for b in self.buckets]
        self._works = []
        self.launched = 0
        self._sent = [False] * len(self.buckets)

    def _bucket_of(self, lo):
        import bisect
        return bisect.bisect_right(self._starts, lo) - 1

    def ready(self, lo, hi):
        """Parameter range [lo, hi) holds its final local gradient (all producing kernels are enqueued)."""
        b = self._bucket_of(lo)
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.world > 1 and self.overlap:
            self._launch(b)

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        self._sent[b] = True
        chunk = self.flat[lo:hi]
        avg = self.backend == "nccl"
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        wire = chunk
        if self.wire_dtype != chunk.dtype:
            wire = self._wire.get(b)
            if wire is None or wire.numel() != chunk.numel():
                wire = self._wire[b] = torch.empty(chunk.numel(), dtype=self.wire_dtype, device=chunk.device)
        if self.stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                if wire is not chunk:
                    wire.copy_(chunk)
                w = dist.all_reduce(wire, op=op, group=self.group, async_op=True)
        else:
            if wire is not chunk:
                wire.copy_(chunk)
            w = dist.all_reduce(wire, op=op, group=self.group, async_op=True)
        self._works.append((w, chunk, avg, wire))
        self.launched += 1

    def finish(self):
        """Join: every bucket reduced; the compute stream may read the averaged gradient afterwards."""
        if self.world > 1:
            for b in range(len(self.buckets)):
                if not self._sent[b]:                          # held back (overlap off) or never completed (unused params)
                    self._pending[b] = 0
                    self._launch(b)

            def settle(w, chunk, avg, wire):
                w.wait()
                if wire is not chunk:
                    chunk.copy_(wire)
                if not avg:
                    chunk.mul_(1.0 / self.world)

            for work in self._works:
                if self.stream is not None:
                    with torch.cuda.stream(self.stream):
                        settle(*work)
                else:
                    settle(*work)
            if self.stream is not None:
                torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
        self._works = []


def attach(model, bucket_bytes=32 << 20, group=None):
    """Make WavUNetModel's backward all-reduce its gradients across the process group (call after model.to(device))."""
    eng = model.train_engine()
    dev = next(model.parameters()).device
    flat = eng.flat_grad(dev)
    sync = GradSync(flat, eng.flat_offsets(), bucket_bytes=bucket_bytes, group=group)
    eng.grad_sync = sync
    eng.grad_ready_hook = sync.ready
    return sync


def broadcast_parameters(model, src=0, group=None):
    """Replicate rank `src`'s weights (DDP does this at construction)."""
    for p in model.parameters():
        dist.broadcast(p.data, src=src, group=group)
    for name in ("_engine", "_train_engine"):
        eng = getattr(model, name, None)
        if eng is not None:
            eng.invalidate()
