"""Per-volume synthesis pipeline = the loop body of the reference's scripts/sample.py:92-131, on fcwdm kernels.

    cond   = cat(DWT(cond_1), DWT(cond_2), DWT(cond_3)) with LLL/3        (sample.py:92-97)   3 launches
    sample = diffusion.p_sample_loop(model, noise=noise, cond=cond)        (sample.py:104-111) fused sampler
    image  = clamp(IDWT(sample; LLL*3), 0, 1); image[cond_1 == 0] = 0      (sample.py:113-125) 1 launch
    image  = image[..., :155]                                              (sample.py:131)     view

`shard_indices` partitions a list of volumes over ranks for multi-GPU sampling: volumes are independent, so
there is no collective on the data path."""
import torch

from . import ops


def build_cond(cond_1, cond_2, cond_3):
    """3 x (N,1,D,H,W) image-space modalities -> (N,24,D/2,H/2,W/2) conditioning sub-bands."""
    N, _, D, H, W = cond_1.shape
    out = torch.empty((N, 24, D // 2, H // 2, W // 2), dtype=torch.float32, device=cond_1.device)
    for k, c in enumerate((cond_1, cond_2, cond_3)):
        out[:, 8 * k:8 * k + 8] = ops.dwt3d_planar(c.float(), lll_scale=1.0 / 3.0, concat=True)
    return out


def synthesize(diffusion, model, cond_1, cond_2, cond_3, noise, clip_denoised=True, crop=155, progress=False,
               post="sample"):
    """Returns the synthesised modality (N, D, H, min(W, crop)).  post='sample': clamped to [0, 1] and masked by
    cond_1's background (sample.py:113-125); post='auto': values <= 0.04 zeroed, nothing else (sample_auto.py:137)."""
    with torch.no_grad():
        cond = build_cond(cond_1, cond_2, cond_3)
        sample = diffusion.p_sample_loop(model, tuple(noise.shape), noise=noise, cond=cond,
                                         clip_denoised=clip_denoised, model_kwargs={}, progress=progress)
        if post == "sample":
            image = ops.sample_to_image(sample, cond_1)
        elif post == "auto":
            image = ops.idwt3d_planar(sample, lll_scale=3.0, concat=True)
            image = torch.where(image <= 0.04, torch.zeros_like(image), image)
        else:
            raise ValueError(f"unknown post-processing {post!r}")
    return image.squeeze(1)[:, :, :, :crop]


def shard_indices(n_items, rank, world_size):
    """Contiguous, balanced, disjoint partition of range(n_items) over ranks (first `n % world` ranks get one
    extra item)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


class VolumeStream:
    """Throughput-oriented synthesis of a sequence of host-resident cases on one GPU (what the loop of
    scripts/sample.py:56-149 does one case at a time, synchronously).

    Two device input slots and two output slots: the pinned-host -> device copy of case i+1 and the device -> host
    copy of case i-1 run on a copy stream underneath the denoising of case i, so the PCIe traffic (128 MB in, 31 MB out
    per 224x224x160 case) is hidden behind ~50 ms of compute.  ``raw=True`` feeds un-normalised (4, 240, 240, 155)
    volumes and runs the loader's clip/normalise/pad/crop on the GPU first (fcwdm.preprocess).

        stream = VolumeStream(diffusion, model, device)
        for volume_host, noise_host, out_host in cases:      # pinned tensors
            stream.submit(volume_host, noise_host, out_host)
        stream.finish()                                       # all results are in their out_host buffers
    """

    def __init__(self, diffusion, model, device, raw=False, crop=155, post="sample", file_order=False):
        """file_order=True (with raw=True): volumes arrive as (N, 4, Z, Y, X) -- the NIfTI file order, first index
        fastest -- and results leave as (N, crop, H, W), so the host never transposes a volume; the two axis swaps ride
        on device copies that exist anyway (the contiguous() in front of the preprocessing, the D2H staging copy)."""
        self.diffusion, self.model, self.device, self.raw, self.crop = diffusion, model, device, raw, crop
        self.post = post
        self.file_order = file_order
        if file_order and not raw:
            raise ValueError("file_order=True is for raw NIfTI volumes (raw=True)")
        self.copy_stream = torch.cuda.Stream(device)
        self._slots = None
        self._i = 0
        self._pending = None          # (host volume, host noise) whose H2D has been issued into the next slot

    def _alloc(self, vol, noise, out):
        dev = self.device
        self._slots = [(torch.empty(vol.shape, dtype=vol.dtype, device=dev), torch.empty(noise.shape, dtype=noise.dtype, device=dev))
                       for _ in range(2)]
        self._out = [torch.empty(out.shape, dtype=torch.float32, device=dev) for _ in range(2)]
        self._h2d_done = [torch.cuda.Event() for _ in range(2)]
        self._slot_free = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        for ev in self._slot_free:
            ev.record(cur)

    def _prefetch(self, slot, vol, noise):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._slot_free[slot])      # the compute that read this slot has finished
            self._slots[slot][0].copy_(vol, non_blocking=True)
            self._slots[slot][1].copy_(noise, non_blocking=True)
            self._h2d_done[slot].record(self.copy_stream)

    def prefetch(self, vol, noise):
        """Optionally announce the NEXT case early so that its upload overlaps the current case's compute."""
        if self._slots is not None:
            self._prefetch(self._i % 2, vol, noise)
            self._pending = (vol, noise)

    def submit(self, vol, noise, out_host, next_case=None):
        """vol: (N, 4, D, H, W) pinned host tensor (channel 0 = the modality to synthesise, unused; 1..3 = conditions) or,
        with raw=True, (N, 4, 240, 240, 155) raw intensities; noise: (N, 8, D/2, H/2, W/2); out_host: pinned
        (N, D, H, crop) destination.  next_case = (vol, noise) of the following call, uploaded during this one."""
        if self._slots is None:
            self._alloc(vol, noise, out_host)
        cur = torch.cuda.current_stream(self.device)
        slot = self._i % 2
        if self._pending is None or self._pending[0] is not vol:
            self._prefetch(slot, vol, noise)
        self._pending = None
        cur.wait_event(self._h2d_done[slot])
        self._i += 1
        if next_case is not None:
            self.prefetch(*next_case)
        v, nz = self._slots[slot]
        if self.raw:
            from . import preprocess
            N = v.shape[0]
            if self.file_order:
                v = v.permute(0, 1, 4, 3, 2)                              # (N, 4, X, Y, Z) view of the file-order buffer
            v = preprocess.clip_and_normalize(v.reshape((N * 4,) + tuple(v.shape[2:]))).reshape(
                (N, 4) + (v.shape[2] - 16, v.shape[3] - 16, 160))
        img = synthesize(self.diffusion, self.model, v[:, 1:2], v[:, 2:3], v[:, 3:4], nz, crop=self.crop, post=self.post)
        self._out[slot].copy_(img.permute(0, 3, 2, 1) if self.file_order else img)
        self._slot_free[slot].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._slot_free[slot])
            out_host.copy_(self._out[slot], non_blocking=True)      # D2H of the finished case, off the compute stream
        return img

    def finish(self):
        torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)
