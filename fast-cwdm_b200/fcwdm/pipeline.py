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


def synthesize(diffusion, model, cond_1, cond_2, cond_3, noise, clip_denoised=True, crop=155, progress=False):
    """Returns the synthesised modality (N, D, H, min(W, crop)) in [0, 1], masked by cond_1's background."""
    with torch.no_grad():
        cond = build_cond(cond_1, cond_2, cond_3)
        sample = diffusion.p_sample_loop(model, tuple(noise.shape), noise=noise, cond=cond,
                                         clip_denoised=clip_denoised, model_kwargs={}, progress=progress)
        image = ops.sample_to_image(sample, cond_1)
    return image.squeeze(1)[:, :, :, :crop]


def shard_indices(n_items, rank, world_size):
    """Contiguous, balanced, disjoint partition of range(n_items) over ranks (first `n % world` ranks get one
    extra item)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))
