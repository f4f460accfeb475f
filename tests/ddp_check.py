"""Run under torchrun (2+ GPUs): the bucketed, backward-overlapped NCCL all-reduce of fcwdm.ddp gives every rank the
mean of the per-rank gradients (checked against a plain all-reduce of gradients computed without the hook)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fcwdm import ddp  # noqa: E402
from oracle import wunet as ow  # noqa: E402
from oracle.make_golden import SMALL_CFG  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from guided_diffusion.wunet import WavUNetModel
    cfg = dict(SMALL_CFG, model_channels=64)
    model = WavUNetModel(**cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = ow.tie_output_blocks(ow.seeded_state_dict(shapes, seed=rank), 2)      # different weights per rank on purpose
    model.load_state_dict(sd)
    model.to(dev).train()
    ddp.broadcast_parameters(model)                                              # ... replicated from rank 0 here
    w0 = model.out[2].weight.detach().clone()
    dist.broadcast(w0, 0)
    assert torch.equal(w0, model.out[2].weight.detach())
    g = torch.Generator().manual_seed(50 + rank)
    x = torch.randn(1, 32, 8, 16, 8, generator=g).to(dev)
    t = torch.tensor([3 + rank], device=dev)

    def grads():
        for p in model.parameters():
            p.grad = None
        model(x, t).square().mean().backward()
        return torch.cat([p.grad.flatten() for p in model.parameters()])

    local_g = grads()
    want = local_g.clone()
    dist.all_reduce(want)
    want /= world
    sync = ddp.attach(model, bucket_bytes=1 << 20)
    got = grads()
    torch.cuda.synchronize()
    err = float((got - want).abs().max() / want.abs().max())
    print(f"rank {rank}: buckets {len(sync.buckets)} launched {sync.launched} max rel err {err:.2e}", flush=True)
    assert sync.launched == len(sync.buckets) and len(sync.buckets) > 1
    assert err <= 1e-6, err
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
