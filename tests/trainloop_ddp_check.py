"""Run under torchrun (2+ GPUs): the sequence scripts/train.py performs (reference train.py:55-101) -- setup_dist,
create_model_and_diffusion, model.to(dist_util.dev()), create_named_schedule_sampler, TrainLoop(...).run_loop() -- with
one process per GPU.  Checks: every rank starts from rank 0's weights although each built its own; after a few steps
on DIFFERENT per-rank data the weights are still identical on all ranks (the gradients were averaged) and have moved;
only rank 0 wrote the checkpoint."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200")]
import numpy as np  # noqa: E402
import torch as th  # noqa: E402
import torch.distributed as dist  # noqa: E402

from guided_diffusion import dist_util, logger  # noqa: E402
from guided_diffusion.resample import create_named_schedule_sampler  # noqa: E402
from guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults  # noqa: E402
from guided_diffusion.train_util import TrainLoop  # noqa: E402

KEYS = ("t1n", "t1c", "t2w", "t2f")


class Volumes(th.utils.data.Dataset):
    def __init__(self, seed, n=4):
        g = th.Generator().manual_seed(seed)
        self.items = [{k: th.rand(1, 32, 32, 16, generator=g) for k in KEYS} for _ in range(n)]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return dict(self.items[i], missing="none", subj="dummy_string")


def main():
    root = sys.argv[1]
    os.environ["FCWDM_CHECKPOINT_ROOT"] = root
    dist_util.setup_dist(devices=[0])
    rank, world = dist.get_rank(), dist.get_world_size()
    seed = 0 + rank                                         # seed + rank for data / noise / t
    th.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    logger.configure(dir=os.path.join(root, "log"), format_strs=["csv"])
    args = model_and_diffusion_defaults()
    args.update(image_size=32, num_channels=64, num_res_blocks=2, channel_mult="1,2", in_channels=32, out_channels=8,
                dims=3, num_groups=32, bottleneck_attention=False, resblock_updown=True, use_freq=True,
                attention_resolutions="", predict_xstart=True, use_scale_shift_norm=False, diffusion_steps=10, sample_schedule="sampled",
                mode="i2i", learn_sigma=False)
    model, diffusion = create_model_and_diffusion(**args)
    with th.no_grad():                                      # zero-initialised convs would hide most of the backward
        g = th.Generator().manual_seed(100 + rank)          # ... and every rank starts from DIFFERENT weights
        for p in model.parameters():
            if float(p.abs().max()) == 0.0:
                p.copy_(th.randn(p.shape, generator=g) * 0.02)
            else:
                p.add_(th.randn(p.shape, generator=g) * 0.01)
    model.to(dist_util.dev())
    assert next(model.parameters()).device == th.device("cuda", int(os.environ["LOCAL_RANK"]))
    before = th.cat([p.detach().flatten() for p in model.parameters()]).clone()
    sampler = create_named_schedule_sampler("uniform", diffusion, maxt=diffusion.num_timesteps)
    data = th.utils.data.DataLoader(Volumes(seed=7 + rank), batch_size=2, shuffle=False)
    loop = TrainLoop(model=model, diffusion=diffusion, data=data, batch_size=2, in_channels=32, image_size=32,
                     microbatch=-1, lr=1e-3, ema_rate="0.9999", log_interval=1, contr="t1n", save_interval=2,
                     resume_checkpoint="", resume_step=0, use_fp16=False, schedule_sampler=sampler, weight_decay=0.0,
                     lr_anneal_steps=4, dataset="brats", summary_writer=None, mode="i2i", sample_schedule="sampled",
                     diffusion_steps=10)
    assert loop.global_batch == 2 * world and (loop.grad_sync is not None) == (world > 1)
    start = th.cat([p.detach().flatten() for p in model.parameters()]).clone()
    ref = start.clone()
    dist.broadcast(ref, 0)
    assert th.equal(ref, start), "weights were not replicated from rank 0"
    if rank != 0:
        assert not th.equal(before, start)
    loop.run_loop()
    th.cuda.synchronize()
    after = th.cat([p.detach().flatten() for p in model.parameters()])
    ref = after.clone()
    dist.broadcast(ref, 0)
    drift = float((after - ref).abs().max())
    moved = float((after - start).abs().max())
    assert drift == 0.0, f"rank {rank}: weights diverged across ranks by {drift}"
    assert moved > 1e-4, "weights did not move"
    dist.barrier()
    ck = os.path.join(root, "checkpoints", "brats_t1n_BEST_sampled_10.pt")
    if rank == 0:
        assert os.path.exists(ck)
        print(f"trainloop ddp ok: world {world}, weights moved {moved:.3e}, cross-rank drift {drift}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
