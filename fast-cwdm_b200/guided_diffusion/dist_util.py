"""Process-group helpers with the reference's names (guided_diffusion/dist_util.py:20-107) made to work with one
process per GPU -- SURVEY.md section 8f row 2.

The reference hard-codes a world of one (``RANK=0``, ``WORLD_SIZE=1``, a random port; dist_util.py:41-53) and its
``sync_params`` is commented out (:90-96), so ``scripts/train.py`` cannot run data-parallel.  Here:

* ``setup_dist`` honours the launcher's environment (``RANK`` / ``WORLD_SIZE`` / ``LOCAL_RANK`` / ``MASTER_ADDR`` /
  ``MASTER_PORT`` as ``torch.distributed.run`` sets them): NCCL on the GPU ``LOCAL_RANK`` names, gloo on CPU.  Without
  a launcher it falls back to the reference's behaviour, a single-process group on the loopback interface;
* ``dev()`` returns this rank's device;
* ``sync_params`` really broadcasts rank 0's tensors; ``load_state_dict`` reads the file once (rank 0) and broadcasts
  the bytes.
"""
import io
import os
import socket

import torch as th
import torch.distributed as dist

GPUS_PER_NODE = 8
SETUP_RETRY_COUNT = 3


def _launched():
    return "RANK" in os.environ and "WORLD_SIZE" in os.environ


def local_rank():
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    return int(os.environ.get("RANK", "0")) % GPUS_PER_NODE


def setup_dist(devices=(0,)):
    """Create the default process group (idempotent).  `devices` is the reference's single-process device list
    (train.py:55): it selects the GPU only when no launcher set LOCAL_RANK."""
    if dist.is_initialized():
        return
    cuda = th.cuda.is_available()
    if _launched():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if cuda:
            th.cuda.set_device(local_rank())
            dist.init_process_group(backend="nccl", init_method="env://", device_id=th.device("cuda", local_rank()))
        else:
            dist.init_process_group(backend="gloo", init_method="env://")
        return
    try:
        first = int(list(devices)[0])
    except (TypeError, IndexError, ValueError):
        first = int(devices) if isinstance(devices, (int, str)) and str(devices).isdigit() else 0
    if cuda and first < th.cuda.device_count():
        th.cuda.set_device(first)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_find_free_port())
    os.environ["RANK"] = "0"
    os.environ["WORLD_SIZE"] = "1"
    dist.init_process_group(backend="nccl" if cuda else "gloo", init_method="env://")


def get_rank():
    return dist.get_rank() if dist.is_initialized() else int(os.environ.get("RANK", "0"))


def get_world_size():
    return dist.get_world_size() if dist.is_initialized() else int(os.environ.get("WORLD_SIZE", "1"))


def dev(device_number=0):
    """This process's device.  Under a launcher it is cuda:LOCAL_RANK whatever `device_number` says (one process per
    GPU); otherwise the reference's rules (dist_util.py:56-71): the current device for 0, an explicit index checked
    against the device count, a list for a list."""
    if isinstance(device_number, (list, tuple)):
        return [dev(k) for k in device_number]
    if not th.cuda.is_available():
        return th.device("cpu")
    if "LOCAL_RANK" in os.environ:
        return th.device("cuda", local_rank())
    count = th.cuda.device_count()
    if device_number >= count:
        raise ValueError(f"requested device number {device_number} (0-indexed) but only {count} devices available")
    if device_number == 0:
        return th.device("cuda", th.cuda.current_device())
    return th.device("cuda", device_number)


def load_state_dict(path, **kwargs):
    """torch.load of `path`, the file being read by rank 0 only and its bytes broadcast to the other ranks."""
    world = get_world_size()
    data = None
    if get_rank() == 0:
        with open(path, "rb") as f:
            data = f.read()
    if world > 1:
        box = [data]
        dist.broadcast_object_list(box, src=0)
        data = box[0]
    return th.load(io.BytesIO(data), **kwargs)


def sync_params(params):
    """Overwrite every tensor in `params` with rank 0's copy (no-op in a world of one)."""
    if get_world_size() == 1:
        return
    with th.no_grad():
        for p in params:
            dist.broadcast(p.data if hasattr(p, "data") else p, src=0)


def _find_free_port():
    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    try:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]
    finally:
        s.close()
