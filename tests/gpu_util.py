"""Helpers shared by the -m gpu tests (all compute goes through the C-ABI via fcwdm.ops)."""
import numpy as np
import torch

from fcwdm import ops


def to_cl(x, ld=None):
    """planar f32 (N,C,D,H,W) CUDA -> cl bf16 buffer (N*S, ld) (pad channels zero), through the product converter."""
    N, C = x.shape[:2]
    S = x[0, 0].numel()
    ld = ld or ((C + 63) // 64 * 64)
    buf = torch.zeros((N * S, ld), dtype=torch.bfloat16, device=x.device)
    ops.planar_to_cl(x.float(), buf, C)
    return buf


def from_cl(buf, shape):
    """cl bf16 buffer -> planar f32 tensor of `shape` (N,C,D,H,W)."""
    out = torch.empty(shape, dtype=torch.float32, device=buf.device)
    ops.cl_to_planar(buf, out, shape[1])
    return out


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def np_t(a, device="cuda"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)
