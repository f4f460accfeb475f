"""TEST INFRASTRUCTURE ONLY -- CPU restatement of one training step of the reference.

TrainLoop.forward_backward / run_step (guided_diffusion/train_util.py:364-460): training_losses in mode 'i2i'
(gaussian_diffusion.py:1084-1166), loss = mean(mse_wav * ones(8)) (:447-449), loss.backward() through the U-Net,
torch.optim.AdamW step (:75-82, :391).  The U-Net forward is oracle.wunet.wunet_forward (torch CPU fp32 functional
ops, differentiable end to end), so torch autograd over it IS the reference's backward arithmetic.
"""
import torch

from . import diffusion as od
from . import wunet as ow


def training_step_grads(sd, tab, batch, t, noise, *, model_channels, channel_mult, contr="t1n", timestep_map=None,
                        forward=None):
    """sd: state dict of fp32 tensors (tied keys share one tensor object).  Returns (loss, mse_wav, model_output,
    grads) with grads[key] for every key of sd (tied keys share the accumulated gradient).  forward: the denoiser
    restatement (default oracle.wunet.wunet_forward; oracle.unet.unet_forward for the plain U-Net)."""
    forward = forward or ow.wunet_forward
    leaves = {}
    for k, v in sd.items():
        if id(v) not in leaves:
            leaves[id(v)] = v.detach().clone().requires_grad_(True)
    live = {k: leaves[id(v)] for k, v in sd.items()}
    model = lambda x, tt: forward(live, x, tt, model_channels=model_channels, channel_mult=channel_mult)
    terms, out, _ = od.training_losses(tab, model, batch, t, contr=contr, timestep_map=timestep_map, noise=noise)
    loss = (terms["mse_wav"] * torch.ones(8)).mean()                                   # train_util.py:447-449
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in live.items()}
    return loss.detach(), terms["mse_wav"].detach(), out.detach(), grads


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW's update rule written out (decoupled weight decay, bias-corrected moments)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    denom = v.sqrt() / (1.0 - beta2 ** step) ** 0.5 + eps
    p = p - (lr / (1.0 - beta1 ** step)) * m / denom
    return p, m, v
