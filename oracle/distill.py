"""Oracle: the teacher half of the distillation step and q_sample / p_losses on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows /root/reference/utils/diffusion.py:68-100 (q_sample, p_losses)
and /root/reference/scripts/train_students.py:131-141 (teacher targets under no_grad), with the network evaluated by
oracle/unet.py on a bare state_dict.
"""
import torch
import torch.nn.functional as F

from . import unet


def extract(a, t, x_shape):
    out = a.gather(-1, torch.clamp(t, 0, a.shape[0] - 1))
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def q_sample(x_start, t, params, noise=None):
    """utils/diffusion.py:68-80; ``noise`` may be injected (else drawn from the global generator like the reference)."""
    if noise is None:
        noise = torch.randn_like(x_start)
    a = extract(params["sqrt_alphas_cumprod"], t, x_start.shape)
    b = extract(params["sqrt_one_minus_alphas_cumprod"], t, x_start.shape)
    return a * x_start + b * noise, noise


@torch.no_grad()
def p_losses(sd, x_start, t, params, cond=None, noise=None):
    """utils/diffusion.py:82-100 (forward value)."""
    x_noisy, noise = q_sample(x_start, t, params, noise)
    return F.mse_loss(unet.unet_forward(sd, x_noisy, t, cond), noise)


@torch.no_grad()
def teacher_targets(sd, images, t_teacher, params, noise=None):
    """scripts/train_students.py:131-141."""
    x_noisy, noise = q_sample(images, t_teacher, params, noise)
    ones = torch.ones(images.shape[0], 1)
    return x_noisy, noise, unet.unet_forward(sd, x_noisy, t_teacher, ones), unet.unet_forward(sd, x_noisy, t_teacher, None)
