"""Forward-only half of the distillation step.

scripts/train_students.py:131-141 computes, under ``torch.no_grad()``, the teacher's targets for a batch:

    t_teacher ~ randint(0, teacher_steps, (B,));  x_noisy, noise = q_sample(images, t_teacher, teacher_params)
    teacher_pred_cond   = teacher(x_noisy, t_teacher, cond=ones(B, 1))
    teacher_pred_uncond = teacher(x_noisy, t_teacher, cond=None)

Here both predictions come from ONE forward of 2B rows (per-row timesteps, per-row conditioning variant) through the same
kernels as the sampling loops; the student's forward/backward and the optimizer stay with the caller (out of scope).
"""
import torch

from ._lib import VAR_COND1, VAR_NONE
from .engine import UNetEngine, check_device_errors, get_precision
from .utils.diffusion import q_sample


@torch.no_grad()
def teacher_targets(teacher_model, images, t_teacher, teacher_params, precision=None):
    """Returns (x_noisy, noise, teacher_pred_cond, teacher_pred_uncond), all [B, C, H, W] on the images' device.
    ``t_teacher``: int64 [B].  Noise is drawn exactly as the reference's q_sample does (global generator)."""
    teacher_model.eval()
    x_noisy, noise = q_sample(images, t_teacher, teacher_params)
    B = images.shape[0]
    n_t = int(teacher_params["betas"].shape[0])
    eng = UNetEngine.for_model(teacher_model, images.shape[2], n_t, precision or get_precision("forward"), images.device)
    x2 = torch.cat([x_noisy, x_noisy])
    t2 = torch.cat([t_teacher, t_teacher])
    variants = torch.cat([torch.full((B,), VAR_COND1, dtype=torch.int32), torch.full((B,), VAR_NONE, dtype=torch.int32)])
    eps = eng.forward(x2, t2, variants)
    check_device_errors()
    return x_noisy, noise, eps[:B], eps[B:]
