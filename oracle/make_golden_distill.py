"""Generate tests/golden/distill.npz by running the UNMODIFIED reference (/root/reference) on CPU:
q_sample, p_losses (forward value) and the teacher half of the distillation step with a timestep per row.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python -m oracle.make_golden_distill

Teachers are the ones of tests/golden/tiny16.npz / tiny32.npz (same seeds, checked by weight checksum); inputs are
seeded; outputs are whatever utils/diffusion.py:68-100 and the no_grad block of scripts/train_students.py:131-141
compute (that block is restated here line by line because it lives inside a training loop).
"""
import os

import numpy as np
import torch

from . import refload
from .make_golden import OUT, make_model


def main():
    ref = refload.load()
    g = {}
    for name, C, H, T, sf_t in (("tiny16", 1, 16, 6, 0.1), ("tiny32", 3, 32, 4, 0.2)):
        cfg = refload.RefConfig(channels=C, image_size=H, timesteps=T)
        teacher = make_model(ref, cfg, sf_t, 100)
        wsum = float(sum(v.double().sum() for v in teacher.state_dict().values()))
        old = np.load(os.path.join(OUT, name + ".npz"))
        assert abs(wsum - float(old["teacher_wsum"][0])) <= 1e-6 * max(1.0, abs(wsum)), "teacher differs from the committed fixture"
        params = ref.diffusion.get_diffusion_params(T, cfg)
        B = 5
        torch.manual_seed(77)
        images = torch.randn(B, C, H, H).clamp(-1, 1)
        t_teacher = torch.tensor([T - 1, 0, 1, T // 2, 0][:B], dtype=torch.long)
        g[f"{name}/images"] = images.numpy()
        g[f"{name}/t"] = t_teacher.numpy()
        # ---- scripts/train_students.py:131-141
        torch.manual_seed(78)
        with torch.no_grad():
            x_noisy, noise = ref.diffusion.q_sample(images, t_teacher, params)
            pred_cond = teacher(x_noisy, t_teacher, cond=torch.ones(B, 1))
            pred_uncond = teacher(x_noisy, t_teacher, cond=None)
        g[f"{name}/x_noisy"], g[f"{name}/noise"] = x_noisy.numpy(), noise.numpy()
        g[f"{name}/pred_cond"], g[f"{name}/pred_uncond"] = pred_cond.numpy(), pred_uncond.numpy()
        # ---- utils/diffusion.py:82-100 (loss value)
        for tag, cond in (("none", None), ("cond1", torch.ones(B, 1))):
            torch.manual_seed(79)
            with torch.no_grad():
                loss = ref.diffusion.p_losses(teacher, images, t_teacher, params, cond)
            g[f"{name}/p_losses_{tag}"] = np.array([float(loss)], np.float64)
    np.savez_compressed(os.path.join(OUT, "distill.npz"), **g)
    print("wrote", os.path.join(OUT, "distill.npz"), {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
