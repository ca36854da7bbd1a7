"""Forward-only half of the training-side steps: q_sample, p_losses (forward value) and the
teacher targets of the distillation step, with a timestep PER ROW.

CPU: the oracle restatement against tests/golden/distill.npz (made by the unmodified reference,
oracle/make_golden_distill.py).  GPU: the product path (libdtraj per-row forward) against fixture and oracle."""
import numpy as np
import pytest
import torch

from oracle import distill as odist
from oracle import samplers as osmp
from helpers import assert_close, cpu_sd, golden_models, load_golden

CASES = ["tiny16", "tiny32"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_distill_matches_reference_fixture(name):
    g, cfg, teacher, _ = golden_models(name)
    d = load_golden("distill")
    sd = cpu_sd(teacher)
    params = osmp.diffusion_params(cfg.timesteps)
    images, t = torch.from_numpy(d[f"{name}/images"]), torch.from_numpy(d[f"{name}/t"])
    torch.manual_seed(78)
    x_noisy, noise, pc, pu = odist.teacher_targets(sd, images, t, params)
    np.testing.assert_array_equal(noise.numpy(), d[f"{name}/noise"])
    np.testing.assert_allclose(x_noisy.numpy(), d[f"{name}/x_noisy"], rtol=0, atol=1e-7)
    assert_close(pc.numpy(), d[f"{name}/pred_cond"], 0.0, 2e-6, "pred_cond")
    assert_close(pu.numpy(), d[f"{name}/pred_uncond"], 0.0, 2e-6, "pred_uncond")
    for tag, cond in (("none", None), ("cond1", torch.ones(images.shape[0], 1))):
        torch.manual_seed(79)
        loss = float(odist.p_losses(sd, images, t, params, cond))
        assert abs(loss - float(d[f"{name}/p_losses_{tag}"][0])) <= 1e-5 * abs(loss)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("tf32x3", 6e-5), ("tf32", 4e-3), ("f16", 4e-3)])
@pytest.mark.parametrize("name", CASES)
def test_teacher_targets_per_row_timesteps(name, prec, tol):
    from distillation_trajectories_b200 import distill, set_precision, umma_error_flag
    from distillation_trajectories_b200.utils import diffusion
    g, cfg, teacher, _ = golden_models(name, device="cuda")
    d = load_golden("distill")
    params = diffusion.get_diffusion_params(cfg.timesteps, cfg)
    images, t = torch.from_numpy(d[f"{name}/images"]), torch.from_numpy(d[f"{name}/t"])
    # the reference run drew its noise on the CPU: draw there, then move (same stream as the fixture)
    torch.manual_seed(78)
    x_ref, n_ref = diffusion.q_sample(images, t, {k: v.cpu() for k, v in params.items()})
    np.testing.assert_array_equal(n_ref.numpy(), d[f"{name}/noise"])
    np.testing.assert_allclose(x_ref.numpy(), d[f"{name}/x_noisy"], rtol=0, atol=1e-7)
    set_precision(prec, "forward")
    try:
        # product forward with a timestep per row, both conditioning variants in one batch
        B = images.shape[0]
        pc = teacher(x_ref.cuda(), t.cuda(), torch.ones(B, 1, device="cuda")).cpu().numpy()
        pu = teacher(x_ref.cuda(), t.cuda(), None).cpu().numpy()
        assert_close(pc, d[f"{name}/pred_cond"], 0.0, tol, f"pred_cond [{prec}]")
        assert_close(pu, d[f"{name}/pred_uncond"], 0.0, tol, f"pred_uncond [{prec}]")
        # the fused helper (one forward of 2B rows) on device-drawn noise: against the oracle on the same x_noisy
        torch.manual_seed(5)
        x_noisy, noise, tc, tu = distill.teacher_targets(teacher, images.cuda(), t.cuda(), params)
        sd = cpu_sd(teacher)
        from oracle import unet as ounet
        want_c = ounet.unet_forward(sd, x_noisy.cpu(), t, torch.ones(B, 1)).numpy()
        want_u = ounet.unet_forward(sd, x_noisy.cpu(), t, None).numpy()
        assert_close(tc.cpu().numpy(), want_c, 0.0, tol, f"teacher_targets cond [{prec}]")
        assert_close(tu.cpu().numpy(), want_u, 0.0, tol, f"teacher_targets uncond [{prec}]")
        # p_losses forward value
        torch.manual_seed(9)
        loss = float(diffusion.p_losses(teacher, images.cuda(), t.cuda(), params, None))
        torch.manual_seed(9)
        xn, nz = diffusion.q_sample(images.cuda(), t.cuda(), params)
        want = float(torch.nn.functional.mse_loss(ounet.unet_forward(sd, xn.cpu(), t, None), nz.cpu()))
        assert abs(loss - want) <= max(10 * tol, 1e-4) * abs(want)
        assert umma_error_flag() == 0
    finally:
        set_precision("tf32x3", "forward")
