"""CPU, build container only: the oracle restatement against the LIVE reference in /root/reference
(skipped on the GPU box, where the committed golden fixtures stand in)."""
import contextlib
import functools
import io

import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import refload, samplers as osmp, unet

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return refload.load()


def _model(ref, cfg, sf, seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.models.DiffusionUNet(cfg, sf).eval()


@pytest.mark.parametrize("C,H,sf", [(1, 16, 1.0), (3, 32, 0.3), (1, 16, 0.6)])
def test_forward_bit_exact(ref, C, H, sf):
    cfg = refload.RefConfig(channels=C, image_size=H)
    m = _model(ref, cfg, sf, 0)
    x, t = torch.randn(2, C, H, H), torch.tensor([17, 17])
    for cond in (None, torch.ones(2, 1), torch.zeros(2, 1)):
        with torch.no_grad():
            assert torch.equal(m(x, t, cond), unet.unet_forward(m.state_dict(), x, t, cond))


def test_samplers_bit_exact_teacher(ref):
    cfg = refload.RefConfig(channels=1, image_size=16, timesteps=4)
    m = _model(ref, cfg, 1.0, 0)
    f = functools.partial(unet.unet_forward, m.state_dict())
    torch.manual_seed(9)
    with contextlib.redirect_stderr(io.StringIO()):
        _, a = ref.diffusion.p_sample_loop(m, (2, 1, 16, 16), 4, ref.diffusion.get_diffusion_params(4, cfg), device="cpu",
                                           config=cfg, track_trajectory=True, guidance_scale=7.5)
    torch.manual_seed(9)
    _, b = osmp.s1_p_sample_loop(f, (2, 1, 16, 16), 4, osmp.diffusion_params(4), 4, 7.5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    noise = torch.randn(1, 1, 16, 16)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        a = ref.trajectory_engine.generate_trajectory(m, noise, 4, "cpu", seed=1, guidance_scale=20.0)
    b = osmp.s2_generate_trajectory(f, noise, 4, seed=1, guidance_scale=20.0)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_metrics_match_reference(ref):
    rng = np.random.RandomState(0)
    T = [torch.from_numpy(rng.randn(1, 3, 32, 32).astype(np.float32)) for _ in range(9)]
    S = [t + 0.05 * torch.from_numpy(rng.randn(1, 3, 32, 32).astype(np.float32)) for t in T]
    np.random.seed(4)
    a = ref.trajectory_metrics.compute_trajectory_metrics(T, S)
    np.random.seed(4)
    b = om.trajectory_metrics(T, S)
    assert list(a) == list(b)
    for k in a:
        np.testing.assert_allclose(np.asarray(a[k], np.float64), np.asarray(b[k], np.float64), rtol=1e-10, atol=1e-12,
                                   equal_nan=True, err_msg=k)
        assert type(a[k]) is type(b[k]), k
