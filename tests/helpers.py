"""Shared test helpers: tiny configs, seeded models, golden fixtures."""
import contextlib
import functools
import io
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Cfg:
    """Duck-typed config (config/config.py:5-95): only what the hot path reads."""

    def __init__(self, channels=1, image_size=16, timesteps=6, **kw):
        self.channels, self.image_size, self.timesteps = channels, image_size, timesteps
        self.sample_steps = self.teacher_steps = self.student_steps = timesteps
        self.beta_start, self.beta_end, self.dropout = 1e-4, 0.02, 0.3
        self.force_cpu, self.mps_enabled = False, False
        self.trajectory_dir = "/tmp/dtraj_test_trajectories"
        for k, v in kw.items():
            setattr(self, k, v)


def bn_stress(model, seed):
    """Same recipe as oracle/make_golden.py::bn_stress (kept in sync by the weight checksum test)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            n = m.num_features
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(n, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(n, generator=g) * 0.1)


def make_model(cfg, sf, seed, stress=True, device="cpu"):
    from distillation_trajectories_b200.models import DiffusionUNet
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = DiffusionUNet(cfg, sf).eval()
    if stress:
        bn_stress(m, seed + 1)
    return m.to(device)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_models(name, device="cpu"):
    """(cfg, teacher, student) of a golden case, rebuilt from seeds; weights are checked against the
    fixture (exactly when stored, by checksum otherwise)."""
    g = load_golden(name)
    C, H, T = (int(v) for v in g["meta"])
    cfg = Cfg(C, H, T)
    sf_t, sf_s = (float(v) for v in g["sf"])
    teacher = make_model(cfg, sf_t, 100)
    student = make_model(cfg, sf_s, 200)
    for who, m in (("teacher", teacher), ("student", student)):
        sd = m.state_dict()
        stored = {k[len(who) + 1:]: v for k, v in g.items() if k.startswith(who + "/")}
        if stored:
            merged = dict(sd)
            merged.update({k: torch.from_numpy(v) for k, v in stored.items()})
            m.load_state_dict(merged, strict=True)
        wsum = float(sum(v.double().sum() for v in m.state_dict().values()))
        assert abs(wsum - float(g[who + "_wsum"][0])) <= 1e-6 * max(1.0, abs(wsum)), \
            f"{name}/{who}: seeded weights drifted from the fixture (torch RNG changed?)"
    return g, cfg, teacher.to(device), student.to(device)


def cpu_sd(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


def oracle_fn(model):
    from oracle import unet
    return functools.partial(unet.unet_forward, cpu_sd(model))


def assert_close(a, b, rtol, atol_frac, what=""):
    """|a-b| <= rtol*|b| + atol_frac*max|b| elementwise (b = reference)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    tol = rtol * np.abs(b) + atol_frac * max(np.abs(b).max(), 1e-30)
    bad = np.abs(a - b) > tol
    assert not bad.any(), (f"{what}: {bad.sum()}/{bad.size} outside rtol={rtol} atol={atol_frac}*max; "
                           f"max|d|={np.abs(a - b).max():.3e} max|ref|={np.abs(b).max():.3e}")
