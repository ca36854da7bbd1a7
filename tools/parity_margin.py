"""GPU: how much of the stated trajectory tolerance (|d| <= 1e-3 |ref| + 1e-4 max|ref|) each arithmetic mode uses on the
BASELINE-size S2 cases of tests/test_gpu_samplers.py (50 steps, CFG): prints max over elements of |d| / tol."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import samplers as osmp
from distillation_trajectories_b200 import sampling, set_precision
from distillation_trajectories_b200.analysis import trajectory_engine as te
from helpers import Cfg, make_model, oracle_fn

sampling.set_noise_device("cpu")
stack = lambda tr: torch.stack(list(tr)).cpu().numpy().astype(np.float64)
cases = [(1, 16, 1.0, 7.5), (1, 16, 0.5, 20.0), (3, 32, 1.0, 7.5), (3, 32, 0.1, 7.5)]
seeds = [42, 43, 44] if len(sys.argv) < 2 else [int(a) for a in sys.argv[1].split(",")]
for C, H, sf, w in cases:
    cfg = Cfg(C, H, 50)
    model = make_model(cfg, sf, 1000 + int(sf * 100), stress=False, device="cuda")
    for seed in seeds:
        torch.manual_seed(seed)
        noise = torch.randn(1, C, H, H)
        want = stack(osmp.s2_generate_trajectory(oracle_fn(model), noise, 50, seed=seed, guidance_scale=w))
        tol = 1e-3 * np.abs(want) + 1e-4 * np.abs(want).max()
        line = f"{C}x{H} sf={sf} w={w} seed={seed} max|ref|={np.abs(want).max():.2f}:"
        for prec in ("fp32", "tf32x3", "tf32", "f16"):
            set_precision(prec, "S2")
            got = stack(te.generate_trajectory(model, noise, 50, "cuda", seed=seed, guidance_scale=w))
            line += f"  {prec} {np.max(np.abs(got - want) / tol):.3f}"
        print(line, flush=True)
