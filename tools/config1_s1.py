"""GPU: BASELINE configs[0] -- teacher 1x16x16, p_sample_loop (S1), 50 timesteps, batch 64, guidance 1.0 (both forwards
still run, utils/diffusion.py:122-126), trajectory tracked; timed through the public drop-in call, next to the CPU oracle
port of the same loop on the box's host cores (test infrastructure used as a baseline only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import Cfg, make_model, oracle_fn
from oracle import samplers as osmp
from distillation_trajectories_b200 import get_precision
from distillation_trajectories_b200.utils import diffusion

cfg = Cfg(1, 16, 50)
model = make_model(cfg, 1.0, 0, stress=False, device="cuda")
params = diffusion.get_diffusion_params(50, cfg)
for rep in range(3):
    torch.manual_seed(123)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    img, traj = diffusion.p_sample_loop(model, (64, 1, 16, 16), 50, params, device="cuda", config=cfg, track_trajectory=True,
                                        guidance_scale=1.0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"GPU run {rep} [{get_precision('S1')}]: 64 trajectories x 51 frames in {dt*1e3:.1f} ms = {64/dt:.0f} trajectories/s", flush=True)
torch.set_num_threads(os.cpu_count() or 1)
torch.manual_seed(123)
t0 = time.perf_counter()
osmp.s1_p_sample_loop(oracle_fn(model), (64, 1, 16, 16), 50, osmp.diffusion_params(50), 50, 1.0)
dt = time.perf_counter() - t0
print(f"CPU oracle port ({torch.get_num_threads()} threads): {dt:.2f} s = {64/dt:.1f} trajectories/s")
