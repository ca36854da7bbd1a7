"""GPU: small-batch latency of the two reference-shaped calls (BASELINE configs[0] and the batch-1 trajectory):
    p_sample_loop(teacher, (64,1,16,16), 50 steps, track_trajectory=True)   [S1, default tf32x3]
    generate_trajectory(teacher, noise[1,1,16,16], 50, seed, w=7.5)         [S2, default f16]
Run with DTRAJ_PDL=0/1 to compare programmatic dependent launch inside the captured loops."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillation_trajectories_b200.analysis import trajectory_engine as te
from distillation_trajectories_b200.utils import diffusion

dev = torch.device("cuda", 0)
teacher = bench.make_model(bench.Cfg, 1.0, 0, dev)
params = diffusion.get_diffusion_params(50, bench.Cfg)


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


s1 = timed(lambda: diffusion.p_sample_loop(teacher, (64, 1, 16, 16), 50, params, device=dev, config=bench.Cfg, track_trajectory=True,
                                           guidance_scale=1.0), 10)
torch.manual_seed(42)
noise = torch.randn(1, 1, 16, 16)
b1 = timed(lambda: te.generate_trajectory(teacher, noise, 50, dev, seed=42, guidance_scale=7.5), 20)
print(f"DTRAJ_PDL={os.environ.get('DTRAJ_PDL', 'auto')}: S1 batch 64: {s1 * 1e3:.2f} ms per loop = {64 / s1:.0f} trajectories/s;  "
      f"batch-1 S2: {b1 * 1e3:.2f} ms per trajectory")
