"""GPU: BASELINE configs[2]-shaped run (3x32x32, w = 7.5, several student sizes vs the teacher) through grid.sweep:
throughput and a parity spot-check of one (seed, student) against the CPU oracle."""
import contextlib, functools, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distillation_trajectories_b200 import grid, sampling
from distillation_trajectories_b200.models import DiffusionUNet

class Cfg:
    channels, image_size, timesteps, dropout = 3, 32, 50, 0.3

dev = torch.device("cuda", 0)
sizes = [float(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.1, 0.3, 0.5, 1.0]
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 64
def mk(sf, seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return DiffusionUNet(Cfg, sf).eval().to(dev)
teacher = mk(1.0, 0)
students = {f"sf{sf}": mk(sf, 1000 + int(sf * 100)) for sf in sizes}
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stats = {}
    res = grid.sweep(teacher, students, Cfg, [7.5], seeds, dev, max_pairs=seeds, stats=stats)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"run {it}: {stats['trajectories']} trajectories in {dt:.2f} s = {stats['trajectories']/dt:.0f} traj/s; "
          f"trajectory_mse sf{sizes[0]} = {res['sf%s' % sizes[0]][7.5]['trajectory_mse']:.5f}", flush=True)
