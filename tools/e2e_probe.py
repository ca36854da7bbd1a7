"""GPU: host time of the three phases of one sweep chunk (stage / run / finish) at the bench batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from distillation_trajectories_b200 import grid
from distillation_trajectories_b200.models import DiffusionUNet
from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm
dev = torch.device("cuda", 0)
models = []
for sf, seed in ((1.0, 0), (0.5, 1050)):
    torch.manual_seed(seed)
    with bench.quiet():
        models.append(DiffusionUNet(bench.Cfg, sf).eval().to(dev))
S = 296
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); ck = grid.stage_chunk(list(range(rep * S, rep * S + S)), bench.Cfg, bench.GUIDANCE, dev); t1 = time.perf_counter(); t1s = T()
    red, w1, n = grid.run_chunk(models[0], [models[1]], ck, dev, "f16"); t2 = time.perf_counter(); t2s = T()
    sums = np.zeros((1, 8, len(tm.SCALAR_KEYS) + 1))
    grid.finish_chunk(red, w1, ck, bench.Cfg, sums); t3 = T()
    print(f"stage host {1e3*(t1-t0):.1f} ms (+sync {1e3*(t1s-t1):.1f}); run_chunk host {1e3*(t2-t1s):.1f} ms, device done after {1e3*(t2s-t1s):.1f}; finish {1e3*(t3-t2s):.1f} ms")
