"""GPU: per-model, per-kernel-class device time of one 50-step loop at the bench batch (TrajectorySampler.profile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillation_trajectories_b200 import grid
from distillation_trajectories_b200.engine import UNetEngine
from distillation_trajectories_b200.models import DiffusionUNet
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 296
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
dev = torch.device("cuda", 0)
models = []
for sf, seed in ((1.0, 0), (0.5, 1050)):
    torch.manual_seed(seed)
    with bench.quiet():
        models.append(DiffusionUNet(bench.Cfg, sf).eval().to(dev))
ck = grid.stage_chunk(list(range(seeds)), bench.Cfg, bench.GUIDANCE, dev)
grid.run_chunk(models[0], [models[1]], ck, dev, prec)
torch.cuda.synchronize()
for name, m in zip(("teacher", "student"), models):
    s = next(iter(UNetEngine.for_model(m, 16, 50, prec, dev)._samplers.values()))
    p = s.profile(); p = s.profile()
    n = s.n_updates
    print(f"{name}: per forward (us): conv {p['ms'][0]/n*1e3:.0f} ({p['conv_flops']/p['ms'][0]/1e9:.0f} TF/s)  first {p['ms'][1]/n*1e3:.0f}  "
          f"resample {p['ms'][2]/n*1e3:.0f}  step {p['ms'][3]/n*1e3:.0f}  enc1 {p['ms'][4]/n*1e3:.0f} "
          f"({(p['enc1_flops']/p['ms'][4]/1e9) if p['ms'][4] else 0:.0f} TF/s)   launches {p['launches']}")
