"""Small profiling target: one device-resident step of the bench workload (teacher + student S2 loops +
metric kernels) at a reduced seed count.  Used under ncu for the launch list and the --set full capture."""
import os
import sys

os.environ.setdefault("DTRAJ_OVERLAP_MODELS", "0")     # one stream: a deterministic launch order for --launch-skip

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from distillation_trajectories_b200 import grid
from distillation_trajectories_b200.models import DiffusionUNet

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
use_graph = (sys.argv[3] != "eager") if len(sys.argv) > 3 else True
dev = torch.device("cuda", 0)
models = []
for sf, seed in ((1.0, 0), (0.5, 1050)):
    torch.manual_seed(seed)
    with bench.quiet():
        models.append(DiffusionUNet(bench.Cfg, sf).eval().to(dev))
ck = grid.stage_chunk(list(range(seeds)), bench.Cfg, bench.GUIDANCE, dev)
red, w1, n = grid.run_chunk(models[0], [models[1]], ck, dev, prec)
torch.cuda.synchronize()
print("trajectories", n, "red", tuple(red.shape), float(red.sum()))
