import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from distillation_trajectories_b200 import grid
dev = torch.device("cuda", 0)
ck = grid.stage_chunk(list(range(592)), bench.Cfg, bench.GUIDANCE, dev)
m = bench.make_model(bench.Cfg, 1.0, 0, dev)
grid.run_chunk(m, [m], ck, dev, "f16")
torch.cuda.synchronize()
