"""GPU (torchrun): BASELINE configs[3] -- the large-batch sweep: 65,536 seeds x 8 guidance scales, 1x16x16, teacher vs one
student, seeds sharded over the ranks, one NCCL all-reduce of the metric sums at the end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_sweep.py [seeds]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from distillation_trajectories_b200 import grid
from distillation_trajectories_b200.models import DiffusionUNet

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    torch.set_num_threads(max(1, (os.cpu_count() or world) // world))
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
models = []
for sf, seed in ((1.0, 0), (0.5, 1050)):
    torch.manual_seed(seed)
    with bench.quiet():
        models.append(DiffusionUNet(bench.Cfg, sf).eval().to(dev))
G = len(bench.GUIDANCE)
grid.sweep(models[0], {"student": models[1]}, bench.Cfg, bench.GUIDANCE, 592 * world, dev, rank, world, max_pairs=592 * G)   # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
stats = {}
res = grid.sweep(models[0], {"student": models[1]}, bench.Cfg, bench.GUIDANCE, n_seeds, dev, rank, world, max_pairs=592 * G, stats=stats)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    traj = 2 * n_seeds * G
    out = {"config": f"configs[3]: {n_seeds} seeds x {G} guidance scales, 1x16x16, teacher vs student sf=0.5, {world} GPU(s)",
           "seconds": dt, "trajectories": traj, "trajectories_per_s": traj / dt, "pairs": n_seeds * G,
           "metrics_at_w": {str(w): {k: res["student"][w][k] for k in ("trajectory_mse", "distribution_similarity",
                                                                       "mean_directional_consistency", "path_length_similarity")}
                            for w in bench.GUIDANCE}}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
