"""GPU (optionally torchrun): BASELINE configs[4] -- the metric-kernel bandwidth test over the WHOLE synthetic pair
[N = 262144, T = 50, 3, 32, 32] fp32 (2 x 150 GiB): N is sharded over the ranks and processed in chunks generated on the
device (teacher = x0 + 0.1 cumsum(N(0,1)), student = teacher + 0.05 N(0,1)); only the streaming pair
reductions (path length + directional consistency + MSE sums) are timed, with CUDA events, max over ranks.

    python [-m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1] tools/config5_metrics.py [N] [chunk]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
L, D = 50, 3 * 32 * 32
mine = range(rank * (N // world), (rank + 1) * (N // world), chunk)
ms, sums = 0.0, torch.zeros(6, dtype=torch.float64, device=dev)
for c, n0 in enumerate(mine):
    n = min(chunk, (rank + 1) * (N // world) - n0)
    gen = torch.Generator(device=dev).manual_seed(1234 + n0 // chunk)
    t = torch.randn(n, 1, D, device=dev, generator=gen) + 0.1 * torch.cumsum(torch.randn(n, L, D, device=dev, generator=gen), dim=1)
    s = t + 0.05 * torch.randn(n, L, D, device=dev, generator=gen)
    if c == 0:
        tm.pair_reductions(t, s)                      # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    red = tm.pair_reductions(t, s)
    e1.record()
    torch.cuda.synchronize()
    ms += e0.elapsed_time(e1)
    sums += red.double().sum(dim=(0, 1))
    del t, s, red
tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(sums)
if rank == 0:
    gb = 2.0 * N * L * D * 4 / 1e9
    print(json.dumps({"config": f"configs[4]: [{N}, {L}, 3, 32, 32] x 2 fp32, {world} GPU(s), chunks of {chunk}",
                      "algorithmic_GB": gb, "kernel_ms_max_over_ranks": float(tmax.item()),
                      "GB_per_s": gb / float(tmax.item()) * 1e3, "GB_per_s_per_gpu": gb / float(tmax.item()) * 1e3 / world,
                      "checksum_of_sums": [float(v) for v in sums.tolist()]}))
if world > 1:
    dist.destroy_process_group()
