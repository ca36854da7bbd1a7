"""CPU, world_size 2 over gloo: seed sharding + the single all-reduce of metric sums (grid.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distillation_trajectories_b200 import grid
from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_metric(sample, g, j):
    return (sample + 1) * 0.5 + g * 0.01 + j


def _worker(rank, world, port, n_samples, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nk = len(tm.SCALAR_KEYS)
    sums = np.zeros((2, 3, nk + 1))
    for s in grid.shard_samples(n_samples, rank, world):
        for i in range(2):
            for g in range(3):
                sums[i, g, :nk] += [_fake_metric(s, g, j) * (i + 1) for j in range(nk)]
                sums[i, g, -1] += 1
    total = grid.reduce_sums(sums)
    if rank == 0:
        np.save(out, total)
    dist.destroy_process_group()


def test_two_rank_reduce(tmp_path):
    n = 7
    out = str(tmp_path / "total.npy")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    total = np.load(out)
    assert np.all(total[:, :, -1] == n)
    avg = grid.averages_from_sums(total, ["a", "b"], [1.0, 3.0, 7.5])
    for i, name in enumerate(["a", "b"]):
        for g, gs in enumerate([1.0, 3.0, 7.5]):
            for j, k in enumerate(tm.SCALAR_KEYS):
                want = np.mean([_fake_metric(s, g, j) * (i + 1) for s in range(n)])
                assert abs(avg[name][gs][k] - want) < 1e-12


def test_single_process_reduce_is_identity():
    x = np.arange(6.0).reshape(1, 2, 3)
    assert grid.reduce_sums(x) is x
