"""(seed x guidance scale x student) sweep: the batched, sharded form of the reference's CFG analysis
(scripts/analysis/analyze_trajectory_metrics.py:476-511 calling compare_trajectories per student).

* every (seed, guidance) of a model is one row group of a captured sampling loop (large batch);
* the teacher trajectory of a (seed, guidance) is generated ONCE and compared with every student
  (the reference regenerates it per student although it only depends on seed and guidance);
* seeds are sharded round-robin over ranks -- the grid is embarrassingly parallel -- and the only
  communication is one all-reduce (sum, f64) of the per-(student, guidance) metric sums and counts
  at the end (NCCL over NVLink when the process group is NCCL; gloo in the CPU tests).
"""
import numpy as np
import torch

from .analysis import trajectory_engine as te
from .analysis.metrics import trajectory_metrics as tm


def shard_samples(num_samples, rank, world_size):
    """Round-robin seed ownership: sample s belongs to rank s % world_size."""
    return list(range(rank, num_samples, world_size))


def reduce_sums(sums, group=None):
    """All-reduce (sum) of the f64 partial sums [n_students, n_gs, n_keys + 1] (last slot = count)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(sums)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def averages_from_sums(sums, student_keys, guidance_scales):
    """{student: {gs: {metric: mean over all seeds}}} (analysis/trajectory_engine.py:171-175)."""
    out = {}
    for i, sk in enumerate(student_keys):
        out[sk] = {}
        for g, gs in enumerate(guidance_scales):
            cnt = sums[i, g, -1]
            out[sk][gs] = {k: (float(sums[i, g, j] / cnt) if cnt > 0 else float("nan"))
                           for j, k in enumerate(tm.SCALAR_KEYS)}
    return out


@torch.no_grad()
def sweep(teacher_model, students, config, guidance_scales, num_samples, device=None, rank=0, world_size=1,
          max_pairs=4096, precision=None, reduce=True, stats=None):
    """Run the sweep for ``students`` (dict name -> model) against ``teacher_model``.

    Returns {name: {gs: {18 scalar metrics averaged over all ``num_samples`` seeds}}}; with
    ``reduce=True`` every rank returns the global averages.  ``stats`` (dict) receives counters:
    trajectories generated, pairs measured, kernel launches.
    """
    if device is None:
        device = next(teacher_model.parameters()).device
    device = torch.device(device)
    G = len(guidance_scales)
    names = list(students)
    sums = np.zeros((len(names), G, len(tm.SCALAR_KEYS) + 1), np.float64)
    mine = shard_samples(num_samples, rank, world_size)
    per_chunk = max(1, max_pairs // G)
    C, H, T = config.channels, config.image_size, config.timesteps
    n_traj = n_pairs = 0
    for c0 in range(0, len(mine), per_chunk):
        chunk = mine[c0:c0 + per_chunk]
        noises = []
        for s in chunk:                                   # analysis/trajectory_engine.py:144-149
            torch.manual_seed(42 + s)
            np.random.seed(42 + s)
            noises.append(torch.randn(1, C, H, H))
        x = torch.cat(noises).repeat_interleave(G, dim=0)
        seeds = [42 + s for s in chunk for _ in range(G)]
        ws = [gs for _ in chunk for gs in guidance_scales]
        tt = te.generate_trajectories_batched(teacher_model, x, seeds, ws, T, device, precision)
        t_flat = tt.reshape(tt.shape[0], tt.shape[1], -1)
        n_traj += len(seeds)
        L, D = t_flat.shape[1], t_flat.shape[2]
        idx = te.wasserstein_index_sets([42 + s for s in chunk], T, L, D)
        idx_dev = None if idx is None else torch.from_numpy(idx).to(device)
        idx_set = None if idx is None else torch.arange(len(chunk), dtype=torch.int32, device=device).repeat_interleave(G)
        for i, name in enumerate(names):
            sm_model = students[name]
            if sm_model is teacher_model:
                s_flat = t_flat
            else:
                st = te.generate_trajectories_batched(sm_model, x, seeds, ws, T, device, precision)
                s_flat = st.reshape(st.shape[0], st.shape[1], -1)
                n_traj += len(seeds)
            red = tm.pair_reductions(t_flat, s_flat)
            w1 = tm.wasserstein_frames(t_flat, s_flat, idx_dev, idx_set)
            sm = tm.scalar_metrics_batched(red.cpu().numpy(), w1.cpu().numpy(), H * H, D)
            n_pairs += len(seeds)
            for j, k in enumerate(tm.SCALAR_KEYS):
                sums[i, :, j] += sm[k].reshape(len(chunk), G).sum(axis=0)
            sums[i, :, -1] += len(chunk)
    if stats is not None:
        stats["trajectories"] = stats.get("trajectories", 0) + n_traj
        stats["pairs"] = stats.get("pairs", 0) + n_pairs
    if reduce:
        sums = reduce_sums(sums)
    return averages_from_sums(sums, names, guidance_scales)
