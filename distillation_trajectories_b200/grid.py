"""(seed x guidance scale x student) sweep: the batched, sharded form of the reference's CFG analysis
(scripts/analysis/analyze_trajectory_metrics.py:476-511 calling compare_trajectories per student).

* every (seed, guidance) of a model is one row group of a captured sampling loop (large batch);
* the teacher trajectory of a (seed, guidance) is generated ONCE and compared with every student
  (the reference regenerates it per student although it only depends on seed and guidance);
* seeds are sharded round-robin over ranks -- the grid is embarrassingly parallel -- and the only
  communication is one all-reduce (sum, f64) of the per-(student, guidance) metric sums and counts
  at the end (NCCL over NVLink when the process group is NCCL; gloo in the CPU tests).

A chunk of the sweep goes through three stages so that the device part can be timed on its own:
``stage_chunk`` (host RNG draws exactly as the reference makes them + host-to-device copies),
``run_chunk`` (device only: sampling loops + metric kernels) and ``finish_chunk`` (device-to-host copy
of the per-frame reductions + the f64 scalar formulas).
"""
import numpy as np
import torch

from . import sampling
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


class Chunk:
    """Device-resident inputs of one batch of (seed, guidance) pairs; pair p = seed-major, guidance-minor."""
    __slots__ = ("samples", "G", "x", "seeds", "ws", "bank", "z_index", "idx", "idx_set", "T", "h2d_bytes")


def stage_chunk(samples, config, guidance_scales, device):
    """Host side of a chunk: x_T per sample from the CPU generator seeded 42 + s
    (analysis/trajectory_engine.py:144-149), the per-step noise bank (every distinct seed + t drawn once
    with the reference's calls), the noise index table and the Wasserstein subsample indices; everything
    is copied to ``device`` from pinned host memory."""
    device = torch.device(device)
    C, H, T = config.channels, config.image_size, config.timesteps
    G = len(guidance_scales)
    ck = Chunk()
    ck.samples, ck.G, ck.T = list(samples), G, T
    noises = []
    for s in ck.samples:
        torch.manual_seed(42 + s)
        np.random.seed(42 + s)
        noises.append(torch.randn(1, C, H, H))
    x = torch.cat(noises).repeat_interleave(G, dim=0)
    ck.seeds = [42 + s for s in ck.samples for _ in range(G)]
    ck.ws = [gs for _ in ck.samples for gs in guidance_scales]
    ndev = sampling.noise_device(device)
    bank, first = te._noise_bank(ck.seeds, (1, C, H, H), ndev, T)
    ts = np.arange(T - 1, 0, -1)
    zi = (np.asarray(ck.seeds)[None, :] + ts[:, None] - first).astype(np.int32) if len(ts) else np.zeros((1, len(ck.seeds)), np.int32)
    L, D = T + 1, C * H * H
    idx = te.wasserstein_index_sets([42 + s for s in ck.samples], T, L, D)
    nbytes = 0

    def up(t):
        nonlocal nbytes
        if t.device == device:
            return t
        nbytes += t.numel() * t.element_size()
        return t.pin_memory().to(device, non_blocking=True)

    ck.x = up(x)
    ck.bank = up(bank)
    ck.z_index = up(torch.from_numpy(zi))
    ck.idx = None if idx is None else up(torch.from_numpy(idx))
    ck.idx_set = None if idx is None else torch.arange(len(ck.samples), dtype=torch.int32, device=device).repeat_interleave(G)
    ck.h2d_bytes = nbytes
    return ck


def run_chunk(teacher_model, student_models, ck, device, precision=None):
    """Device side of a chunk: one captured S2 loop per model over all pairs, then the streaming metric
    kernels for every (teacher, student).  Returns device tensors (red [n_students, N, L, 6],
    w1 [n_students, N, L]) and the number of trajectories generated."""
    device = torch.device(device)
    prec = precision or te.get_precision("S2")

    def gen(model):
        model.eval()
        eng = te.UNetEngine.for_model(model, ck.x.shape[2], ck.T, prec, device)
        return sampling.s2_sample(eng, ck.x, ck.T, ck.ws, ck.bank, ck.z_index)

    tt = gen(teacher_model)
    t_flat = tt.reshape(tt.shape[0], tt.shape[1], -1)
    n_traj = tt.shape[0]
    reds, w1s = [], []
    for sm_model in student_models:
        if sm_model is teacher_model:
            s_flat = t_flat
        else:
            st = gen(sm_model)
            s_flat = st.reshape(st.shape[0], st.shape[1], -1)
            n_traj += st.shape[0]
        reds.append(tm.pair_reductions(t_flat, s_flat))
        w1s.append(tm.wasserstein_frames(t_flat, s_flat, ck.idx, ck.idx_set))
    return torch.stack(reds), torch.stack(w1s), n_traj


def finish_chunk(red, w1, ck, config, sums):
    """Device-to-host copy of the reductions and the f64 scalar formulas; accumulates into ``sums``.
    Returns the bytes copied."""
    H, D = config.image_size, config.channels * config.image_size ** 2
    red_h, w1_h = red.cpu().numpy(), w1.cpu().numpy()
    for i in range(red_h.shape[0]):
        sm = tm.scalar_metrics_batched(red_h[i], w1_h[i], H * H, D)
        for j, k in enumerate(tm.SCALAR_KEYS):
            sums[i, :, j] += sm[k].reshape(len(ck.samples), ck.G).sum(axis=0)
        sums[i, :, -1] += len(ck.samples)
    return red_h.nbytes + w1_h.nbytes


@torch.no_grad()
def sweep(teacher_model, students, config, guidance_scales, num_samples, device=None, rank=0, world_size=1,
          max_pairs=4096, precision=None, reduce=True, stats=None):
    """Run the sweep for ``students`` (dict name -> model) against ``teacher_model``.

    Returns {name: {gs: {18 scalar metrics averaged over all ``num_samples`` seeds}}}; with
    ``reduce=True`` every rank returns the global averages.  ``stats`` (dict) receives counters:
    trajectories generated, pairs measured, bytes moved each way.
    """
    if device is None:
        device = next(teacher_model.parameters()).device
    device = torch.device(device)
    guidance_scales = list(guidance_scales)
    G = len(guidance_scales)
    names = list(students)
    sums = np.zeros((len(names), G, len(tm.SCALAR_KEYS) + 1), np.float64)
    mine = shard_samples(num_samples, rank, world_size)
    per_chunk = max(1, max_pairs // G)
    n_traj = n_pairs = h2d = d2h = 0
    for c0 in range(0, len(mine), per_chunk):
        ck = stage_chunk(mine[c0:c0 + per_chunk], config, guidance_scales, device)
        red, w1, nt = run_chunk(teacher_model, [students[n] for n in names], ck, device, precision)
        d2h += finish_chunk(red, w1, ck, config, sums)
        h2d += ck.h2d_bytes
        n_traj += nt
        n_pairs += len(ck.seeds) * len(names)
    if stats is not None:
        for k, v in (("trajectories", n_traj), ("pairs", n_pairs), ("h2d_bytes", h2d), ("d2h_bytes", d2h)):
            stats[k] = stats.get(k, 0) + v
    if reduce:
        sums = reduce_sums(sums)
    return averages_from_sums(sums, names, guidance_scales)
