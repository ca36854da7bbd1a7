"""Trajectory generation S2 + teacher/student comparison (drop-in for
/root/reference/analysis/trajectory_engine.py:24-180).

``generate_trajectory`` keeps the reference's batch-1 signature; ``compare_trajectories`` runs
every (sample, guidance scale) of a model as ONE batched captured loop (per-sample guidance
weight, per-sample noise index) and computes all pair metrics with the streaming kernels.
Noise is drawn with the reference's own calls (``torch.manual_seed(seed + t)`` then a draw on the
model's device), so trajectories agree with the reference run on the same device type.
"""
import numpy as np
import torch

from .. import sampling
from ..engine import UNetEngine, check_device_errors, get_precision
from .metrics import trajectory_metrics as tm


def extract(a, t, x_shape):
    """analysis/trajectory_engine.py:14-22 (unused duplicate of utils.diffusion.extract, kept for API parity)."""
    out = a.gather(-1, torch.clamp(t, 0, a.shape[0] - 1))
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def _draw_step_noise(shape, device, timesteps, seed):
    """The per-step draws of analysis/trajectory_engine.py:86-95 for one trajectory, t = T-1 .. 1."""
    zs = []
    for t in range(timesteps - 1, 0, -1):
        if seed is not None:
            torch.manual_seed(seed + t)
            np.random.seed(seed + t)
        zs.append(torch.randn(shape, device=device))
    return zs


@torch.no_grad()
def generate_trajectory(model, noise, timesteps, device, seed=None, guidance_scale=None):
    """One denoising trajectory from ``noise`` [1, C, H, W]: list of T+1 CPU tensors (frame 0 = noise,
    last frame duplicated because no update happens at t = 0)."""
    model.eval()
    device = torch.device(device)
    x = noise.clone().to(device)
    if x.shape[0] != 1 and guidance_scale is not None and guidance_scale > 1.0:
        raise ValueError("the reference's CFG branch only works for batch 1 (trajectory_engine.py:68-76)")
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    eng = UNetEngine.for_model(model, x.shape[2], timesteps, get_precision("S2"), device)
    B = x.shape[0]
    zs = _draw_step_noise(tuple(x.shape), sampling.noise_device(device), timesteps, seed)
    n = len(zs)
    if n:
        bank = torch.stack(zs).reshape(n * B, -1)            # step-major, then sample
        z_index = (np.arange(n)[:, None] * B + np.arange(B)[None, :]).astype(np.int32)
    else:
        bank, z_index = torch.zeros(1, x[0].numel(), device=device), np.zeros((1, B), np.int32)
    traj = sampling.s2_sample(eng, x, timesteps, [guidance_scale] * B, bank, z_index)
    return sampling.frames_to_cpu_list(traj)


_generators = {}


def _generator(device):
    """A private generator per device.  ``g.manual_seed(k); torch.randn(..., generator=g)`` yields exactly the
    stream of ``torch.manual_seed(k); torch.randn(...)`` on that device (same engine, same seeding) without
    re-seeding every CUDA device's global generator for each of the hundreds of keys a batch needs."""
    key = str(device)
    if key not in _generators:
        _generators[key] = torch.Generator(device=device)
    return _generators[key]


def _noise_bank(base_seeds, shape, device, timesteps):
    """z(sample, t) depends on seed + t only: draw each distinct key once.
    Returns (bank [n_keys, D], first_key)."""
    lo, hi = min(base_seeds) + 1, max(base_seeds) + timesteps - 1
    g = _generator(device)
    rows = []
    for key in range(lo, hi + 1):
        g.manual_seed(key)
        rows.append(torch.randn(shape, device=device, generator=g).reshape(-1))
    if not rows:
        return torch.zeros(1, int(np.prod(shape)), device=device), lo
    return torch.stack(rows), lo


@torch.no_grad()
def generate_trajectories_batched(model, noises, seeds, guidance, timesteps, device, precision=None, groups=None):
    """S2 for B independent (noise, seed, w) triples in one captured loop.
    noises [B, C, H, W]; seeds list[int]; guidance list[float|None]; ``groups``: optional group id per triple --
    triples of a group have IDENTICAL noise rows, which lets the first step share forward rows.
    Returns DEVICE [B, T+1, C, H, W]."""
    model.eval()
    device = torch.device(device)
    eng = UNetEngine.for_model(model, noises.shape[2], timesteps, precision or get_precision("S2"), device)
    bank, first = _noise_bank(seeds, (1,) + tuple(noises.shape[1:]), sampling.noise_device(device), timesteps)
    ts = np.arange(timesteps - 1, 0, -1)
    z_index = (np.asarray(seeds)[None, :] + ts[:, None] - first).astype(np.int32) if len(ts) else np.zeros((1, len(seeds)), np.int32)
    return sampling.s2_sample(eng, noises.to(device), timesteps, list(guidance), bank, z_index, groups=groups)


def wasserstein_index_sets(seeds, timesteps, n_frames, numel, sample_size=1000):
    """The np.random.choice subsamples compute_trajectory_metrics would draw after the student's
    generate_trajectory left the global numpy RNG at seed + 1 (trajectory_engine.py:91-93 at t = 1;
    trajectory_metrics.py:301-306).  None when every element is used (numel <= 1000)."""
    K = min(sample_size, numel)
    if K == numel:
        return None
    out = np.empty((len(seeds), n_frames, K), np.int32)

    for i, s in enumerate(seeds):
        rs = np.random.RandomState(s + 1 if timesteps > 1 else s)
        for f in range(n_frames):
            out[i, f] = rs.choice(numel, K, replace=False)
    return out


def wasserstein_index_sets_device(seeds, timesteps, n_frames, numel, device, sample_size=1000):
    """The same index sets drawn ON THE DEVICE (dtraj_numpy_choice_sets: numpy's legacy MT19937 + Fisher-Yates stream
    reproduced bit for bit, one thread per seed): int32 device tensor [n_seeds, n_frames, K], or None when every element
    is used.  The host loop above costs ~90 us per frame (2.4 s per 512 seeds at 3x32x32 -- more than the chunk's device
    time) and its 100 MB result would have to cross PCIe; numpy's shuffle does not scale over host threads."""
    from .. import _lib
    K = min(sample_size, numel)
    if K == numel:
        return None
    device = torch.device(device)
    sd = torch.tensor([(s + 1 if timesteps > 1 else s) & 0xFFFFFFFF for s in seeds], dtype=torch.int64).to(torch.int32)
    sd = sd.pin_memory().to(device, non_blocking=True) if device.type == "cuda" else sd
    out = torch.empty(len(seeds), n_frames, K, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().dtraj_numpy_choice_sets(_lib.ptr(sd), len(seeds), n_frames, numel, K, _lib.ptr(out), _lib.stream_ptr()))
    return out


@torch.no_grad()
def compare_trajectories_batched(teacher_model, student_model, config, guidance_scales, num_samples, device=None,
                                 first_sample=0, sample_stride=1, precision=None):
    """All (sample, guidance scale) pairs of ``compare_trajectories`` at once.
    Samples are first_sample, first_sample + stride, ... (< num_samples) so ranks can shard them.
    Returns (per-pair scalar metrics dict key -> [n_local_samples, n_gs], sample indices)."""
    if device is None:
        device = next(teacher_model.parameters()).device
    device = torch.device(device)
    samples = list(range(first_sample, num_samples, sample_stride))
    C, H, T = config.channels, config.image_size, config.timesteps
    G = len(guidance_scales)
    if not samples:
        return {k: np.zeros((0, G)) for k in tm.SCALAR_KEYS}, samples
    noises = []
    for s in samples:                                   # trajectory_engine.py:144-149
        torch.manual_seed(42 + s)
        np.random.seed(42 + s)
        noises.append(torch.randn(1, C, H, H))
    x = torch.cat(noises).repeat_interleave(G, dim=0)   # pair p = sample-major, guidance-minor
    seeds = [42 + s for s in samples for _ in range(G)]
    ws = [gs for _ in samples for gs in guidance_scales]
    groups = [i for i in range(len(samples)) for _ in range(G)]      # every scale of a sample starts from the same noise
    tt = generate_trajectories_batched(teacher_model, x, seeds, ws, T, device, precision, groups)
    t_flat = tt.reshape(tt.shape[0], tt.shape[1], -1)
    if student_model is teacher_model:
        s_flat = t_flat
    else:
        t_flat = t_flat.clone()                         # the sampler's buffer is reused by the next run
        st = generate_trajectories_batched(student_model, x, seeds, ws, T, device, precision, groups)
        s_flat = st.reshape(st.shape[0], st.shape[1], -1)
    red = tm.pair_reductions(t_flat, s_flat).cpu().numpy()
    L, D = t_flat.shape[1], t_flat.shape[2]
    idx = wasserstein_index_sets([42 + s for s in samples], T, L, D)
    if idx is None:
        w1 = tm.wasserstein_frames(t_flat, s_flat)
    else:
        idx_set = torch.arange(len(samples), dtype=torch.int32).repeat_interleave(G)
        w1 = tm.wasserstein_frames(t_flat, s_flat, torch.from_numpy(idx), idx_set)
    sm = tm.scalar_metrics_batched(red, w1.cpu().numpy(), H * H, D)
    check_device_errors()
    return {k: sm[k].reshape(len(samples), G) for k in tm.SCALAR_KEYS}, samples


def compare_trajectories(teacher_model, student_model, config, guidance_scales=[1.0, 3.0, 5.0], size_factor=1.0,
                         num_samples=3):
    """Teacher-vs-student metrics per guidance scale, averaged over ``num_samples`` seeds 42, 43, ...
    (analysis/trajectory_engine.py:117-180).  Returns {'teacher_metrics': {gs: {18 scalars}},
    'student_metrics': same} -- the two entries hold the same numbers, as in the reference (:162-164)."""
    from .. import grid
    res = grid.sweep(teacher_model, {"student": student_model}, config, list(guidance_scales), num_samples,
                     reduce=False)["student"]
    return {"teacher_metrics": {gs: dict(res[gs]) for gs in guidance_scales},
            "student_metrics": {gs: dict(res[gs]) for gs in guidance_scales}}
