"""Oracle: CPU restatement of the reference's trajectory metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py).

  Q1  ``compute_trajectory_metrics`` ........ /root/reference/analysis/metrics/trajectory_metrics.py:12-325
  Q2  ``analyze_time_dependent_distances`` .. /root/reference/analysis/metrics/time_dependent.py:42-120
  Q3  ``transform_metrics`` ................. /root/reference/utils/metric_transformations.py:3-38

Per-frame reductions are fp32 torch reductions over ALL elements of the frame tensor
(as the reference's ``torch.norm`` / ``torch.mean``); per-trajectory scalars are f64
Python/numpy arithmetic.  The 1-d Wasserstein distance is scipy's
(``scipy.stats.wasserstein_distance``, pinned scipy==1.13.1 in the reference's
requirements.txt:43; not vendored): for two equal-size, equal-weight samples it is
the mean absolute difference of the sorted samples, restated here in f64.
"""
import numpy as np
import torch


def wasserstein_1d(u, v):
    """W1 between the empirical distributions of u and v (any sizes), f64.
    Published algorithm of scipy.stats._cdf_distance(p=1): integrate |U_cdf - V_cdf|
    over the merged support."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    if u.size == v.size:
        return float(np.mean(np.abs(np.sort(u) - np.sort(v))))
    allv = np.sort(np.concatenate([u, v]))
    deltas = np.diff(allv)
    ucdf = np.searchsorted(np.sort(u), allv[:-1], side="right") / u.size
    vcdf = np.searchsorted(np.sort(v), allv[:-1], side="right") / v.size
    return float(np.sum(np.abs(ucdf - vcdf) * deltas))


def _images(traj):
    # trajectory_metrics.py:29-37 -- S3 stores (tensor, t) tuples
    return [it[0] for it in traj] if isinstance(traj[0], tuple) else list(traj)


def _ratio(a, b):
    hi = max(a, b)
    return min(a, b) / hi if hi > 0 else 1.0


def trajectory_metrics(teacher_trajectory, student_trajectory, rng=None, sample_size=1000):
    """Q1.  ``rng`` is the numpy RandomState-like object whose ``choice`` is used for the
    Wasserstein subsample (the reference uses the GLOBAL numpy RNG,
    trajectory_metrics.py:306); default is the ``np.random`` module itself."""
    rng = np.random if rng is None else rng
    T = _images(teacher_trajectory)
    S = _images(student_trajectory)
    if T[-1].shape != S[-1].shape and T[-1].shape[2:] != S[-1].shape[2:]:
        # trajectory_metrics.py:40-52
        S = [torch.nn.functional.interpolate(s, size=T[0].shape[2:], mode="bilinear",
                                             align_corners=True) for s in S]
    m = {}
    n = min(len(T), len(S))
    pixels = T[0].shape[2] * T[0].shape[3]

    m["endpoint_distance"] = torch.norm(T[-1] - S[-1]).item()           # :55
    mse = torch.mean((T[-1] - S[-1]) ** 2).item()                       # :59
    m["mse"] = mse

    acc = 0.0                                                           # :63-86
    for i in range(n):
        acc += torch.mean((T[i] - S[i]) ** 2).item()
    acc = acc / n * 1000
    m["trajectory_mse"] = np.log1p(1.0 - acc)

    pos = [torch.norm(T[i] - S[i]).item() for i in range(n)]            # :90-101, :180-187
    m["point_by_point_similarity"] = np.exp(-5.0 * (np.mean(pos) if pos else float("inf")))
    m["log_mse_similarity"] = max(0, 1.0 - np.log1p(mse * 5000) / np.log1p(5000))   # :106-108

    tl = sl = 0                                                         # :111-131
    for i in range(1, n):
        tl += torch.norm(T[i] - T[i - 1]).item() / pixels
        sl += torch.norm(S[i] - S[i - 1]).item() / pixels
    tl /= (n - 1)
    sl /= (n - 1)
    m["teacher_path_length"] = tl
    m["student_path_length"] = sl
    m["path_length_similarity"] = np.log1p(_ratio(tl, sl))              # :134-137

    te = torch.norm(T[-1] - T[0]).item()                                # :140-153
    se = torch.norm(S[-1] - S[0]).item()
    teff = te / tl if tl > 0 else 0
    seff = se / sl if sl > 0 else 0
    m["teacher_efficiency"] = teff
    m["student_efficiency"] = seff
    m["efficiency_similarity"] = np.log1p(_ratio(teff, seff))

    tv = [torch.norm(T[i] - T[i - 1]).item() for i in range(1, len(T))]  # :156-177
    sv = [torch.norm(S[i] - S[i - 1]).item() for i in range(1, len(S))]
    m["teacher_velocities"] = tv
    m["student_velocities"] = sv
    vs = [_ratio(a, b) for a, b in zip(tv, sv)]
    m["velocity_similarities"] = vs
    m["mean_velocity_similarity"] = np.mean(vs) if vs else 0.0

    m["position_differences"] = pos
    m["mean_position_difference"] = np.mean(pos) if pos else 0.0
    m["max_position_difference"] = np.max(pos) if pos else 0.0

    dc, wdc = [], []                                                    # :190-231
    for i in range(n - 1):
        dt = T[i + 1] - T[i]
        ds = S[i + 1] - S[i]
        nt, ns = torch.norm(dt), torch.norm(ds)
        if nt > 0 and ns > 0:
            cos = (torch.sum(dt.flatten() * ds.flatten())
                   / (torch.norm(dt.flatten()) * torch.norm(ds.flatten()))).item()
            dc.append(cos)
            wdc.append(cos * (nt.item() + ns.item()) / 2)
    m["directional_consistency"] = dc
    m["mean_directional_consistency"] = np.mean(dc) if dc else 0.0
    if wdc:
        tw = sum((tv[i] + sv[i]) / 2 for i in range(min(len(tv), len(sv))))
        wm = sum(wdc) / tw if tw > 0 else 0
        m["weighted_directional_consistency"] = wm ** 2
    else:
        m["weighted_directional_consistency"] = 0.0

    tf = [t.flatten().cpu().numpy() for t in T]                          # :239-293
    sf = [s.flatten().cpu().numpy() for s in S]
    if len(tf) != len(sf):
        longer, shorter = (tf, sf) if len(tf) > len(sf) else (sf, tf)
        lt = np.linspace(0, 1, len(longer))
        st = np.linspace(0, 1, len(shorter))
        stack = np.stack(longer).astype(np.float64)            # interp1d works in f64
        res = np.stack([np.interp(st, lt, stack[:, d]) for d in range(stack.shape[1])], axis=1)
        res = [r for r in res]
        ta, sa = (res, shorter) if len(tf) > len(sf) else (shorter, res)
    else:
        ta, sa = tf, sf
    dist = [np.linalg.norm(a - b) for a, b in zip(ta, sa)]
    m["path_alignment"] = np.exp(-10.0 * np.sum(dist) / len(dist))

    wd = []                                                             # :296-323
    for a, b in zip(tf, sf):
        idx = rng.choice(len(a), min(sample_size, len(a)), replace=False)
        wd.append(wasserstein_1d(a[idx], b[idx]))
    m["wasserstein_distances"] = wd
    m["mean_wasserstein"] = np.mean(wd)
    m["distribution_similarity"] = np.log1p(np.exp(-m["mean_wasserstein"]))
    return m


def time_dependent_distances(teacher_trajectories, student_trajectories, size_factor=None):
    """Q2 (compute part only; the plot at time_dependent.py:122-150 is out of scope)."""
    res = {"teacher_distances": [], "student_distances": [], "teacher_avg_distance": 0,
           "student_avg_distance": 0, "teacher_std_distance": 0, "student_std_distance": 0,
           "size_factor": size_factor}
    if not teacher_trajectories or not student_trajectories:
        return res

    def per_traj(trajs):
        out = []
        for tr in trajs:
            im = _images(tr)
            d = [torch.norm(im[i] - im[i - 1]).item() for i in range(1, len(im))]
            if d:
                out.append(d)
        return out

    for who, trajs in (("teacher", teacher_trajectories), ("student", student_trajectories)):
        res[who + "_distances"] = per_traj(trajs)
    avg = {"teacher": [], "student": []}
    if res["teacher_distances"] and res["student_distances"]:
        for who in avg:
            dd = res[who + "_distances"]
            for t in range(min(len(d) for d in dd)):
                avg[who].append(sum(d[t] for d in dd) / len(dd))
    for who in avg:
        a = avg[who]
        res[who + "_avg_per_timestep"] = a
        res[who + "_avg_distance"] = sum(a) / len(a) if a else 0
        if a:
            mu = res[who + "_avg_distance"]
            res[who + "_std_distance"] = (sum((d - mu) ** 2 for d in a) / len(a)) ** 0.5
    return res


def transform_metrics(path_length_similarity, trajectory_mse, directional_consistency,
                      distribution_similarity):
    """Q3."""
    mse = np.log1p(np.clip(trajectory_mse, 0, None))
    ds = np.log1p(distribution_similarity)
    return {"path_length_similarity": path_length_similarity,
            "trajectory_mse": np.clip(1 - mse / np.log1p(1.0), 0, 1),
            "mean_directional_consistency": np.abs(directional_consistency),
            "distribution_similarity": np.clip(ds / np.log1p(1.0), 0, 1)}


def average_scalar_metrics(per_sample):
    """analysis/trajectory_engine.py:171-175: mean over samples of every key whose value
    is a Python int/float (np.float64 qualifies, np.float32 and lists do not)."""
    out = {}
    for key, v in per_sample[0].items():
        if isinstance(v, (int, float)) and not isinstance(v, bool):
            out[key] = sum(m[key] for m in per_sample) / len(per_sample)
    return out
