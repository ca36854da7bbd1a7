"""Generate tests/golden/batch_metrics.npz by running the UNMODIFIED reference's TrajectoryManager on CPU
(utils/trajectory_manager.py:207-263 generate_and_save_trajectories, :434-548 compute_trajectory_metrics_batch).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python -m oracle.make_golden_batch

Per case the fixture holds the trajectories the reference pickled (so that a test can rebuild the very same pickle
files without the reference) and every list / ``*_avg`` value its ``compute_trajectory_metrics_batch`` returned.  The
global numpy RNG is seeded right before that call because ``compute_trajectory_metrics`` subsamples frames larger than
1000 elements with ``np.random.choice`` (analysis/metrics/trajectory_metrics.py:301-306).
"""
import os
import tempfile

import numpy as np
import torch

from . import refload
from .make_golden import OUT, make_model, quiet

NP_SEED = 7
CASES = [("b16", 1, 16, 6, 1.0, 0.5, 6, 3), ("b32", 3, 32, 4, 0.2, 0.05, 2, 2)]   # name, C, H, T, sf_t, sf_s, student_steps, n


def main():
    ref = refload.load()
    g = {}
    for name, C, H, T, sf_t, sf_s, ss, n in CASES:
        with tempfile.TemporaryDirectory() as d:
            cfg = refload.RefConfig(C, H, T, trajectory_dir=d, student_steps=ss)
            teacher = make_model(ref, cfg, sf_t, 100)
            student = make_model(ref, cfg, sf_s, 200)
            with quiet():
                mgr = ref.trajectory_manager.TrajectoryManager(teacher, student, cfg, size_factor=sf_s)
                mgr.generate_and_save_trajectories(n)
                tt, st = mgr.load_trajectories()
                np.random.seed(NP_SEED)
                allm = mgr.compute_trajectory_metrics_batch(batch_size=2)
        g[f"{name}/meta"] = np.array([C, H, T, ss, n], np.int64)
        g[f"{name}/sf"] = np.array([sf_t, sf_s], np.float64)
        g[f"{name}/teacher"] = np.stack([torch.stack([x for x, _ in tr]).numpy() for tr in tt])     # [n, L, 1, C, H, W]
        g[f"{name}/student"] = np.stack([torch.stack([x for x, _ in tr]).numpy() for tr in st])
        g[f"{name}/teacher_t"] = np.array([t for _, t in tt[0]], np.int64)
        g[f"{name}/student_t"] = np.array([t for _, t in st[0]], np.int64)
        for k, v in allm.items():
            if k == "architecture_type":
                continue
            g[f"{name}/m/{k}"] = np.asarray(v, np.float64)
    path = os.path.join(OUT, "batch_metrics.npz")
    np.savez_compressed(path, **g)
    print(path, os.path.getsize(path), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
