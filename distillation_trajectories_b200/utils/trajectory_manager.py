"""Teacher/student trajectory pairs with the placeholder sampler S3, pickle cache and batched metric
aggregation (drop-in for /root/reference/utils/trajectory_manager.py:9-581).

Each model's loop (store frame, forward, x <- (x - 0.1 eps)/sqrt(0.9) + 0.1 (t/teacher_steps) z) runs
as one captured CUDA loop; noise is drawn from torch's generators with the reference's calls in the
reference's order.  The on-disk format stays ``pickle((teacher_list, student_list))`` of
``(tensor, t)`` tuples per (size factor, sample).
"""
import os
import pickle

import numpy as np
import torch

from .. import sampling
from ..engine import UNetEngine, get_precision
from ..analysis.metrics.trajectory_metrics import compute_trajectory_metrics


class TrajectoryManager:
    def __init__(self, teacher_model, student_model, config, size_factor=1.0, fixed_samples=None):
        self.teacher_model = teacher_model
        self.student_model = student_model
        self.config = config
        self.size_factor = size_factor
        self.fixed_samples = fixed_samples
        os.makedirs(config.trajectory_dir, exist_ok=True)
        self.device = next(teacher_model.parameters()).device

    # ------------------------------------------------------------------ generation
    def _run_model(self, model, x, steps):
        """frames-before-update loop of one model (trajectory_manager.py:98-111) -> [(tensor, t), ...]"""
        cfg = self.config
        idx = sampling.s3_timestep_indices(cfg.sample_steps, steps)
        ts = list(reversed(idx))
        eng = UNetEngine.for_model(model, x.shape[2], max(ts) + 1, get_precision("S3"), self.device)
        n_upd = sum(1 for t in ts if t > 0)
        ndev = sampling.noise_device(self.device)
        noise = torch.stack([torch.randn(x.shape, device=ndev) for _ in range(n_upd)]) if n_upd else None
        traj = sampling.s3_sample(eng, x, ts, cfg.teacher_steps, noise)
        return [(traj[:, k].clone(), t) for k, t in enumerate(ts)]

    def _pair(self, seed, sample):
        cfg = self.config
        self.teacher_model.eval()
        self.student_model.eval()
        shape = (1, cfg.channels, cfg.image_size, cfg.image_size)
        out = []
        for k, (model, steps) in enumerate(((self.teacher_model, cfg.teacher_steps),
                                            (self.student_model, cfg.student_steps))):
            if seed is not None and (sample is None or k == 0):
                torch.manual_seed(seed)
                np.random.seed(seed)
            x = (torch.randn(shape) if sample is None else sample.clone()).to(self.device)
            out.append(self._run_model(model, x, steps))
        return out[0], out[1]

    def generate_trajectory(self, seed=None):
        """One (teacher, student) pair from ``seed`` (trajectory_manager.py:65-165); both start from the
        same x_T and see the same noise sequence because the student re-seeds."""
        return self._pair(seed, None)

    def generate_trajectory_from_sample(self, sample, seed=None):
        """Pair starting from a fixed noise sample (trajectory_manager.py:265-387); errors are reported
        and yield empty lists, as in the reference."""
        try:
            return self._pair(seed, sample)
        except Exception as e:            # reference semantics: print and return empties
            print(f"Error generating trajectory from sample: {e}")
            return [], []

    def _path(self, i, size_factor=None):
        sf = self.size_factor if size_factor is None else size_factor
        return os.path.join(self.config.trajectory_dir, f"trajectory_size_{sf}_sample_{i}.pkl")

    def generate_and_save_trajectories(self, num_samples=10):
        """trajectory_manager.py:207-263: one pickle per pair; a failing sample is reported and skipped."""
        paths = []
        fixed = self.fixed_samples is not None and num_samples <= len(self.fixed_samples)
        for i in range(num_samples):
            try:
                if fixed:
                    pair = self.generate_trajectory_from_sample(self.fixed_samples[i], i)
                else:
                    pair = self.generate_trajectory(seed=i)
            except Exception as e:
                print(f"Error generating trajectory {i}: {e}")
                continue
            with open(self._path(i), "wb") as f:
                pickle.dump(pair, f)
            paths.append(self._path(i))
        return paths

    # ------------------------------------------------------------------ cache
    def _files(self, size_factor):
        prefix = f"trajectory_size_{size_factor}_sample_"
        names = [f for f in os.listdir(self.config.trajectory_dir) if f.startswith(prefix) and f.endswith(".pkl")]
        names.sort(key=lambda s: int(s.split("_sample_")[1].split(".")[0]))
        return names

    def load_trajectories(self, size_factor=None, indices=None):
        """trajectory_manager.py:389-432."""
        sf = self.size_factor if size_factor is None else size_factor
        names = self._files(sf)
        if indices is not None:
            names = [n for n in names if int(n.split("_sample_")[1].split(".")[0]) in indices]
        teachers, students = [], []
        for n in names:
            with open(os.path.join(self.config.trajectory_dir, n), "rb") as f:
                t, s = pickle.load(f)
            teachers.append(t)
            students.append(s)
        return teachers, students

    def compute_trajectory_metrics_batch(self, size_factor=None, batch_size=10):
        """trajectory_manager.py:434-548: per-pair metrics appended to 13 named lists, then ``*_avg`` means."""
        sf = self.size_factor if size_factor is None else size_factor
        names = self._files(sf)
        lists = {"wasserstein_distances": "mean_wasserstein", "wasserstein_distances_per_timestep": "wasserstein_distances",
                 "endpoint_distances": "endpoint_distance", "teacher_path_lengths": "teacher_path_length",
                 "student_path_lengths": "student_path_length", "teacher_efficiency": "teacher_efficiency",
                 "student_efficiency": "student_efficiency"}
        new = ["path_length_similarity", "efficiency_similarity", "mean_velocity_similarity",
               "mean_directional_consistency", "mean_position_difference", "distribution_similarity"]
        allm = {k: [] for k in list(lists) + new + ["architecture_type"]}
        for i in range(0, len(names), batch_size):
            for n in names[i:i + batch_size]:
                with open(os.path.join(self.config.trajectory_dir, n), "rb") as f:
                    t, s = pickle.load(f)
                m = compute_trajectory_metrics(t, s, self.config)
                for dst, src in lists.items():
                    allm[dst].append(m[src])
                for k in new:
                    if k in m:
                        allm[k].append(m[k])
                if hasattr(self, "architecture_type"):
                    allm["architecture_type"].append(self.architecture_type)
        for k in ["endpoint_distances", "teacher_path_lengths", "student_path_lengths", "teacher_efficiency",
                  "student_efficiency", "wasserstein_distances"] + new:
            if allm.get(k):
                allm[k + "_avg"] = sum(allm[k]) / len(allm[k])
        return allm


def generate_trajectories_with_disk_storage(teacher_model, student_model, config, size_factor=1.0, num_samples=10,
                                            fixed_samples=None):
    """trajectory_manager.py:550-581: reuse cached pickles, generate the missing count."""
    manager = TrajectoryManager(teacher_model, student_model, config, size_factor, fixed_samples)
    existing = manager._files(size_factor)
    if len(existing) < num_samples:
        print(f"Generating {num_samples - len(existing)} new trajectories...")
        manager.generate_and_save_trajectories(num_samples - len(existing))
    else:
        print(f"Using {num_samples} existing trajectories...")
    return manager
