"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python -m oracle.make_golden

The reference's own tests hold no golden vectors for this path, so these files
are the pin: inputs are seeded, outputs are whatever the reference computes.  The fixtures are
committed; this script is how they were made.
"""
import contextlib
import io
import os
import shutil

import numpy as np
import torch

from . import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def bn_stress(model, seed):
    """Randomise BatchNorm statistics/affine so that folding bugs cannot hide behind the
    near-identity default init (SURVEY.md 8c golden-vector recipe)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            n = m.num_features
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(n, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(n, generator=g) * 0.1)


def make_model(ref, cfg, sf, seed, stress=True):
    torch.manual_seed(seed)
    with quiet():
        m = ref.models.DiffusionUNet(cfg, sf).eval()
    if stress:
        bn_stress(m, seed + 1)
    return m


def sd_arrays(model, prefix):
    return {prefix + k: v.numpy() for k, v in model.state_dict().items() if torch.is_floating_point(v)}


def stack(traj):
    ims = [t[0] if isinstance(t, tuple) else t for t in traj]
    return torch.stack(ims).numpy()


def metrics_arrays(m, prefix):
    out = {}
    for k, v in m.items():
        out[prefix + k] = np.asarray(v, dtype=np.float64)
    return out


def case(ref, name, C, H, T, sf_t, sf_s, store_weights):
    cfg = refload.RefConfig(channels=C, image_size=H, timesteps=T,
                            trajectory_dir=f"/tmp/dtraj_golden_{name}")
    teacher = make_model(ref, cfg, sf_t, 100)
    student = make_model(ref, cfg, sf_s, 200)
    g = {"meta": np.array([C, H, T], np.int64), "sf": np.array([sf_t, sf_s], np.float64)}
    if store_weights:
        g.update(sd_arrays(teacher, "teacher/"))
        g.update(sd_arrays(student, "student/"))
    g["teacher_wsum"] = np.array([float(sum(v.double().sum() for v in teacher.state_dict().values()))])
    g["student_wsum"] = np.array([float(sum(v.double().sum() for v in student.state_dict().values()))])

    # ---- single forwards (models.py:159-224)
    torch.manual_seed(1)
    x = torch.randn(3, C, H, H)
    g["fwd_x"] = x.numpy()
    for tval in (0, T - 1):
        t = torch.full((3,), tval, dtype=torch.long)
        with torch.no_grad():
            g[f"fwd_t{tval}_none"] = teacher(x, t, None).numpy()
            g[f"fwd_t{tval}_cond1"] = teacher(x, t, torch.ones(3, 1)).numpy()
            g[f"fwd_t{tval}_cond0"] = teacher(x, t, torch.zeros(3, 1)).numpy()
            g[f"fwd_student_t{tval}_cond1"] = student(x, t, torch.ones(3, 1)).numpy()

    # ---- S1 p_sample_loop (utils/diffusion.py:160-212), seed 5, w = 3.0 and w = 1.0
    params = ref.diffusion.get_diffusion_params(T, cfg)
    for w in (3.0, 1.0):
        torch.manual_seed(5)
        with contextlib.redirect_stderr(io.StringIO()):
            _, traj = ref.diffusion.p_sample_loop(teacher, (2, C, H, H), T, params, device="cpu", config=cfg,
                                                  track_trajectory=True, guidance_scale=w)
        g[f"s1_w{w}"] = stack(traj)
    # S1 with a strided schedule: sample_steps = 3T, config.timesteps = T
    params3 = ref.diffusion.get_diffusion_params(3 * T, cfg)
    torch.manual_seed(6)
    with contextlib.redirect_stderr(io.StringIO()):
        _, traj = ref.diffusion.p_sample_loop(teacher, (2, C, H, H), 3 * T, params3, device="cpu", config=cfg,
                                              track_trajectory=True, guidance_scale=2.0)
    g["s1_strided"] = stack(traj)

    # ---- S2 generate_trajectory (analysis/trajectory_engine.py:24-115)
    torch.manual_seed(42)
    noise = torch.randn(1, C, H, H)
    g["s2_noise"] = noise.numpy()
    s2 = {}
    for who, model in (("teacher", teacher), ("student", student)):
        for w in (None, 1.0, 3.0, 7.5):
            with quiet(), contextlib.redirect_stderr(io.StringIO()):
                tr = ref.trajectory_engine.generate_trajectory(model, noise, T, "cpu", seed=42, guidance_scale=w)
            s2[(who, w)] = tr
            g[f"s2_{who}_w{w}"] = stack(tr)

    # ---- Q1 compute_trajectory_metrics on the w = 3.0 pair (numpy RNG as generate_trajectory leaves it)
    np.random.seed(43)
    m = ref.trajectory_metrics.compute_trajectory_metrics(s2[("teacher", 3.0)], s2[("student", 3.0)], cfg)
    g.update(metrics_arrays(m, "q1/"))
    # ---- Q2
    with quiet():
        td = ref.time_dependent.analyze_time_dependent_distances(
            [s2[("teacher", 3.0)], s2[("teacher", 7.5)]], [s2[("student", 3.0)], s2[("student", 7.5)]], cfg, size_factor=sf_s)
    for k in ("teacher_distances", "student_distances", "teacher_avg_per_timestep", "student_avg_per_timestep",
              "teacher_avg_distance", "student_avg_distance", "teacher_std_distance", "student_std_distance"):
        g["q2/" + k] = np.asarray(td[k], np.float64)

    # ---- compare_trajectories (analysis/trajectory_engine.py:117-180)
    with quiet(), contextlib.redirect_stderr(io.StringIO()):
        cmp_ = ref.trajectory_engine.compare_trajectories(teacher, student, cfg, guidance_scales=[1.0, 3.0],
                                                          size_factor=sf_s, num_samples=2)
    for gs, d in cmp_["student_metrics"].items():
        for k, v in d.items():
            g[f"cmp/{gs}/{k}"] = np.array([v], np.float64)

    # ---- S3 TrajectoryManager (utils/trajectory_manager.py:65-205), equal and unequal step counts
    shutil.rmtree(cfg.trajectory_dir, ignore_errors=True)
    for tag, ts_, ss_ in (("eq", T, T), ("uneq", T, max(2, T // 2))):
        cfg.sample_steps, cfg.teacher_steps, cfg.student_steps = T, ts_, ss_
        with quiet():
            mgr = ref.trajectory_manager.TrajectoryManager(teacher, student, cfg, size_factor=sf_s)
            tt, st = mgr.generate_trajectory(seed=3)
        g[f"s3_{tag}_teacher"] = stack(tt)
        g[f"s3_{tag}_student"] = stack(st)
        g[f"s3_{tag}_teacher_t"] = np.array([t for _, t in tt], np.int64)
        g[f"s3_{tag}_student_t"] = np.array([t for _, t in st], np.int64)
        np.random.seed(3)
        m3 = ref.trajectory_metrics.compute_trajectory_metrics(tt, st, cfg)
        g.update(metrics_arrays(m3, f"q1_s3_{tag}/"))
    cfg.sample_steps = cfg.teacher_steps = cfg.student_steps = T

    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
    print(name, "->", sum(v.nbytes for v in g.values()) // 1024, "KiB raw")


def main():
    ref = refload.load()
    torch.set_num_threads(1)   # fixtures must not depend on the thread count
    case(ref, "tiny16", 1, 16, 6, 0.1, 0.05, store_weights=True)
    case(ref, "tiny32", 3, 32, 4, 0.2, 0.1, store_weights=False)
    # Q3 transform_metrics (utils/metric_transformations.py:3-38)
    rows = []
    for args in ((0.6, 0.3, -0.4, 0.5), (0.0, -2.0, 1.0, 0.0), (0.69, 5.0, 0.2, 3.0)):
        r = ref.metric_transformations.transform_metrics(*args)
        rows.append(list(args) + [r["path_length_similarity"], r["trajectory_mse"], r["mean_directional_consistency"],
                                  r["distribution_similarity"]])
    np.savez_compressed(os.path.join(OUT, "transform.npz"), rows=np.array(rows, np.float64))


if __name__ == "__main__":
    main()
