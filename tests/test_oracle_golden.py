"""CPU: the oracle restatement replayed against the committed reference outputs (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import samplers as osmp
from helpers import golden_models, load_golden, oracle_fn

CASES = ["tiny16", "tiny32"]
TIGHT = dict(rtol=1e-6, atol=1e-6)


@pytest.fixture(scope="module", params=CASES)
def case(request):
    torch.set_num_threads(1)
    g, cfg, teacher, student = golden_models(request.param)
    return g, cfg, oracle_fn(teacher), oracle_fn(student)


def test_forward(case):
    g, cfg, ft, fs = case
    x = torch.from_numpy(g["fwd_x"])
    for tval in (0, cfg.timesteps - 1):
        t = torch.full((3,), tval, dtype=torch.long)
        np.testing.assert_allclose(ft(x, t, None).numpy(), g[f"fwd_t{tval}_none"], **TIGHT)
        np.testing.assert_allclose(ft(x, t, torch.ones(3, 1)).numpy(), g[f"fwd_t{tval}_cond1"], **TIGHT)
        np.testing.assert_allclose(ft(x, t, torch.zeros(3, 1)).numpy(), g[f"fwd_t{tval}_cond0"], **TIGHT)
        np.testing.assert_allclose(fs(x, t, torch.ones(3, 1)).numpy(), g[f"fwd_student_t{tval}_cond1"], **TIGHT)


def test_s1(case):
    g, cfg, ft, _ = case
    C, H, T = cfg.channels, cfg.image_size, cfg.timesteps
    for w in (3.0, 1.0):
        torch.manual_seed(5)
        _, traj = osmp.s1_p_sample_loop(ft, (2, C, H, H), T, osmp.diffusion_params(T), T, w)
        np.testing.assert_allclose(torch.stack(traj).numpy(), g[f"s1_w{w}"], rtol=1e-5, atol=1e-5)
    torch.manual_seed(6)
    _, traj = osmp.s1_p_sample_loop(ft, (2, C, H, H), 3 * T, osmp.diffusion_params(3 * T), T, 2.0)
    assert len(traj) == g["s1_strided"].shape[0]
    np.testing.assert_allclose(torch.stack(traj).numpy(), g["s1_strided"], rtol=1e-5, atol=1e-5)


def test_s2(case):
    g, cfg, ft, fs = case
    noise = torch.from_numpy(g["s2_noise"])
    for who, f in (("teacher", ft), ("student", fs)):
        for w in (None, 1.0, 3.0, 7.5):
            tr = osmp.s2_generate_trajectory(f, noise, cfg.timesteps, seed=42, guidance_scale=w)
            ref = g[f"s2_{who}_w{w}"]
            assert len(tr) == cfg.timesteps + 1 == ref.shape[0]
            np.testing.assert_allclose(torch.stack(tr).numpy(), ref, rtol=1e-5, atol=1e-5)
            assert torch.equal(tr[-1], tr[-2])        # no update at t == 0


def test_s3_and_metrics(case):
    g, cfg, ft, fs = case
    C, H, T = cfg.channels, cfg.image_size, cfg.timesteps
    for tag, ss in (("eq", T), ("uneq", max(2, T // 2))):
        tt, st = osmp.s3_generate_pair(ft, fs, (1, C, H, H), T, T, ss, seed=3)
        assert [t for _, t in tt] == list(g[f"s3_{tag}_teacher_t"])
        assert [t for _, t in st] == list(g[f"s3_{tag}_student_t"])
        np.testing.assert_allclose(torch.stack([x for x, _ in tt]).numpy(), g[f"s3_{tag}_teacher"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(torch.stack([x for x, _ in st]).numpy(), g[f"s3_{tag}_student"], rtol=1e-5, atol=1e-5)
        # Q1 on the exact reference trajectories (tuples), same numpy seed as the fixture
        tt_ref = [(torch.from_numpy(a), int(t)) for a, t in zip(g[f"s3_{tag}_teacher"], g[f"s3_{tag}_teacher_t"])]
        st_ref = [(torch.from_numpy(a), int(t)) for a, t in zip(g[f"s3_{tag}_student"], g[f"s3_{tag}_student_t"])]
        np.random.seed(3)
        m = om.trajectory_metrics(tt_ref, st_ref)
        _check_metrics(m, g, f"q1_s3_{tag}/")


def _check_metrics(m, g, prefix):
    keys = [k[len(prefix):] for k in g if k.startswith(prefix)]
    assert sorted(keys) == sorted(m.keys())
    assert len(keys) == 25
    for k in keys:
        np.testing.assert_allclose(np.asarray(m[k], np.float64), g[prefix + k], rtol=1e-9, atol=1e-12,
                                   equal_nan=True, err_msg=k)


def test_q1_q2(case):
    g, cfg, _, _ = case
    T_ = [torch.from_numpy(a) for a in g["s2_teacher_w3.0"]]
    S_ = [torch.from_numpy(a) for a in g["s2_student_w3.0"]]
    np.random.seed(43)
    m = om.trajectory_metrics(T_, S_)
    _check_metrics(m, g, "q1/")
    assert isinstance(m["path_alignment"], np.float32)      # dropped by compare_trajectories' averaging
    T2 = [torch.from_numpy(a) for a in g["s2_teacher_w7.5"]]
    S2 = [torch.from_numpy(a) for a in g["s2_student_w7.5"]]
    td = om.time_dependent_distances([T_, T2], [S_, S2], size_factor=0.5)
    for k in ("teacher_distances", "student_distances", "teacher_avg_per_timestep", "student_avg_per_timestep",
              "teacher_avg_distance", "student_avg_distance", "teacher_std_distance", "student_std_distance"):
        np.testing.assert_allclose(np.asarray(td[k], np.float64), g["q2/" + k], rtol=1e-12, err_msg=k)


def test_compare_trajectories(case):
    """analysis/trajectory_engine.py:117-180 restated with the oracle pieces."""
    g, cfg, ft, fs = case
    C, H, T = cfg.channels, cfg.image_size, cfg.timesteps
    per = {1.0: [], 3.0: []}
    for s in range(2):
        seed = 42 + s
        torch.manual_seed(seed)
        np.random.seed(seed)
        noise = torch.randn(1, C, H, H)
        for gs in (1.0, 3.0):
            a = osmp.s2_generate_trajectory(ft, noise, T, seed=seed, guidance_scale=gs)
            b = osmp.s2_generate_trajectory(fs, noise, T, seed=seed, guidance_scale=gs)
            per[gs].append(om.trajectory_metrics(a, b))
    for gs in (1.0, 3.0):
        avg = om.average_scalar_metrics(per[gs])
        keys = [k.split("/")[2] for k in g if k.startswith(f"cmp/{gs}/")]
        assert sorted(keys) == sorted(avg) and len(keys) == 18
        for k in keys:
            np.testing.assert_allclose(avg[k], g[f"cmp/{gs}/{k}"][0], rtol=1e-5, atol=1e-7, equal_nan=True, err_msg=k)


def test_transform_metrics():
    rows = load_golden("transform")["rows"]
    for r in rows:
        out = om.transform_metrics(*r[:4])
        got = [out["path_length_similarity"], out["trajectory_mse"], out["mean_directional_consistency"],
               out["distribution_similarity"]]
        np.testing.assert_allclose(got, r[4:], rtol=1e-15)


def test_wasserstein_matches_scipy():
    from scipy.stats import wasserstein_distance
    rng = np.random.RandomState(0)
    for n, mth in ((1000, 1000), (256, 256), (7, 7), (50, 31)):
        u, v = rng.randn(n).astype(np.float32), rng.randn(mth).astype(np.float32) * 1.5 + 0.2
        assert abs(om.wasserstein_1d(u, v) - wasserstein_distance(u, v)) < 1e-12
