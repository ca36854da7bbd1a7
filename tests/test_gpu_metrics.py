"""GPU: trajectory-metric kernels and their drop-in wrappers against the committed reference
outputs and the CPU oracle.  Tolerance (north_star): rtol 1e-4 on metric scalars."""
import numpy as np
import pytest
import torch

from oracle import metrics as om
from distillation_trajectories_b200 import sampling, set_precision
from distillation_trajectories_b200.analysis import trajectory_engine as te
from distillation_trajectories_b200.analysis.metrics import time_dependent as td
from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm
from distillation_trajectories_b200.utils.trajectory_manager import TrajectoryManager, generate_trajectories_with_disk_storage
from helpers import golden_models, load_golden

pytestmark = pytest.mark.gpu
RT = 1e-4


def _frames(a):
    return [torch.from_numpy(x) for x in a]


def _check25(m, g, prefix, rtol=RT):
    keys = [k[len(prefix):] for k in g if k.startswith(prefix)]
    assert list(m.keys()) and sorted(keys) == sorted(m.keys()) and len(keys) == 25
    for k in keys:
        np.testing.assert_allclose(np.asarray(m[k], np.float64), g[prefix + k], rtol=rtol, atol=1e-7, equal_nan=True, err_msg=k)


@pytest.mark.parametrize("name", ["tiny16", "tiny32"])
def test_q1_q2_on_reference_trajectories(name):
    g = load_golden(name)
    T_, S_ = _frames(g["s2_teacher_w3.0"]), _frames(g["s2_student_w3.0"])
    np.random.seed(43)
    m = tm.compute_trajectory_metrics(T_, S_)
    _check25(m, g, "q1/")
    assert isinstance(m["path_alignment"], np.float32) and isinstance(m["endpoint_distance"], float)
    assert isinstance(m["teacher_velocities"], list)
    T2, S2 = _frames(g["s2_teacher_w7.5"]), _frames(g["s2_student_w7.5"])
    r = td.analyze_time_dependent_distances([T_, T2], [S_, S2], None, size_factor=0.5)
    for k in ("teacher_distances", "student_distances", "teacher_avg_per_timestep", "student_avg_per_timestep",
              "teacher_avg_distance", "student_avg_distance", "teacher_std_distance", "student_std_distance"):
        np.testing.assert_allclose(np.asarray(r[k], np.float64), g["q2/" + k], rtol=RT, err_msg=k)
    assert r["size_factor"] == 0.5
    # S3 tuples, equal and unequal lengths (trajectory_metrics.py:29-37,239-279)
    for tag in ("eq", "uneq"):
        tt = [(torch.from_numpy(a), int(t)) for a, t in zip(g[f"s3_{tag}_teacher"], g[f"s3_{tag}_teacher_t"])]
        st = [(torch.from_numpy(a), int(t)) for a, t in zip(g[f"s3_{tag}_student"], g[f"s3_{tag}_student_t"])]
        np.random.seed(3)
        _check25(tm.compute_trajectory_metrics(tt, st), g, f"q1_s3_{tag}/")


def test_empty_inputs():
    r = td.analyze_time_dependent_distances([], [], None)
    assert r["teacher_distances"] == [] and r["teacher_avg_distance"] == 0
    out = tm.pair_reductions(torch.zeros(0, 5, 256, device="cuda"), torch.zeros(0, 5, 256, device="cuda"))
    assert out.shape == (0, 5, 6)


@pytest.mark.parametrize("N,L,D", [(5, 51, 256), (3, 51, 3072), (130, 7, 256), (2, 1, 256), (9, 50, 768), (4, 3, 4096)])
def test_pair_reductions_vs_numpy(N, L, D):
    rng = np.random.RandomState(N * 1000 + L)
    T = np.cumsum(rng.randn(N, L, D).astype(np.float32) * 0.1, axis=1).astype(np.float32)
    S = (T + rng.randn(N, L, D).astype(np.float32) * 0.05).astype(np.float32)
    S[0] = T[0]
    got = tm.pair_reductions(torch.from_numpy(T).cuda(), torch.from_numpy(S).cuda()).cpu().numpy()
    T64, S64 = T.astype(np.float64), S.astype(np.float64)
    want = np.zeros((N, L, 6))
    want[:, :, 0] = ((T64 - S64) ** 2).sum(-1)
    if L > 1:
        dT, dS = np.diff(T64, axis=1), np.diff(S64, axis=1)
        want[:, :-1, 1], want[:, :-1, 2], want[:, :-1, 3] = (dT ** 2).sum(-1), (dS ** 2).sum(-1), (dT * dS).sum(-1)
    want[:, 0, 4], want[:, 0, 5] = ((T64[:, -1] - T64[:, 0]) ** 2).sum(-1), ((S64[:, -1] - S64[:, 0]) ** 2).sum(-1)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-6 * np.abs(want).max())
    assert (got[0, :, 0] == 0).all()


@pytest.mark.parametrize("D,K", [(256, 256), (3072, 1000), (768, 768), (1024, 1000), (3072, 7)])
def test_wasserstein_vs_oracle(D, K):
    rng = np.random.RandomState(D + K)
    N, L = 4, 6
    T = rng.randn(N, L, D).astype(np.float32)
    S = (rng.randn(N, L, D) * 1.3 + 0.2).astype(np.float32)
    if K == D:
        idx, idx_set, sets = None, None, None
    else:
        sets = np.stack([np.stack([rng.choice(D, K, replace=False) for _ in range(L)]) for _ in range(2)]).astype(np.int32)
        idx, idx_set = torch.from_numpy(sets), torch.tensor([0, 1, 1, 0], dtype=torch.int32)
    got = tm.wasserstein_frames(torch.from_numpy(T).cuda(), torch.from_numpy(S).cuda(), idx, idx_set).cpu().numpy()
    for n in range(N):
        for i in range(L):
            sel = np.arange(D) if sets is None else sets[[0, 1, 1, 0][n], i]
            want = om.wasserstein_1d(T[n, i, sel], S[n, i, sel])
            np.testing.assert_allclose(got[n, i], want, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("D,K,L", [(3072, 1000, 51), (1024, 1000, 3), (4096, 4096, 2), (2, 1, 5), (1500, 7, 4)])
def test_device_index_sets_reproduce_numpy_legacy_stream(D, K, L):
    """dtraj_numpy_choice_sets == np.random.RandomState(seed).choice(D, K, replace=False) called L times (the subsample
    draws of trajectory_metrics.py:301-306): bit-exact, including seeds at the ends of the 32-bit range and a seed count
    that is not a multiple of the kernel's block."""
    import ctypes as C
    from distillation_trajectories_b200 import _lib
    seeds = [0, 1, 43, 44, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 123456789] + list(range(1000, 1011))
    sd = torch.tensor(seeds, dtype=torch.int64).to(torch.int32).cuda()
    out = torch.full((len(seeds), L, K), -1, dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().dtraj_numpy_choice_sets(_lib.ptr(sd), len(seeds), L, D, K, _lib.ptr(out), _lib.stream_ptr()))
    got = out.cpu().numpy()
    for i, s in enumerate(seeds):
        rs = np.random.RandomState(s)
        want = np.stack([rs.choice(D, K, replace=False) for _ in range(L)])
        np.testing.assert_array_equal(got[i], want, err_msg=f"seed {s}")
    # the wrapper the sweeps use (seed + 1 as compare_trajectories' callers leave the RNG) agrees with the host loop
    a = te.wasserstein_index_sets_device([42, 43, 44], 50, L, D, "cuda")
    b = te.wasserstein_index_sets([42, 43, 44], 50, L, D)
    if K == min(1000, D) and K < D:
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    else:
        assert (a is None) == (b is None)


@pytest.mark.parametrize("name", ["tiny16", "tiny32"])
def test_compare_trajectories_fixture(name):
    """analysis/trajectory_engine.py:117-180 end to end (fp32 convolutions so that the metric scalars
    can be held to 1e-3 although they sit behind a 6-step sampling loop)."""
    g, cfg, teacher, student = golden_models(name, device="cuda")
    sampling.set_noise_device("cpu")
    set_precision("fp32", "S2")
    try:
        res = te.compare_trajectories(teacher, student, cfg, guidance_scales=[1.0, 3.0], size_factor=0.5, num_samples=2)
    finally:
        sampling.set_noise_device(None)
        set_precision("f16", "S2")
    assert set(res) == {"teacher_metrics", "student_metrics"}
    for gs in (1.0, 3.0):
        d = res["student_metrics"][gs]
        keys = [k.split("/")[2] for k in g if k.startswith(f"cmp/{gs}/")]
        assert sorted(keys) == sorted(d) and len(keys) == 18
        assert res["teacher_metrics"][gs] == d
        for k in keys:
            np.testing.assert_allclose(d[k], g[f"cmp/{gs}/{k}"][0], rtol=1e-3, atol=1e-6, equal_nan=True, err_msg=f"{gs}/{k}")


def test_trajectory_manager_disk_roundtrip(tmp_path):
    g, cfg, teacher, student = golden_models("tiny16", device="cuda")
    cfg.trajectory_dir = str(tmp_path)
    mgr = generate_trajectories_with_disk_storage(teacher, student, cfg, size_factor=0.05, num_samples=3)
    assert isinstance(mgr, TrajectoryManager)
    tt, ss = mgr.load_trajectories()
    assert len(tt) == 3 and len(tt[0]) == cfg.timesteps and isinstance(tt[0][0], tuple)
    again = generate_trajectories_with_disk_storage(teacher, student, cfg, size_factor=0.05, num_samples=3)
    allm = again.compute_trajectory_metrics_batch(batch_size=2)
    assert len(allm["endpoint_distances"]) == 3 and "endpoint_distances_avg" in allm
    assert len(allm["wasserstein_distances_per_timestep"][0]) == cfg.timesteps


# ------------------------------------------------------------------ size-independent properties at BASELINE sizes
def test_properties_full_size():
    """[N, 50, 3, 32, 32] chunk of BASELINE config 5: identities that hold for any input."""
    N, L, D = 512, 50, 3072
    gen = torch.Generator(device="cuda").manual_seed(1234)
    x0 = torch.randn(N, 1, D, device="cuda", generator=gen)
    T = x0 + 0.1 * torch.cumsum(torch.randn(N, L, D, device="cuda", generator=gen), dim=1)
    S = T + 0.05 * torch.randn(N, L, D, device="cuda", generator=gen)
    r = tm.pair_reductions(T, S)
    # (a) a trajectory against itself: zero distance, equal velocities, dot = |v|^2
    rs = tm.pair_reductions(T, T)
    assert (rs[:, :, 0] == 0).all() and torch.equal(rs[:, :, 1], rs[:, :, 2]) and torch.equal(rs[:, :, 1], rs[:, :, 3])
    # (b) symmetry: swapping the arguments swaps the velocity columns and the end-to-end columns
    rsw = tm.pair_reductions(S, T)
    assert torch.equal(rsw[:, :, 0], r[:, :, 0]) and torch.equal(rsw[:, :, 1], r[:, :, 2]) and torch.equal(rsw[:, :, 4], r[:, :, 5])
    # (c) exact scaling by a power of two: every quadratic sum scales by 4
    r2 = tm.pair_reductions(2 * T, 2 * S)
    assert torch.equal(r2, 4 * r)
    # (d) agreement with torch reductions
    want_d = ((T.double() - S.double()) ** 2).sum(-1)
    torch.testing.assert_close(r[:, :, 0].double(), want_d, rtol=2e-5, atol=0)
    want_v = ((T[:, 1:].double() - T[:, :-1].double()) ** 2).sum(-1)
    torch.testing.assert_close(r[:, :-1, 1].double(), want_v, rtol=2e-5, atol=0)
    # (e) Cauchy-Schwarz on the direction dot product
    assert (r[:, :-1, 3].abs() <= (r[:, :-1, 1].sqrt() * r[:, :-1, 2].sqrt()) * (1 + 1e-5)).all()
    # (f) W1 is invariant under a common permutation of elements and W1(x, x + c) = |c|
    w_all = tm.wasserstein_frames(T[:8, :, :1024].contiguous(), (T[:8, :, :1024] + 0.25).contiguous())
    torch.testing.assert_close(w_all, torch.full_like(w_all, 0.25), rtol=1e-4, atol=1e-6)


def test_packed_store_matches_per_sample_pickles(tmp_path):
    """the packed format (one batched loop, one file) holds exactly the frames the per-sample
    path generates, its reader yields the reference structure, and the batched metrics agree with the per-pair ones."""
    from distillation_trajectories_b200.utils import trajectory_store as store
    g, cfg, teacher, student = golden_models("tiny16", device="cuda")
    sampling.set_noise_device("cpu")
    try:
        for ss in (cfg.timesteps, max(2, cfg.timesteps // 2)):          # equal and unequal step counts
            cfg.student_steps = ss
            cfg.trajectory_dir = str(tmp_path / f"pk_{ss}")
            a = TrajectoryManager(teacher, student, cfg, size_factor=0.05)
            a.generate_and_save_trajectories(4)
            cfg.trajectory_dir = str(tmp_path / f"pack_{ss}")
            b = generate_trajectories_with_disk_storage(teacher, student, cfg, size_factor=0.05, num_samples=4, packed=True)
            assert len(store.list_packs(cfg.trajectory_dir, 0.05)) == 1 and store.stored_samples(cfg.trajectory_dir, 0.05) == {0, 1, 2, 3}
            cfg.trajectory_dir = str(tmp_path / f"pk_{ss}")
            ta, sa = a.load_trajectories()
            cfg.trajectory_dir = str(tmp_path / f"pack_{ss}")
            tb, sb = b.load_trajectories()
            assert len(ta) == len(tb) == 4
            for la, lb in zip(ta + sa, tb + sb):
                assert [t for _, t in la] == [t for _, t in lb]
                for (xa, _), (xb, _) in zip(la, lb):
                    assert torch.equal(xa.cpu(), xb.cpu())
            np.random.seed(0)
            mb = b.compute_trajectory_metrics_batch()
            cfg.trajectory_dir = str(tmp_path / f"pk_{ss}")
            np.random.seed(0)
            ma = a.compute_trajectory_metrics_batch()
            for k in ("endpoint_distances", "teacher_path_lengths", "wasserstein_distances", "distribution_similarity",
                      "mean_directional_consistency", "wasserstein_distances_per_timestep"):
                np.testing.assert_allclose(np.asarray(mb[k], np.float64), np.asarray(ma[k], np.float64), rtol=1e-5, atol=1e-8, err_msg=k)
            assert abs(mb["endpoint_distances_avg"] - ma["endpoint_distances_avg"]) < 1e-6
    finally:
        sampling.set_noise_device(None)
        cfg.student_steps = cfg.timesteps
