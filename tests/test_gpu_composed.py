"""GPU: the COMPOSED paths at the precision and the batch shape the benchmark runs -- the multi-student sweep
(SURVEY 8f rank 2), compare_trajectories at its default (f16) arithmetic (E1), a sampler chunk at the bench's row count,
compute_trajectory_metrics_batch values against the reference's own output (E2), and the model aliases (M4).

Tolerances (north_star): trajectories |got - ref| <= 1e-3 |ref| + 1e-4 max|ref| elementwise; metric scalars rtol 1e-4
GIVEN IDENTICAL TRAJECTORIES: the scalars a sweep returns are compared with the oracle's metrics evaluated on the very
frames the CUDA samplers produced (the samplers are deterministic, so re-generating them yields the sweep's frames bit
for bit), while the frames themselves are compared with the oracle's frames under the trajectory tolerance.
"""
import contextlib
import ctypes as Ct
import io
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import samplers as osmp
from distillation_trajectories_b200 import _lib, grid, sampling, get_precision, umma_error_flag
from distillation_trajectories_b200.analysis import trajectory_engine as te
from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm
from distillation_trajectories_b200.engine import UNetEngine
from distillation_trajectories_b200.models import DiffusionUNet, SimpleUNet, StudentUNet
from distillation_trajectories_b200.utils.trajectory_manager import TrajectoryManager
from helpers import Cfg, assert_close, load_golden, make_model, oracle_fn

pytestmark = pytest.mark.gpu
RTOL, ATOL, RT_METRIC = 1e-3, 1e-4, 1e-4


@pytest.fixture(autouse=True)
def cpu_noise():
    sampling.set_noise_device("cpu")
    yield
    sampling.set_noise_device(None)


def _frames(a):
    """[L, C, H, W] array -> list of [1, C, H, W] CPU tensors (the structure generate_trajectory returns)"""
    return [torch.from_numpy(np.ascontiguousarray(f))[None] for f in a]


def _oracle_pair_metrics(t_frames, s_frames, seed):
    """compute_trajectory_metrics as compare_trajectories reaches it: the global numpy RNG was last seeded with
    seed + 1 inside the student's generate_trajectory (analysis/trajectory_engine.py:91-93 at t = 1)."""
    np.random.seed(seed + 1)
    return om.trajectory_metrics(_frames(t_frames), _frames(s_frames))


def _check_scalars(got, want, what):
    for k in tm.SCALAR_KEYS:
        np.testing.assert_allclose(got[k], want[k], rtol=RT_METRIC, atol=1e-9, equal_nan=True, err_msg=f"{what}: {k}")


# ------------------------------------------------------------------ f2 + E1: multi-student sweep, default precision
@pytest.mark.parametrize("C,H,T", [(1, 16, 8), (3, 32, 5)])
def test_sweep_three_students_default_precision_vs_oracle(C, H, T):
    """grid.sweep (the batched compare_trajectories, teacher generated once for all students) in the DEFAULT arithmetic
    against the oracle restatement of analysis/trajectory_engine.py:117-180 run per student."""
    assert get_precision("S2") == "f16"
    cfg = Cfg(C, H, T)
    teacher = make_model(cfg, 1.0 if H == 16 else 0.5, 31, device="cuda")
    students = {sf: make_model(cfg, sf, 1000 + int(sf * 100), device="cuda") for sf in (0.05, 0.3, 0.5)}
    scales, n_seeds = [1.0, 3.0, 7.5], 3
    chunk_seeds = 2                                      # max_pairs below: the sweep runs seeds {0, 1} and {2} as two chunks
    res = grid.sweep(teacher, students, cfg, scales, n_seeds, reduce=False, max_pairs=chunk_seeds * len(scales))
    assert umma_error_flag() == 0
    ft = oracle_fn(teacher)
    x, seeds, ws = [], [], []
    for s in range(n_seeds):
        torch.manual_seed(42 + s)
        noise = torch.randn(1, C, H, H)
        for gs in scales:
            x.append(noise); seeds.append(42 + s); ws.append(gs)
    x = torch.cat(x)

    def regenerate(model):
        """the sweep's own frames: same chunking, hence the same launch shapes (kernel forms -- column split, split-K -- are
        chosen per launch shape and sum in different orders), and the samplers are deterministic"""
        out = []
        for c0 in range(0, n_seeds, chunk_seeds):
            sl = slice(c0 * len(scales), min(n_seeds, c0 + chunk_seeds) * len(scales))
            groups = [i // len(scales) for i in range(sl.stop - sl.start)]
            out.append(te.generate_trajectories_batched(model, x[sl], seeds[sl], ws[sl], T, "cuda", groups=groups).cpu().numpy().copy())
        return np.concatenate(out)

    t_gpu = regenerate(teacher)
    t_ref = [torch.stack(osmp.s2_generate_trajectory(ft, x[p:p + 1], T, seed=seeds[p], guidance_scale=ws[p]))[:, 0].numpy()
             for p in range(len(seeds))]
    assert_close(t_gpu, np.stack(t_ref), RTOL, ATOL, "teacher frames of the sweep")
    for sf, model in students.items():
        fs = oracle_fn(model)
        s_gpu = regenerate(model)
        s_ref = [torch.stack(osmp.s2_generate_trajectory(fs, x[p:p + 1], T, seed=seeds[p], guidance_scale=ws[p]))[:, 0].numpy()
                 for p in range(len(seeds))]
        assert_close(s_gpu, np.stack(s_ref), RTOL, ATOL, f"student {sf} frames of the sweep")
        for g, gs in enumerate(scales):
            per = [_oracle_pair_metrics(t_gpu[s * len(scales) + g], s_gpu[s * len(scales) + g], 42 + s) for s in range(n_seeds)]
            _check_scalars(res[sf][gs], om.average_scalar_metrics(per), f"sf={sf} w={gs}")


@pytest.mark.parametrize("name", ["tiny16", "tiny32"])
def test_compare_trajectories_default_precision(name):
    """E1 through the reference signature in the default (f16) arithmetic: same keys and structure as the committed
    reference output; values at 1e-4 against the oracle's metrics of the CUDA frames; and -- the end-to-end view --
    within the spread the trajectory tolerance allows of the reference's own fp32 numbers."""
    from helpers import golden_models
    g, cfg, teacher, student = golden_models(name, device="cuda")
    C, H, T = cfg.channels, cfg.image_size, cfg.timesteps
    scales = [1.0, 3.0]
    res = te.compare_trajectories(teacher, student, cfg, guidance_scales=scales, size_factor=0.5, num_samples=2)
    assert set(res) == {"teacher_metrics", "student_metrics"} and res["teacher_metrics"] == res["student_metrics"]
    x, seeds, ws = [], [], []
    for s in range(2):
        torch.manual_seed(42 + s)
        noise = torch.randn(1, C, H, H)
        for gs in scales:
            x.append(noise); seeds.append(42 + s); ws.append(gs)
    x = torch.cat(x)
    groups = [0, 0, 1, 1]
    tg = te.generate_trajectories_batched(teacher, x, seeds, ws, T, "cuda", groups=groups).cpu().numpy().copy()
    sg = te.generate_trajectories_batched(student, x, seeds, ws, T, "cuda", groups=groups).cpu().numpy().copy()
    for gi, gs in enumerate(scales):
        d = res["student_metrics"][gs]
        keys = [k.split("/")[2] for k in g if k.startswith(f"cmp/{gs}/")]
        assert sorted(keys) == sorted(d) and len(keys) == 18
        per = [_oracle_pair_metrics(tg[2 * s + gi], sg[2 * s + gi], 42 + s) for s in range(2)]
        _check_scalars(d, om.average_scalar_metrics(per), f"{name} w={gs}")
        for k in keys:          # vs the unmodified reference (fp32 CPU): frames differ at the 1e-3 level, so do the scalars
            np.testing.assert_allclose(d[k], g[f"cmp/{gs}/{k}"][0], rtol=5e-3, atol=1e-5, equal_nan=True, err_msg=f"{gs}/{k}")


# ------------------------------------------------------------------ the bench's chunk shape
def test_bench_shape_chunk_vs_oracle():
    """One chunk exactly as bench.py runs it (592 seeds x 8 guidance scales = 4736 samples, 8880 forward rows per model,
    CTA-pair / halo / persistent multi-tile forms, two overlapped streams, default f16): sampled rows of both models
    against the oracle, and the chunk's metric reductions against the oracle's metrics of those rows."""
    from bench import Cfg as BCfg, GUIDANCE
    S, G, T = 592, len(GUIDANCE), BCfg.timesteps
    models = []
    for sf, seed in ((1.0, 0), (0.5, 1050)):
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            models.append(DiffusionUNet(BCfg, sf).eval().cuda())
    teacher, student = models
    ck = grid.stage_chunk(list(range(S)), BCfg, GUIDANCE, "cuda")
    keep = []
    red, w1, n_traj = grid.run_chunk(teacher, [student], ck, "cuda", out_traj=keep)
    torch.cuda.synchronize()
    assert n_traj == 2 * S * G and umma_error_flag() == 0
    picks = [0, 295, 591]
    rows = [s * G + g for s in picks for g in range(G)]
    t_gpu, s_gpu = (k[rows].reshape(len(rows), T + 1, 1, 16, 16).cpu().numpy() for k in keep[0])
    red_h, w1_h = red[0][rows].cpu().numpy(), w1[0][rows].cpu().numpy()
    sm = tm.scalar_metrics_batched(red_h, w1_h, 256, 256)
    ft, fs = oracle_fn(teacher), oracle_fn(student)
    worst = 0.0
    for i, (s, g) in enumerate((s, g) for s in picks for g in range(G)):
        torch.manual_seed(42 + s)
        noise = torch.randn(1, 1, 16, 16)
        for f, got, who in ((ft, t_gpu[i], "teacher"), (fs, s_gpu[i], "student")):
            want = torch.stack(osmp.s2_generate_trajectory(f, noise, T, seed=42 + s, guidance_scale=GUIDANCE[g]))[:, 0].numpy()
            assert_close(got, want, RTOL, ATOL, f"{who} seed {s} w={GUIDANCE[g]} inside the 8880-row chunk")
            worst = max(worst, float((np.abs(got - want) / (RTOL * np.abs(want) + ATOL * np.abs(want).max())).max()))
        want_m = _oracle_pair_metrics(t_gpu[i], s_gpu[i], 42 + s)
        _check_scalars({k: sm[k][i] for k in tm.SCALAR_KEYS}, want_m, f"seed {s} w={GUIDANCE[g]}")
    assert worst < 1.0
    for m in models:
        for eng in list(m.__dict__.get("_dtraj_engines", {}).values()):
            eng[1].close()


@pytest.mark.parametrize("sf", [1.0, 0.5])
def test_forward_guard_bands_at_bench_row_count(sf):
    """8880 forward rows (the bench's rows per model and step, f16): nothing is written outside the workspace the ABI
    asked for, nor outside the eps tensor."""
    cfg = Cfg(1, 16, 50)
    model = make_model(cfg, sf, 17, stress=False, device="cuda")
    eng = UNetEngine.for_model(model, 16, 50, "f16")
    R = 8880
    need = eng.workspace_bytes(R)
    guard = 1 << 20
    ws = torch.full((guard + need + guard,), 0xAB, dtype=torch.uint8, device="cuda")
    base = ws.data_ptr() + guard
    assert base % 256 == 0
    x = torch.randn(R, 1, 16, 16, device="cuda")
    out = torch.full((R + 2, 1, 16, 16), 7.0, device="cuda")
    variants = (torch.arange(R, device="cuda") % 3).to(torch.int32)
    _lib.check(_lib.load().dtraj_unet_forward(eng.handle, _lib.ptr(x), R, 31, _lib.ptr(variants), Ct.c_void_p(out[1:].data_ptr()),
                                              Ct.c_void_p(base), need, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert bool((ws[:guard] == 0xAB).all()) and bool((ws[guard + need:] == 0xAB).all()), "wrote outside the workspace"
    assert bool((out[0] == 7.0).all()) and bool((out[R + 1] == 7.0).all()), "wrote outside the eps tensor"
    assert torch.isfinite(out[1:R + 1]).all() and umma_error_flag() == 0
    # ... and rows of the big launch equal the same rows run as a small batch (within the f16 forward bound)
    sel = [0, 4439, 8879]
    small = eng.forward(x[sel], 31, variants[sel])
    ref = out[1:][sel]
    assert float((small - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
    eng.close()


# ------------------------------------------------------------------ E2: batch metrics against the reference's own output
@pytest.mark.parametrize("name", ["b16", "b32"])
def test_compute_trajectory_metrics_batch_values_vs_reference(name, tmp_path):
    """utils/trajectory_manager.py:434-548 on the very pickles the unmodified reference wrote (rebuilt from
    tests/golden/batch_metrics.npz, oracle/make_golden_batch.py): every list and every ``*_avg`` at 1e-4."""
    g = load_golden("batch_metrics")
    C, H, T, ss, n = (int(v) for v in g[f"{name}/meta"])
    sf_t, sf_s = (float(v) for v in g[f"{name}/sf"])
    cfg = Cfg(C, H, T, trajectory_dir=str(tmp_path), student_steps=ss)
    for i in range(n):
        pair = tuple([(torch.from_numpy(g[f"{name}/{who}"][i, k]), int(t)) for k, t in enumerate(g[f"{name}/{who}_t"])]
                     for who in ("teacher", "student"))
        with open(os.path.join(str(tmp_path), f"trajectory_size_{sf_s}_sample_{i}.pkl"), "wb") as f:
            pickle.dump(pair, f)
    teacher = make_model(cfg, sf_t, 100, device="cuda")
    student = make_model(cfg, sf_s, 200, device="cuda")
    mgr = TrajectoryManager(teacher, student, cfg, size_factor=sf_s)
    np.random.seed(7)                                    # oracle/make_golden_batch.py: NP_SEED
    allm = mgr.compute_trajectory_metrics_batch(batch_size=2)
    keys = [k[len(name) + 3:] for k in g if k.startswith(f"{name}/m/")]
    assert sorted(keys) == sorted(k for k in allm if k != "architecture_type") and allm["architecture_type"] == []
    for k in keys:
        np.testing.assert_allclose(np.asarray(allm[k], np.float64), g[f"{name}/m/{k}"], rtol=RT_METRIC, atol=1e-7,
                                   equal_nan=True, err_msg=k)


# ------------------------------------------------------------------ M4: aliases
def test_simple_and_student_unet_aliases():
    """models.py:227-242: SimpleUNet(config) == DiffusionUNet(config, 1.0); StudentUNet(config, sf, architecture_type)
    == DiffusionUNet(config, sf) whatever the architecture type -- same parameters from the same seed, same outputs."""
    cfg = Cfg(1, 16, 8)
    x = torch.randn(3, 1, 16, 16, device="cuda")
    t = torch.full((3,), 5, dtype=torch.long, device="cuda")
    cond = torch.tensor([[0.0], [1.0], [1.0]], device="cuda")
    for make_alias, sf in ((lambda: SimpleUNet(cfg), 1.0), (lambda: StudentUNet(cfg, 0.3, architecture_type="tiny"), 0.3),
                           (lambda: StudentUNet(cfg, size_factor=0.05), 0.05)):
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(9)
            a = make_alias().eval().cuda()
            torch.manual_seed(9)
            b = DiffusionUNet(cfg, sf).eval().cuda()
        assert isinstance(a, DiffusionUNet) and a.size_factor == sf and a.dims == b.dims
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
        for c in (None, cond):
            assert torch.equal(a(x, t, c), b(x, t, c))
    assert umma_error_flag() == 0


# ------------------------------------------------------------------ stale weights, silent overflow, whose error
def test_engine_is_rebuilt_after_in_place_data_write():
    """``p.data.add_()`` bumps no tensor version; the cached packed engine must still be refreshed (content checksum)."""
    cfg = Cfg(1, 16, 4)
    m = make_model(cfg, 0.2, 41, device="cuda")
    x = torch.randn(2, 1, 16, 16, device="cuda")
    t = torch.full((2,), 3, dtype=torch.long, device="cuda")
    a = m(x, t).clone()
    eng0 = next(iter(m.__dict__["_dtraj_engines"].values()))[1]
    assert torch.equal(m(x, t), a) and next(iter(m.__dict__["_dtraj_engines"].values()))[1] is eng0    # cache hit
    m.final.bias.data.add_(0.25)                                     # the EMA / manual-load idiom
    b = m(x, t)
    assert next(iter(m.__dict__["_dtraj_engines"].values()))[1] is not eng0
    assert torch.allclose(b, a + 0.25, atol=1e-5)
    UNetEngine.invalidate(m)
    assert "_dtraj_engines" not in m.__dict__


def test_s3_long_run_overflow_is_never_silent(tmp_path):
    """S3 divides by sqrt(0.9) at every step (utils/trajectory_manager.py:194-203), so x grows ~190x over 100 steps: in the
    default f16 arithmetic the frames either stay finite or the call raises DtrajError -- never inf/NaN handed out."""
    from distillation_trajectories_b200 import DtrajError
    cfg = Cfg(1, 16, 100, trajectory_dir=str(tmp_path))
    teacher = make_model(cfg, 0.3, 51, device="cuda")
    student = make_model(cfg, 0.1, 52, device="cuda")
    mgr = TrajectoryManager(teacher, student, cfg, size_factor=0.1)
    try:
        tt, st = mgr.generate_trajectory(seed=1)
    except DtrajError as e:
        assert "fp16 range" in str(e)
    else:
        assert len(tt) == 100 and all(torch.isfinite(x).all() for x, _ in tt + st)
    assert umma_error_flag() == 0


def test_error_word_belongs_to_the_engine_that_overflowed():
    """two models on one device: the fp16 overflow of one is reported for THAT model, the other keeps working"""
    from distillation_trajectories_b200 import DtrajError
    cfg = Cfg(1, 16, 4)
    bad, good = make_model(cfg, 0.2, 61, device="cuda"), make_model(cfg, 0.1, 62, device="cuda")
    with torch.no_grad():
        bad.enc2.conv1.weight.mul_(3e4)
    e_bad, e_good = UNetEngine.for_model(bad, 16, 4, "f16"), UNetEngine.for_model(good, 16, 4, "f16")
    x = torch.randn(2, 1, 16, 16, device="cuda")
    e_bad.forward(x, 2)
    out = e_good.forward(x, 2)
    torch.cuda.synchronize()
    e_good.check_errors()                                            # its own word is clean
    assert torch.isfinite(out).all()
    with pytest.raises(DtrajError, match=r"dims=\[25, 50, 50, 50\].*fp16 range"):
        e_bad.check_errors()
    e_bad.check_errors()                                             # read-and-clear
    assert umma_error_flag() == 0
