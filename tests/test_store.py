"""CPU: the packed on-disk trajectory format (utils/trajectory_store.py) -- round trip, listing, and the
reference structure its reader yields (utils/trajectory_manager.py:389-432)."""
import os
import pickle

import numpy as np
import pytest
import torch

from distillation_trajectories_b200.utils import trajectory_store as store


def _fake(n, lt, ls, seed):
    rng = np.random.RandomState(seed)
    return rng.randn(n, lt, 1, 16, 16).astype(np.float32), rng.randn(n, ls, 1, 16, 16).astype(np.float32)


def test_pack_round_trip_and_reference_structure(tmp_path):
    T, S = _fake(5, 6, 4, 0)
    tt, st = [5, 4, 3, 2, 1, 0], [5, 3, 1, 0]
    path = store.write_pack(str(tmp_path), 0.5, [10, 11, 12, 13, 14], T, S, tt, st)
    assert os.path.basename(path) == "trajectory_size_0.5_pack_10_5.npz"
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".tmp.npz")]
    pack = store.read_pack(path)
    np.testing.assert_array_equal(pack["teacher"], T)
    np.testing.assert_array_equal(pack["student"], S)
    teachers, students = store.as_reference_lists(pack, torch.from_numpy)
    assert len(teachers) == len(students) == 5
    assert len(teachers[0]) == 6 and len(students[0]) == 4
    x, t = teachers[2][1]
    assert t == 4 and tuple(x.shape) == (1, 1, 16, 16) and torch.equal(x[0], torch.from_numpy(T[2, 1]))
    only = store.as_reference_lists(pack, torch.from_numpy, indices={11, 14})
    assert len(only[0]) == 2 and torch.equal(only[1][1][3][0][0], torch.from_numpy(S[4, 3]))


def test_listing_mixes_formats_and_size_factors(tmp_path):
    d = str(tmp_path)
    T, S = _fake(3, 2, 2, 1)
    store.write_pack(d, 0.5, [0, 1, 2], T, S, [1, 0], [1, 0])
    store.write_pack(d, 0.5, [7, 8, 9], T, S, [1, 0], [1, 0])
    store.write_pack(d, 0.25, [0, 1, 2], T, S, [1, 0], [1, 0])
    for i in (3, 20):
        with open(os.path.join(d, f"trajectory_size_0.5_sample_{i}.pkl"), "wb") as f:
            pickle.dump(([], []), f)
    assert [os.path.basename(p) for p in store.list_packs(d, 0.5)] == ["trajectory_size_0.5_pack_0_3.npz",
                                                                       "trajectory_size_0.5_pack_7_3.npz"]
    assert [i for i, _ in store.list_pickles(d, 0.5)] == [3, 20]
    assert store.stored_samples(d, 0.5) == {0, 1, 2, 3, 7, 8, 9, 20}
    assert store.stored_samples(d, 0.25) == {0, 1, 2} and store.stored_samples(d, 1.0) == set()


def test_write_pack_rejects_bad_shapes(tmp_path):
    T, S = _fake(2, 3, 3, 2)
    with pytest.raises(ValueError):
        store.write_pack(str(tmp_path), 1.0, [0], T, S, [2, 1, 0], [2, 1, 0])            # 2 arrays rows, 1 sample id
    with pytest.raises(ValueError):
        store.write_pack(str(tmp_path), 1.0, [0, 1], T, S, [2, 1], [2, 1, 0])            # timestep list too short
    with pytest.raises(ValueError):
        store.write_pack(str(tmp_path), 1.0, [0, 1], T[:, :, 0], S, [2, 1, 0], [2, 1, 0])
