"""PCA projection of trajectories.  CPU: the oracle against sklearn's own transform (the
third-party call the reference makes, scripts/analysis/analyze_trajectories.py:79-80,100).  GPU: dtraj_project against
sklearn per trajectory, at BASELINE frame sizes."""
import numpy as np
import pytest
import torch

from oracle import pca as opca


def _walk(N, L, D, seed):
    rng = np.random.RandomState(seed)
    return (rng.randn(N, 1, D) + 0.1 * np.cumsum(rng.randn(N, L, D), axis=1)).astype(np.float32)


@pytest.mark.parametrize("L,D", [(51, 256), (51, 3072), (7, 12)])
def test_oracle_transform_matches_sklearn(L, D):
    from sklearn.decomposition import PCA
    x = _walk(3, L, D, 1)
    pca = PCA(n_components=3).fit(x[0])
    for n in range(3):
        want = pca.transform(x[n])
        got = opca.transform(x[n], pca.mean_, pca.components_)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


@pytest.mark.gpu
@pytest.mark.parametrize("N,L,C,H", [(37, 51, 1, 16), (9, 51, 3, 32), (300, 50, 1, 16)])
def test_project_trajectories_matches_sklearn(N, L, C, H):
    from distillation_trajectories_b200.analysis import trajectory_pca as tp
    D = C * H * H
    x = _walk(N, L, D, 2).reshape(N, L, C, H, H)
    ref_traj = [torch.from_numpy(x[0, i:i + 1]) for i in range(L)]          # a reference trajectory as the samplers return it
    pca = tp.fit_reference_pca(ref_traj)
    got = tp.project_trajectories(torch.from_numpy(x).cuda(), pca).cpu().numpy()
    assert got.shape == (N, L, 3)
    for n in (0, 1, N // 2, N - 1):
        want = pca.transform(x[n].reshape(L, D))
        np.testing.assert_allclose(got[n], want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
        np.testing.assert_allclose(got[n], opca.transform(x[n].reshape(L, D), pca.mean_, pca.components_), rtol=1e-4,
                                   atol=1e-4 * np.abs(want).max())
