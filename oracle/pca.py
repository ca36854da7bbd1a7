"""Oracle: PCA projection of trajectories on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows /root/reference/scripts/analysis/analyze_trajectories.py:66-80,100
(process_trajectory, PCA(n_components=3).fit(reference_features), pca.transform(features)); the arithmetic lives in the
third-party dependency scikit-learn (pinned 1.6.1 in the reference's requirements.txt, 1.9.0 in this image):
transform(X) = X @ components_.T - mean_ @ components_.T.
"""
import numpy as np


def process_trajectory(traj):
    return np.stack([np.asarray(f[0] if isinstance(f, tuple) else f).reshape(-1) for f in traj])


def transform(features, mean, components):
    f = np.asarray(features, np.float64)
    c = np.asarray(components, np.float64)
    return f @ c.T - np.asarray(mean, np.float64) @ c.T
