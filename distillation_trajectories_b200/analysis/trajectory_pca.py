"""PCA view of trajectories: the numeric core of
scripts/analysis/analyze_trajectories.py:52-120 and :215-300 without the plotting.

The reference fits ``sklearn.decomposition.PCA(n_components=3)`` on ONE reference trajectory's features
(``process_trajectory``: every frame flattened, [L, D]) and then calls ``pca.transform`` once per trajectory.  The fit
is a 51-sample SVD and stays on the host with the same sklearn call; the projection of a whole sweep's trajectories
([N, L, C, H, W] on the device, e.g. what ``generate_trajectories_batched`` returns) is one streaming CUDA pass
(``dtraj_project``: every element read once).
"""
import numpy as np
import torch

from .. import _lib


def process_trajectory(traj):
    """analyze_trajectories.py:66-68: list of frames (tensors or (tensor, t) tuples) -> [L, D] float32 features."""
    frames = [f[0] if isinstance(f, tuple) else f for f in traj]
    return np.stack([f.detach().cpu().numpy().reshape(-1) for f in frames])


def fit_reference_pca(reference_trajectory, n_components=3):
    """analyze_trajectories.py:70-80: PCA fitted on the reference trajectory (teacher, first guidance scale).
    ``reference_trajectory``: list of frames, or an [L, D] / [L, C, H, W] array / tensor."""
    from sklearn.decomposition import PCA
    if isinstance(reference_trajectory, (list, tuple)):
        feats = process_trajectory(reference_trajectory)
    else:
        a = reference_trajectory.detach().cpu().numpy() if torch.is_tensor(reference_trajectory) else np.asarray(reference_trajectory)
        feats = a.reshape(a.shape[0], -1)
    pca = PCA(n_components=n_components)
    pca.fit(feats)
    return pca


def project_trajectories(trajectories, pca):
    """``pca.transform`` of every trajectory of a device tensor [N, L, ...] (CUDA fp32) -> device tensor
    [N, L, n_components]; row n equals ``pca.transform(process_trajectory(trajectory n))`` (analyze_trajectories.py:100)."""
    if trajectories.device.type != "cuda":
        raise _lib.DtrajError("project_trajectories runs on CUDA only (no CPU fallback)")
    lib = _lib.load()
    x = trajectories.to(torch.float32).contiguous()
    N, L = x.shape[0], x.shape[1]
    D = int(np.prod(x.shape[2:]))
    comps64 = np.asarray(pca.components_, np.float64)
    K = comps64.shape[0]
    if comps64.shape[1] != D:
        raise ValueError(f"PCA fitted on {comps64.shape[1]} features, trajectories have {D}")
    off = comps64 @ np.asarray(pca.mean_, np.float64)                       # sklearn: X @ C^T - mean @ C^T
    comps = torch.from_numpy(comps64.astype(np.float32)).to(x.device).contiguous()
    offset = torch.from_numpy(off.astype(np.float32)).to(x.device)
    out = torch.empty(N, L, K, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.dtraj_project(_lib.ptr(x), N * L, D, _lib.ptr(comps), _lib.ptr(offset), K, _lib.ptr(out), _lib.stream_ptr()))
    return out
