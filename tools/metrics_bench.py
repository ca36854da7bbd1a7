"""GPU: bandwidth of the metric kernels on BASELINE config 5-shaped chunks ([N, 50, 3, 32, 32] fp32, synthetic
random-walk teacher / noisy student generated on the device).  Prints GB/s = 2*N*L*D*4 / time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm

def make(N, L, D, chunk):
    g = torch.Generator(device="cuda").manual_seed(1234 + chunk)
    T = torch.randn(N, 1, D, device="cuda", generator=g) + 0.1 * torch.cumsum(torch.randn(N, L, D, device="cuda", generator=g), dim=1)
    S = T + 0.05 * torch.randn(N, L, D, device="cuda", generator=g)
    return T.contiguous(), S.contiguous()

for N, L, D in ((4096, 50, 3072), (16384, 50, 3072), (65536, 51, 256), (2048, 51, 256)):
    T, S = make(N, L, D, 0)
    for _ in range(3):
        tm.pair_reductions(T, S)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tm.pair_reductions(T, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = 2 * N * L * D * 4 / 1e9
    print(f"pair_reductions N={N} L={L} D={D}: {gb:.2f} GB in {ms:.3f} ms = {gb/ms*1e3:.0f} GB/s", flush=True)
    del T, S

# PCA projection (dtraj_project): one pass over [N, L, D], K = 3 directions
import numpy as np
from distillation_trajectories_b200.analysis import trajectory_pca as tp
for N, L, D in ((8192, 50, 3072), (65536, 51, 256)):
    T, S = make(N, L, D, 1)
    del S
    pca = tp.fit_reference_pca(T[0].cpu().numpy())
    for _ in range(3):
        tp.project_trajectories(T, pca)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tp.project_trajectories(T, pca)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = N * L * D * 4 / 1e9
    print(f"project_trajectories N={N} L={L} D={D}: {gb:.2f} GB in {ms:.3f} ms = {gb/ms*1e3:.0f} GB/s", flush=True)
    del T
