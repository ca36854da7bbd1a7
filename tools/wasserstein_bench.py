"""GPU: timing of the Wasserstein kernels (the library picks the warp form for K <= 1024 sampled elements, the block form above) and a
check of a few frames against an f64 sort.  (Round 1 also forced the block form everywhere through an environment knob; the knob was
removed in round 2.)
    python tools/wasserstein_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from distillation_trajectories_b200.analysis.metrics import trajectory_metrics as tm

def run(T, S, idx, idx_set):
    for _ in range(2):
        out = tm.wasserstein_frames(T, S, idx, idx_set)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = tm.wasserstein_frames(T, S, idx, idx_set)
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / 5

for N, L, D, K in ((2368, 51, 256, 256), (256, 51, 3072, 1000), (64, 51, 3072, 3072)):
    g = torch.Generator(device="cuda").manual_seed(7)
    T = torch.randn(N, L, D, device="cuda", generator=g)
    S = T + 0.3 * torch.randn(N, L, D, device="cuda", generator=g)
    idx = idx_set = None
    if K != D:
        rng = np.random.RandomState(0)
        idx = torch.from_numpy(np.stack([np.stack([rng.choice(D, K, replace=False) for _ in range(L)]) for _ in range(4)]).astype(np.int32))
        idx_set = torch.arange(N, dtype=torch.int32) % 4
    a, ta = run(T, S, idx, idx_set)
    # f64 reference on a few frames
    ref = []
    for n, i in ((0, 0), (N - 1, L - 1), (N // 2, 7)):
        sel = slice(None) if idx is None else idx[int(idx_set[n]), i].long().cuda()
        u, v = T[n, i][sel].double().sort().values, S[n, i][sel].double().sort().values
        ref.append(((u - v).abs().mean().item(), a[n, i].item()))
    print(f"N={N} L={L} D={D} K={K}: {ta:.3f} ms, (f64 reference, kernel) on three frames {ref}", flush=True)
