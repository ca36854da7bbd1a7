"""GPU: how much of the stated trajectory tolerance (|d| <= 1e-3 |ref| + 1e-4 max|ref|) each arithmetic mode uses on S1
(utils.diffusion.p_sample_loop), BASELINE configs[0] shape: teacher 1x16x16, 50 steps, batch 64; w = 1.0 and CFG w = 3.0.
Prints max over elements and frames of |d| / tol, the frame where it peaks and max|ref| (S1's update rule lets |x| grow)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import samplers as osmp
from distillation_trajectories_b200 import sampling, set_precision
from distillation_trajectories_b200.utils import diffusion
from helpers import Cfg, make_model, oracle_fn

sampling.set_noise_device("cpu")        # the oracle draws from the CPU generator
stack = lambda tr: torch.stack(list(tr)).cpu().numpy().astype(np.float64)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for C, H, sf, wseed in ((1, 16, 1.0, 0), (3, 32, 1.0, 0), (1, 16, 0.5, 1050)):
    cfg = Cfg(C, H, 50)
    model = make_model(cfg, sf, wseed, stress=False, device="cuda")
    b = B if H == 16 else max(2, B // 8)
    for w, seed in ((1.0, 123), (3.0, 124)):
        torch.manual_seed(seed)
        _, want = osmp.s1_p_sample_loop(oracle_fn(model), (b, C, H, H), 50, osmp.diffusion_params(50), 50, w)
        want = stack(want)
        tol = 1e-3 * np.abs(want) + 1e-4 * np.abs(want).max()
        line = f"{C}x{H} sf={sf} B={b} w={w} max|ref|={np.abs(want).max():.1f} (last frame {np.abs(want[-1]).max():.1f}):"
        for prec in ("tf32x3", "tf32", "f16"):
            set_precision(prec, "S1")
            torch.manual_seed(seed)
            _, got = diffusion.p_sample_loop(model, (b, C, H, H), 50, diffusion.get_diffusion_params(50, cfg), device="cuda",
                                             config=cfg, track_trajectory=True, guidance_scale=w)
            r = np.abs(stack(got) - want) / tol
            line += f"  {prec} {r.max():.3f} (frame {int(np.argmax(r.reshape(r.shape[0], -1).max(1)))})"
        print(line, flush=True)
set_precision("tf32x3", "S1")
