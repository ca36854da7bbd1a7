"""GPU probe: error of the tcgen05 conv (TF32 one pass, 3xTF32) against an f64 convolution as K grows.
Prints max|err|/max|ref| and mean signed err/max|ref| (a bias means round-toward-zero accumulation)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_gpu_conv import run_conv, reference, tf32_trunc
from distillation_trajectories_b200 import _lib

rng = np.random.RandomState(0)
print("prec   cin  K     positive_inputs  max_err/max  mean_signed/max  rms/max")
for cin in (32, 64, 128, 256, 512):
    for positive in (False, True):
        n, H, cout = 2, 8, 64
        c0 = cin if cin <= 256 else 256
        c1 = cin - c0
        x0 = rng.randn(n, c0, H, H).astype(np.float32)
        x1 = rng.randn(n, c1, H, H).astype(np.float32) if c1 else None
        if positive:
            x0 = np.abs(x0); x1 = None if x1 is None else np.abs(x1)
        w = (rng.randn(cout, cin, 3, 3) / np.sqrt(cin * 9)).astype(np.float32)
        if positive:
            w = np.abs(w)
        b = np.zeros(cout, np.float32)
        for prec, nm in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_TF32, "tf32"), (_lib.PREC_TF32X3, "tf32x3")):
            a0, a1 = x0, x1
            if prec == _lib.PREC_TF32:
                a0 = tf32_trunc(x0); a1 = None if x1 is None else tf32_trunc(x1)
            got = run_conv(prec, a0, a1, w, b, 3, False, None)
            want = reference(a0, a1, w, b, 3, False, None, round_w=(prec == _lib.PREC_TF32))
            e = (got - want) / np.abs(want).max()
            print(f"{nm:6s} {cin:4d} {cin*9:5d} {str(positive):5s}            {np.abs(e).max():.3e}    {e.mean():+.3e}      {np.sqrt((e**2).mean()):.3e}")
