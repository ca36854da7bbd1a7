"""GPU: kernel-only timing of the conv layer shapes of the bench workload (teacher + student, 3840 rows).
    python tools/conv_bench.py [rows] [precision]
Prints per layer: ms and TFLOP/s (real flops).  (Round 1 also timed operand knock-outs -- no weight loads / no pixel
loads / no epilogue traffic, profiles/r01f_conv_knockout.txt; those debug paths were removed from the kernels in round 2.)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distillation_trajectories_b200 import _lib

lib = _lib.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
prec = _lib.PRECISIONS[sys.argv[2]] if len(sys.argv) > 2 else _lib.PREC_TF32
dbgs = [0]

# (name, c0, c1, cout, H, ksize, flags)  flags: 1 relu, 4 residual
def layers(b, m, H):
    return [("enc1.conv2", b, 0, b, H, 3, 5), ("enc2.res", b, 0, m, H // 2, 1, 0), ("enc2.conv1", b, 0, m, H // 2, 3, 1),
            ("enc2.conv2", m, 0, m, H // 2, 3, 5), ("enc3.conv1", m, 0, m, H // 4, 3, 1), ("enc4.conv1", m, 0, m, H // 8, 3, 1),
            ("bott.conv1", m, 0, m, H // 16, 3, 1), ("dec3.conv1", m, m, m, H // 8, 3, 1), ("dec2.conv1", m, m, m, H // 4, 3, 1),
            ("dec1.res", m, m, b, H // 2, 1, 0), ("dec1.conv1", m, m, b, H // 2, 3, 1), ("dec1.conv2", b, 0, b, H // 2, 3, 5)]

for tag, b, m in (("teacher", 128, 256), ("student0.5", 64, 128)):
    print(f"== {tag} rows={rows}")
    for name, c0, c1, cout, H, k, fl in layers(b, m, 16):
        taps = 1 if (k == 1 or H == 1) else 9
        flops = 2.0 * rows * H * H * cout * (c0 + c1) * taps
        line = f"{name:11s} M={rows*H*H:8d} N={cout:3d} K={(c0+c1)*taps:5d}"
        for d in dbgs:
            ms = C.c_float()
            _lib.check(lib.dtraj_bench_conv(prec, c0, c1, cout, rows, H, k, fl, 10, d, C.byref(ms)))
            line += f" | dbg{d}: {ms.value*1e3:8.1f} us {flops/ms.value/1e9:7.1f} TF/s"
        print(line, flush=True)
print("umma_error", lib.dtraj_debug_umma_error())
