"""GPU: cost of the fused epilogue tails on the layers that use them (kernel-only, zero data)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distillation_trajectories_b200 import _lib
lib = _lib.load()
rows = 3840
RELU, RES, POOL, NOST, RESX, FIN = 1, 4, 8, 16, 32, 64
cases = [("enc1.conv2 unfused", 128, 0, 128, 16, RELU | RES), ("enc1.conv2 resx", 128, 0, 128, 16, RELU | RESX),
         ("enc1.conv2 resx+pool", 128, 0, 128, 16, RELU | RESX | POOL), ("enc1.conv2 resx+pool+nostore", 128, 0, 128, 16, RELU | RESX | POOL | NOST),
         ("enc1.conv2 relu only", 128, 0, 128, 16, RELU), ("enc1.conv2 relu nostore", 128, 0, 128, 16, RELU | NOST),
         ("enc2.conv2 unfused", 256, 0, 256, 8, RELU | RES), ("enc2.conv2 pool", 256, 0, 256, 8, RELU | RES | POOL),
         ("dec1.conv2 unfused", 128, 0, 128, 8, RELU | RES), ("dec1.conv2 final", 128, 0, 128, 8, RELU | RES | FIN),
         ("dec1.conv2 final+nostore", 128, 0, 128, 8, RELU | RES | FIN | NOST)]
for name, c0, c1, cout, H, fl in cases:
    ms = C.c_float()
    _lib.check(lib.dtraj_bench_conv(_lib.PREC_TF32, c0, c1, cout, rows, H, 3, fl, 10, 0, C.byref(ms)))
    flops = 2.0 * rows * H * H * cout * (c0 + c1) * 9
    print(f"{name:32s} {ms.value*1e3:8.1f} us {flops/ms.value/1e9:7.1f} TF/s", flush=True)
