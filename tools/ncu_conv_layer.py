"""ncu target: one conv layer shape through dtraj_bench_conv (3 launches).  args: c0 c1 cout H ksize flags [rows]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distillation_trajectories_b200 import _lib
lib = _lib.load()
if len(sys.argv) < 7:
    sys.exit(__doc__)
c0, c1, cout, H, k, fl = (int(a) for a in sys.argv[1:7])
rows = int(sys.argv[7]) if len(sys.argv) > 7 else 3840
ms = C.c_float()
_lib.check(lib.dtraj_bench_conv(_lib.PREC_TF32, c0, c1, cout, rows, H, k, fl, 1, 0, C.byref(ms)))
print("ms", ms.value)
