"""GPU probe: which shared-memory rows does tcgen05.mma fetch for an A descriptor with a shifted start address
and a non-1024-multiple group stride (see csrc/probe.cuh)?  Prints, per configuration, whether every output row m
saw row  start + (m // 8) * (SBO / 128) + m % 8."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distillation_trajectories_b200 import _lib
from distillation_trajectories_b200 import build as _build
lib = C.CDLL(_build.build(probes=True))      # the probes live in libdtraj_probes.so (-DDTRAJ_PROBES), built on the CPU box before gpurun
lib.dtraj_probe_umma_view.restype = C.c_int
lib.dtraj_probe_umma_view.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
for sbo in (1024, 1280, 1792):
    for start in (0, 1, 3, 8, 11, 21):
        for mode in (0, 1):
            out = np.zeros(128, np.float32)
            _lib.check(lib.dtraj_probe_umma_view(256, start, sbo, mode, out.ctypes.data_as(C.c_void_p)))
            want = np.array([start + (m // 8) * (sbo // 128) + m % 8 for m in range(128)], np.float32)
            want[want >= 256] = -1                      # beyond the 256-row tile: not judged
            ok = np.array_equal(out[want >= 0], want[want >= 0])
            print(f"SBO={sbo:5d} start_row={start:2d} base_offset_field={'(addr>>7)&7' if mode else '0':11s} -> {'OK ' if ok else 'MISMATCH'} "
                  f"first rows got {out[:10].astype(int).tolist()} want {want[:10].astype(int).tolist()}", flush=True)
