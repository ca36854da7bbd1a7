"""GPU hardware probe (see csrc/probe.cuh): TMA box over the permuted dimensions {c, x, image, y} of an NHWC fp16 map ->
shared-memory rows [halo row][image][halo column], zero-filled outside the 8x8 images?"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distillation_trajectories_b200 import _lib

from distillation_trajectories_b200 import build as _build
lib = C.CDLL(_build.build(probes=True))      # the probes live in libdtraj_probes.so (-DDTRAJ_PROBES), built on the CPU box before gpurun
lib.dtraj_probe_tma_permuted.restype = C.c_int
lib.dtraj_probe_tma_permuted.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
for n_img, img0 in ((4, 0), (4, 2), (3, 2)):
    out = np.zeros(200, np.float32)
    _lib.check(lib.dtraj_probe_tma_permuted(n_img, img0, out.ctypes.data_as(C.c_void_p)))
    want = np.zeros(200, np.float32)
    for hy in range(10):
        for im in range(2):
            for hx in range(10):
                y, x, n = hy - 1, hx - 1, img0 + im
                if 0 <= y < 8 and 0 <= x < 8 and n < n_img:
                    want[(hy * 2 + im) * 10 + hx] = n * 64 + y * 8 + x + 1
    print(f"n_img={n_img} img0={img0}: layout [halo row][image][halo col] {'OK' if np.array_equal(out, want) else 'MISMATCH'}; "
          f"first rows {out[:24].tolist()}")
