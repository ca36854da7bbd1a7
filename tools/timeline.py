"""GPU: SM-clock timeline of CTA 0 in the 256-column fused-residual conv layers at 8x8 (enc2.conv2 +res +pool, dec2.conv2 +res) of the
teacher at the benchmark shape -- where a tile's time goes between the MMA issuer and the epilogue warps.

    python tools/timeline.py [seeds]

Needs the probe build (csrc/libdtraj_probes.so, -DDTRAJ_PROBES: build it on the CPU box before gpurun).  Events (cycles, relative to the
tile's first stamp):  issuer: wait for a drained accumulator -> first MMA ... commit of the tile;  epilogue warps 2 (chunks 0,2,4,6)
and 6 (chunks 1,3,5,7) of lane quarter 2: tile start, accumulator ready, then per chunk: TMEM loaded / computed / emitted."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distillation_trajectories_b200 import _lib, build as _build

_lib.LIB_PATH = _build.build(probes=True)
import numpy as np
import torch

import bench
from distillation_trajectories_b200 import grid

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 592
dev = torch.device("cuda", 0)
ck = grid.stage_chunk(list(range(seeds)), bench.Cfg, bench.GUIDANCE, dev)
m = bench.make_model(bench.Cfg, 1.0, 0, dev)
grid.run_chunk(m, [m], ck, dev, "f16")
torch.cuda.synchronize()
lib = _lib.load()
tl = np.zeros((2, 16, 3, 16), dtype=np.int64)
lib.dtraj_probe_timeline.restype = C.c_int
lib.dtraj_probe_timeline.argtypes = [C.c_void_p, C.c_int64]
_lib.check(lib.dtraj_probe_timeline(tl.ctypes.data, tl.size))
names = ["enc2.conv2 +res +pool", "dec2.conv2 +res"]
for s in range(2):
    print(f"== {names[s]}: CTA 0, cycles relative to the issuer's first stamp of tile 2")
    t0 = tl[s, 2, 0, 0]
    for t in range(2, 8):
        iss = tl[s, t, 0, :3] - t0
        print(f"tile {t}: issuer  start {iss[0]:7d}  acc drained {iss[1]:7d}  committed {iss[2]:7d}   (K loop issue {iss[2] - iss[1]} cycles)")
        for w in (1, 2):
            e = tl[s, t, w] - t0
            chunks = "  ".join(f"[ld {e[2 + 3 * k]:7d} cmp {e[3 + 3 * k]:7d} emit {e[4 + 3 * k]:7d}]" for k in range(4))
            print(f"        warp {2 + 4 * (w - 1)}: start {e[0]:7d}  acc ready {e[1]:7d}  {chunks}")
    per_tile = (tl[s, 7, 0, 0] - tl[s, 2, 0, 0]) / 5
    print(f"   tile period {per_tile:.0f} cycles")
