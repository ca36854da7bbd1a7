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
# optional: the conv layer to record, as "<flag mask> <flag value> <coutp> <W>" (csrc/conv_simt.cuh flag bits), and the size factor
sel = [int(a, 0) for a in sys.argv[2:6]] if len(sys.argv) >= 6 else None
sf = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
H = int(sys.argv[7]) if len(sys.argv) > 7 else 16
dev = torch.device("cuda", 0)
cfg = bench.Cfg if H == 16 else bench.Cfg32
ck = grid.stage_chunk(list(range(seeds)), cfg, bench.GUIDANCE if H == 16 else [7.5], dev)
m = bench.make_model(cfg, sf, 0 if sf == 1.0 else 1000 + int(sf * 100), dev)
if sel:
    _lib.load().dtraj_probe_timeline_select.argtypes = [C.c_int] * 4
    _lib.check(_lib.load().dtraj_probe_timeline_select(*sel))
grid.run_chunk(m, [m], ck, dev, "f16")
torch.cuda.synchronize()
lib = _lib.load()
tl = np.zeros((2, 16, 3, 16), dtype=np.int64)
lib.dtraj_probe_timeline.restype = C.c_int
lib.dtraj_probe_timeline.argtypes = [C.c_void_p, C.c_int64]
_lib.check(lib.dtraj_probe_timeline(tl.ctypes.data, tl.size))
print(f"== k_conv_umma_t layer {sel or 'enc2.conv2 +res +pool'}: CTA 0, cycles relative to the issuer's first stamp of tile 2")
t0 = tl[0, 2, 0, 0]
for t in range(2, 8):
    iss = tl[0, t, 0, :3] - t0
    print(f"tile {t}: issuer  start {iss[0]:7d}  acc drained {iss[1]:7d}  committed {iss[2]:7d}   (K loop issue {iss[2] - iss[1]} cycles)")
    for w in (1, 2):
        e = tl[0, t, w] - t0
        chunks = "  ".join(f"[ld {e[2 + 3 * k]:7d} cmp {e[3 + 3 * k]:7d} emit {e[4 + 3 * k]:7d}]" for k in range(4))
        print(f"        warp {2 + 4 * (w - 1)}: start {e[0]:7d}  acc ready {e[1]:7d}  {chunks}")
print(f"   tile period {(tl[0, 7, 0, 0] - tl[0, 2, 0, 0]) / 5:.0f} cycles")
print("== enc1 fused (k_enc1_f16): CTA 0, cycles relative to the issuer's first stamp of tile 4")
t0 = tl[1, 4, 0, 0]
for t in range(4, 10):
    i_ = tl[1, t, 0] - t0
    m_ = tl[1, t, 1] - t0
    e_ = tl[1, t, 2] - t0
    print(f"tile {t}: issuer  top {i_[0]:7d}  acc free {i_[1]:7d}  chunk0 [halo ready {i_[2]:7d} issued {i_[3]:7d}]  chunk1 [halo ready {i_[4]:7d} issued {i_[5]:7d}]  end {i_[6]:7d}")
    print(f"        mid w0:  top {m_[0]:7d}  D1 ready {m_[1]:7d}  patch ready {m_[12]:7d}  A1 gathered {m_[2]:7d}  " +
          "  ".join(f"chunk{c} [ld {m_[3 + 4 * c]:7d} buf free {m_[4 + 4 * c]:7d} written {m_[5 + 4 * c]:7d} barrier {m_[6 + 4 * c]:7d}]" for c in range(2)) + f"  end {m_[11]:7d}")
    print(f"        epi w2:  top {e_[0]:7d}  acc ready {e_[1]:7d}  " + "  ".join(f"[ld {e_[2 + 3 * k]:7d} cmp {e_[3 + 3 * k]:7d} pool {e_[4 + 3 * k]:7d}]" for k in range(4)))
print(f"   tile period {(tl[1, 9, 0, 0] - tl[1, 4, 0, 0]) / 5:.0f} cycles")
