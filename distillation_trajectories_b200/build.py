"""Build libdtraj.so in-tree with nvcc for sm_100a (no torch involved).

    python -m distillation_trajectories_b200.build [--force]

The shared object lands next to the sources (csrc/libdtraj.so); it is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdtraj.so")
STAMP = os.path.join(CSRC, ".libdtraj.stamp")
SOURCES = ["dtraj.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + ["../../include/dtraj.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False, probes=False):
    """``probes=True`` builds csrc/libdtraj_probes.so instead: the same library plus the hardware probes of csrc/probe.cuh
    (-DDTRAJ_PROBES; tools/probe_*.py).  The product library never contains them."""
    dig = _digest() + ("+probes" if probes else "")
    lib = LIB.replace("libdtraj.so", "libdtraj_probes.so") if probes else LIB
    stamp = STAMP + (".probes" if probes else "")
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return lib
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-DDTRAJ_PROBES"] if probes else []) + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, probes="--probes" in sys.argv))
