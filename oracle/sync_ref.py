"""Stage the UNMODIFIED reference hot-path files under the git-ignored ``oracle/_ref/`` so that they travel
to the GPU box (gpurun snapshots /root/repo; /root/reference does not exist there).

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py): ``oracle/_ref`` is what ``bench.py --impl reference``
and the ``cpu_baseline`` leg time (``cpu_baseline.kind == "reference"``).  Nothing in the product package imports it.
The files are byte-for-byte copies made at build time (``__graft_entry__.build()`` calls ``sync()`` whenever
/root/reference is present) -- the equivalent of ``pip install --target baseline/_ref`` for a reference that is a
flat script tree without packaging metadata; they never enter the git history (``.gitignore``: ``oracle/_ref/``).

    python -m oracle.sync_ref            # copy, print what was staged
"""
import hashlib
import os
import shutil
import sys

SRC = os.environ.get("DTRAJ_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# the files SURVEY.md 8(a)/8(c) put on the hot path, plus the packages' __init__ files they need.
# analysis/__init__.py and analysis/metrics/__init__.py are NOT copied: they import every plotting sub-package
# (umap, seaborn ...); oracle/refload.py registers bare namespace packages instead.
FILES = [
    "models.py",
    "config/__init__.py", "config/config.py",
    "utils/__init__.py", "utils/diffusion.py", "utils/trajectory_manager.py", "utils/metric_transformations.py",
    "analysis/trajectory_engine.py",
    "analysis/metrics/trajectory_metrics.py", "analysis/metrics/time_dependent.py",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def sync(verbose=False):
    """Copy FILES from SRC to DST (only when SRC exists).  Returns the list of staged files, [] when SRC is absent."""
    if not os.path.isfile(os.path.join(SRC, "models.py")):
        return []
    staged = []
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            raise FileNotFoundError(f"reference file missing: {s}")
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.isfile(d) or _sha(s) != _sha(d):
            shutil.copyfile(s, d)
        staged.append(rel)
        if verbose:
            print(f"{_sha(d)[:16]}  {rel}")
    with open(os.path.join(DST, "MANIFEST"), "w") as f:
        f.write("".join(f"{_sha(os.path.join(DST, rel))}  {rel}\n" for rel in staged))
    return staged


def staged():
    return os.path.isfile(os.path.join(DST, "models.py"))


if __name__ == "__main__":
    got = sync(verbose=True)
    print(f"staged {len(got)} files under {DST}" if got else f"{SRC} not present: nothing staged")
    sys.exit(0)
