"""Import the UNMODIFIED reference: from /root/reference (build container) or from the byte-for-byte
copies ``oracle/sync_ref.py`` stages under the git-ignored ``oracle/_ref/`` (they travel to the GPU box).

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).  Used by ``oracle/make_golden.py`` and
``tests/test_oracle_vs_reference.py`` to pin the oracle restatement against the real
reference code, and by ``bench.py --impl reference`` / its ``cpu_baseline`` leg to TIME the
reference itself on the host cores.  ``available()`` says whether either tree can be used;
``DTRAJ_REFERENCE_ROOT`` overrides the choice.

The only imports the reference needs that this image lacks are plotting packages
(SURVEY.md 8c); they are replaced by inert stub modules.  ``analysis/__init__.py``
drags in every analysis sub-package (umap, sklearn plots ...), so ``analysis`` and
``analysis.metrics`` are registered as bare namespace packages instead.
"""
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root():
    env = os.environ.get("DTRAJ_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, "models.py")):
            return cand
    return "/root/reference"


REF_ROOT = _pick_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models.py"))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")


_loaded = None


def load():
    """Returns a namespace with the reference modules of the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors",
                 "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "umap", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    # The reference uses top-level module names (models, utils, analysis, config); they
    # stay registered in sys.modules because the reference imports some of them lazily
    # inside functions (analysis/trajectory_engine.py:131).  The product package is
    # namespaced (distillation_trajectories_b200.*), so there is no clash.
    sys.path.insert(0, REF_ROOT)
    try:
        for pkg in ("analysis", "analysis.metrics"):
            mod = types.ModuleType(pkg)
            mod.__path__ = [os.path.join(REF_ROOT, *pkg.split("."))]
            sys.modules[pkg] = mod
        ns = types.SimpleNamespace()
        ns.models = importlib.import_module("models")
        ns.diffusion = importlib.import_module("utils.diffusion")
        ns.metric_transformations = importlib.import_module("utils.metric_transformations")
        ns.trajectory_metrics = importlib.import_module("analysis.metrics.trajectory_metrics")
        ns.time_dependent = importlib.import_module("analysis.metrics.time_dependent")
        ns.trajectory_engine = importlib.import_module("analysis.trajectory_engine")
        ns.trajectory_manager = importlib.import_module("utils.trajectory_manager")
    finally:
        sys.path.remove(REF_ROOT)
    _loaded = ns
    return ns


class RefConfig:
    """Duck-typed stand-in for config/config.py:5-95 with only the attributes the hot
    path reads; avoids the torchvision import."""

    def __init__(self, channels=1, image_size=16, timesteps=50, **kw):
        self.channels = channels
        self.image_size = image_size
        self.timesteps = timesteps
        self.sample_steps = timesteps
        self.teacher_steps = timesteps
        self.student_steps = timesteps
        self.beta_start = 1e-4
        self.beta_end = 0.02
        self.dropout = 0.3
        self.force_cpu = True
        self.mps_enabled = False
        self.progress_bar_leave = False
        self.progress_bar_position = 0
        self.trajectory_dir = "/tmp/dtraj_ref_trajectories"
        for k, v in kw.items():
            setattr(self, k, v)
