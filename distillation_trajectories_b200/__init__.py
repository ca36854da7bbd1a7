"""B200-native trajectory hot path of henriChevreux/distillation_trajectories.

Same Python entry points as the reference for the sampler, the trajectory engine and the
trajectory metrics, executed by libdtraj.so (hand-written sm_100a CUDA behind a C ABI,
include/dtraj.h).  See DESIGN.md and INTEGRATION.md.
"""
from ._lib import DtrajError, LIB_PATH  # noqa: F401
from .engine import UNetEngine, TrajectorySampler, set_precision, get_precision, umma_error_flag  # noqa: F401

__all__ = ["DtrajError", "LIB_PATH", "UNetEngine", "TrajectorySampler", "set_precision", "get_precision",
           "umma_error_flag"]
