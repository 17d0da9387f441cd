"""metran_b200 -- B200-native batched Kalman filter / RTS smoother behind the Metran API.

The hot path of pastas/metran (metran/kalmanfilter.py filter + smoother, the inner
log-likelihood call of metran/solver.py) as hand-written sm_100a CUDA kernels behind a
C ABI (include/metran_b200.h), with a host-side mirror of the reference's Kalman object.
Importing this package does not need a GPU; running anything does (no CPU fallback).
"""
__version__ = "0.1.0"

from .install import install, uninstall  # noqa: F401
from .kalmanfilter import (SPKalmanFilter, kalmansmoother_cuda, pack_observations,  # noqa: F401
                           seqkalmanfilter_cuda)


def __getattr__(name):
    if name in ("KalmanEngine", "default_engine"):
        from . import batched

        return getattr(batched, name)
    if name in ("DynamicFactorModel", "BatchedObjective"):
        from . import dfm

        return getattr(dfm, name)
    if name in ("fit_dfm_batch", "BatchFitResult"):
        from . import fit

        return getattr(fit, name)
    raise AttributeError(name)
