"""Attach the CUDA engine to an importable ``metran`` package without forking it.

``Metran._init_kalmanfilter`` instantiates the module-level name ``SPKalmanFilter`` that
metran/metran.py imports at line 23 and uses at line 243; rebinding that name (and the one
in metran/kalmanfilter.py) to this package's class makes ``Metran.solve()``,
``get_state_means()``, ``get_simulation()``, ``mask_observations()`` ... run on the GPU
with their signatures unchanged ("How a new engine attaches").
"""
import numpy as np

from .kalmanfilter import SPKalmanFilter, kalmansmoother_cuda, seqkalmanfilter_cuda

_saved = {}


def _batched_get_covariance(self, x0, f, callback, epsilon=None, diff="forward"):
    """Replacement for BaseSolver._get_covariance (metran/solver.py:65-140) on the CUDA engine: the
    O(nparam^2) objective values of the finite-difference Hessian come from ONE batched launch
    (dfm.batched_covariance keeps scipy's points and difference formulas).  ``f`` is the solver's
    objfunction (solver.py:42-63): callback(p) then Metran.get_mle (metran.py:605-622), i.e.
    _get_matrices -> set_matrices -> run_filter -> get_mle; the batch builds the same matrices
    per point on the host and filters them together."""
    from .dfm import batched_covariance

    mt = self.mt
    kf = mt.kf
    if not isinstance(kf, SPKalmanFilter) or kf.filtermethod != kf._filtermethod:
        return _saved["_get_covariance"](self, x0, f, callback, epsilon, diff)
    from pandas import Series

    def fun_batch(P):
        mats = []
        for p in np.atleast_2d(P):
            pp = callback(np.array(p, dtype=np.float64)) if callback is not None else p
            mats.append(mt._get_matrices(Series(pp, index=mt.parameters.index)))
        A, Q, H, R = (np.stack([np.asarray(m[i], dtype=np.float64) for m in mats]) for i in range(4))
        return kf.loglik_batch(A, Q, H, R)

    return batched_covariance(fun_batch, np.array(x0, dtype=np.float64), epsilon=epsilon, diff=diff)


def install(metran_module=None):
    """Rebind metran's Kalman object to the CUDA-backed one; returns the metran module."""
    if metran_module is None:
        import metran as metran_module
    mt_mod = metran_module.metran
    kf_mod = metran_module.kalmanfilter
    if "SPKalmanFilter" not in _saved:
        _saved["SPKalmanFilter"] = kf_mod.SPKalmanFilter
        _saved["seqkalmanfilter"] = getattr(kf_mod, "seqkalmanfilter", None)
        _saved["kalmansmoother"] = kf_mod.kalmansmoother
        sv_mod = getattr(metran_module, "solver", None)
        if sv_mod is not None and hasattr(sv_mod, "BaseSolver"):
            _saved["_get_covariance"] = sv_mod.BaseSolver._get_covariance
            sv_mod.BaseSolver._get_covariance = _batched_get_covariance
    mt_mod.SPKalmanFilter = SPKalmanFilter
    kf_mod.SPKalmanFilter = SPKalmanFilter
    kf_mod.seqkalmanfilter_cuda = seqkalmanfilter_cuda
    kf_mod.kalmansmoother_cuda = kalmansmoother_cuda
    return metran_module


def uninstall(metran_module=None):
    """Restore the reference classes."""
    if not _saved:
        return
    if metran_module is None:
        import metran as metran_module
    metran_module.metran.SPKalmanFilter = _saved["SPKalmanFilter"]
    metran_module.kalmanfilter.SPKalmanFilter = _saved["SPKalmanFilter"]
    if "_get_covariance" in _saved:
        metran_module.solver.BaseSolver._get_covariance = _saved["_get_covariance"]
    _saved.clear()
