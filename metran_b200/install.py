"""Attach the CUDA engine to an importable ``metran`` package without forking it.

``Metran._init_kalmanfilter`` instantiates the module-level name ``SPKalmanFilter`` that
metran/metran.py imports at line 23 and uses at line 243; rebinding that name (and the one
in metran/kalmanfilter.py) to this package's class makes ``Metran.solve()``,
``get_state_means()``, ``get_simulation()``, ``mask_observations()`` ... run on the GPU
with their signatures unchanged ("How a new engine attaches").
"""
from .kalmanfilter import SPKalmanFilter, kalmansmoother_cuda, seqkalmanfilter_cuda

_saved = {}


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
    _saved.clear()
