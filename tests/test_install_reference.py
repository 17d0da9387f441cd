"""Drop-in check against the UNMODIFIED reference facade (build container only).

`metran_b200.install()` rebinds the `SPKalmanFilter` name that metran/metran.py:243
instantiates; `Metran.solve()`, the getters and the mask path then run on this package's
Kalman object.  Needs /root/reference (absent on the GPU box -> skipped there); the device
library is replaced by tests/fakelib.py so the host-side mirror is what is under test.
"""
import os
import sys

import numpy as np
import pytest
from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import _refload  # noqa: E402

pytestmark = pytest.mark.skipif(not _refload.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def installed():
    from fakelib import FakeLib

    import metran_b200
    from metran_b200 import kalmanfilter
    from metran_b200.batched import KalmanEngine

    metran = _refload.load_reference()
    eng = KalmanEngine(lib=FakeLib(), device="cpu")
    kalmanfilter.set_default_engine(eng)
    metran_b200.install(metran)
    yield metran, eng
    metran_b200.uninstall(metran)
    kalmanfilter.set_default_engine(None)


def test_metran_solve_and_getters_through_installed_engine(installed, b21):
    metran, eng = installed
    mt = metran.Metran(_refload.load_example_series(), name="B21B0214")
    mt.solve(report=False)                      # default engine="numba" argument is accepted
    from metran_b200.kalmanfilter import SPKalmanFilter

    assert isinstance(mt.kf, SPKalmanFilter)
    # every objective evaluation of solve() was ONE log-likelihood kernel call; the history
    # kernel is deferred until a getter reads state arrays
    names = [c[0] for c in eng.lib.calls]
    # (on the compacted grid: 343 of the 6255 daily steps of B21B0214 carry an observation)
    assert names.count("mkf_dfm_loglik_compact") == mt.fit.nfev and "mkf_filter_history" not in names
    assert abs(mt.fit.obj_func - float(b21["obj_opt"])) < 1e-6
    np.testing.assert_allclose(mt.parameters["optimal"].values.astype(float), b21["alpha_opt"], rtol=1e-4)
    sel = b21["sel"]
    import pandas as pd

    p = pd.Series(b21["alpha_opt"], index=mt.parameters.index)
    np.testing.assert_allclose(mt.get_state_means(p=p).values[sel], b21["state_means"], atol=1e-9)
    assert "mkf_filter_history" in [c[0] for c in eng.lib.calls] and "mkf_smoother_project" in [c[0] for c in eng.lib.calls]
    np.testing.assert_allclose(mt.get_state_variances(p=p).values[sel], b21["state_variances"], atol=1e-9)
    np.testing.assert_allclose(mt.get_simulated_means(p=p).values[sel], b21["sim_means"], atol=1e-9)
    np.testing.assert_allclose(mt.get_simulated_variances(p=p).values[sel], b21["sim_variances"], atol=1e-9)
    np.testing.assert_allclose(mt.get_simulation("B21B0214005", p=p, alpha=0.05).values[sel],
                               b21["simulation_last_ci"], atol=1e-9)
    np.testing.assert_allclose(mt.decompose_simulation("B21B0214001", p=p).values[sel],
                               b21["decompose_first"], atol=1e-9)
    np.testing.assert_allclose(mt.get_state(0, p=p, alpha=0.05).values[sel], b21["state0_ci"], atol=1e-9)


def test_mask_path_through_installed_engine(installed, b21):
    """tests/test_metran.py:32-40 of the reference, with the value pinned."""
    metran, _ = installed
    mt = metran.Metran(_refload.load_example_series(), name="B21B0214")
    mt.solve(report=False)
    proj1 = mt.get_simulation("B21B0214005", alpha=0.05)
    oseries = mt.get_observations()
    mask = (0 * oseries).astype(bool)
    mask.loc["1997-8-28", "B21B0214005"] = True
    mt.mask_observations(mask)
    proj2 = mt.get_simulation("B21B0214005", alpha=0.05)
    mt.unmask_observations()
    assert (proj1 != proj2).any().any()
    np.testing.assert_allclose(proj2.values[int(b21["mask_row"])], b21["masked_sim_row"], atol=1e-6)
    np.testing.assert_allclose(proj2.values[int(b21["mask_row"])],
                               [5.92966842, 0.78340608, 11.07593077], atol=1e-6)


def test_hessian_of_the_reference_solver_is_batched(installed, b21):
    """install() also swaps BaseSolver._get_covariance (solver.py:65-140): same numbers as the
    reference's one-call-at-a-time loop, one kernel launch for all its points."""
    metran, eng = installed
    import importlib

    inst = importlib.import_module("metran_b200.install")

    mt = metran.Metran(_refload.load_example_series(), name="B21B0214")
    mt.solve(report=False)
    solver = mt.fit
    x0 = solver.result.x.copy()
    n0 = len(eng.lib.calls)
    cov = solver._get_covariance(x0.copy(), solver.objfunction, solver._array_todict)
    names = [c[0] for c in eng.lib.calls[n0:]]
    assert names == ["mkf_dfm_loglik_compact"], names                      # ONE launch for (n+1)^2 points
    n1 = len(eng.lib.calls)
    want = inst._saved["_get_covariance"](solver, x0.copy(), solver.objfunction, solver._array_todict)
    assert len(eng.lib.calls) - n1 == (x0.size + 1) ** 2            # the reference's loop, one launch each
    np.testing.assert_allclose(cov, want, rtol=1e-9, atol=1e-12)
    assert np.all(np.diag(cov) > 0)
