"""Oracle against the LIVE reference on fresh random cases (build container only; skipped when
/root/reference is absent).  The committed golden fixtures cover the GPU box; this test keeps the
oracle honest whenever the reference tree is at hand."""
import sys

import numpy as np
import pytest
from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import _refload  # noqa: E402

from oracle import kalman_oracle as ko  # noqa: E402

pytestmark = pytest.mark.skipif(not _refload.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def refkf():
    return _refload.load_reference().kalmanfilter


@pytest.mark.parametrize("seed,n,m,T,miss,full", [(1, 8, 5, 150, 0.3, False), (2, 6, 5, 200, 0.8, False),
                                                 (3, 12, 20, 40, 0.1, True), (4, 3, 2, 90, 0.0, False)])
def test_oracle_matches_live_reference(refkf, seed, n, m, T, miss, full):
    import pandas as pd

    rng = np.random.default_rng(seed)
    phi = rng.uniform(0.6, 0.99, n)
    A = np.diag(phi) + (0.03 * rng.standard_normal((n, n)) if full else 0.0)
    L = 0.1 * rng.standard_normal((n, n))
    Q = np.diag(1 - phi ** 2) + (L @ L.T if full else 0.0)
    H = rng.standard_normal((m, n)) / np.sqrt(n)
    R = rng.uniform(0.05, 0.3, m)
    y = rng.standard_normal((T, m))
    y[rng.random((T, m)) < miss] = np.nan
    kf = refkf.SPKalmanFilter(engine="numba")
    kf.set_observations(pd.DataFrame(y))
    kf.set_matrices(A, Q, H, R)
    kf.run_smoother()
    obs, idx, cnt = ko.pack_observations(y)
    assert np.array_equal(obs, kf.observations) and np.array_equal(idx, kf.observation_indices)
    assert np.array_equal(cnt, kf.observation_count)
    sig, det, ns, fm, fc, pm, pc = ko.seqkalmanfilter(obs, A, Q, H, R, idx, cnt, np.zeros(n), np.eye(n))
    assert ns == len(kf.sigmas) and np.array_equal(sig[:ns], kf.sigmas)
    np.testing.assert_allclose(det[:ns], kf.detfs, rtol=0, atol=4e-15)
    for a, b in ((fm, kf.filtered_state_means), (fc, kf.filtered_state_covariances),
                 (pm, kf.predicted_state_means), (pc, kf.predicted_state_covariances)):
        assert np.array_equal(a, b)
    for w in (0, 1, 2):
        assert abs(ko.get_mle(sig[:ns], det[:ns], cnt, w) - kf.get_mle(warmup=w)) < 1e-11
        assert abs(ko.filter_loglik(y, A, Q, H, R, warmup=w)[0] - kf.get_mle(warmup=w)) < 1e-10
    sm, sc = ko.kalmansmoother(fm, fc, pm, pc, A)
    np.testing.assert_allclose(sm, kf.smoothed_state_means, rtol=0, atol=1e-13)
    np.testing.assert_allclose(sc, kf.smoothed_state_covariances, rtol=0, atol=1e-13)
    means, var = ko.simulate(sm, sc, H)
    rm, rv = kf.simulate(H)
    np.testing.assert_allclose(means, np.asarray(rm), atol=1e-13)
    np.testing.assert_allclose(var, np.asarray(rv), atol=1e-13)
