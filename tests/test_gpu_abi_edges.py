"""Edge cases and error behaviour of the C ABI on the device."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import kalman_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from metran_b200.batched import KalmanEngine

    return KalmanEngine()


def _ssm(rng, n, m, T, miss=0.0):
    phi = rng.uniform(0.7, 0.98, n)
    A, Q = np.diag(phi), np.diag(1 - phi ** 2)
    H = rng.standard_normal((m, n)) / np.sqrt(n)
    R = rng.uniform(0.05, 0.2, m)
    y = rng.standard_normal((T, m))
    y[rng.random((T, m)) < miss] = np.nan
    return y, A, Q, H, R


def test_return_codes(eng):
    from metran_b200 import _lib

    L = eng.lib
    z = ctypes.c_void_p(0)
    d = eng.empty(64)
    p = ctypes.c_void_p(d.data_ptr())
    I = ctypes.c_int64
    # null observation pointer
    rc = L.mkf_filter_loglik(I(1), I(4), 2, 1, z, I(0), p, I(0), p, I(0), p, I(0), p, I(0), z, I(0), z, I(0),
                             1, 3, p, z, z, z)
    assert rc == _lib.MKF_ERR_BAD_ARG
    # state dimension beyond the compiled kernels
    rc = L.mkf_filter_loglik(I(1), I(4), 33, 1, p, I(0), p, I(0), p, I(0), p, I(0), p, I(0), z, I(0), z, I(0),
                             1, 3, p, z, z, z)
    assert rc == _lib.MKF_ERR_UNSUPPORTED
    rc = L.mkf_dfm_loglik_alpha(I(1), I(4), 9, 1, p, I(0), p, I(0), 1.0, p, I(0), 1, p, z, z, z)
    assert rc == _lib.MKF_ERR_UNSUPPORTED
    assert b"supported" in L.mkf_error_string(rc)
    with pytest.raises(_lib.MetranB200Error):
        _lib.check(L, rc, "probe")
    # B = 0 is a no-op
    rc = L.mkf_filter_loglik(I(0), I(4), 2, 1, p, I(0), p, I(0), p, I(0), p, I(0), p, I(0), z, I(0), z, I(0),
                             1, 3, p, z, z, z)
    assert rc == _lib.MKF_OK


@pytest.mark.parametrize("n,m,T", [(1, 1, 1), (2, 3, 1), (8, 5, 2), (32, 1, 7), (17, 40, 5), (4, 9, 33)])
def test_tiny_and_ragged_shapes(eng, n, m, T):
    rng = np.random.default_rng(100 * n + m)
    B = 37                                   # not a multiple of any group size
    cases = [_ssm(rng, n, m, T, miss=0.3) for _ in range(B)]
    y, A, Q, H, R = (np.stack([c[i] for c in cases]) for i in range(5))
    want, wobs = ko.filter_loglik_batch(y, A, Q, H, R)
    for warm in (0, 1):
        mle, nobs, status = eng.filter_loglik(y, A, Q, H, R, warmup=warm)
        ref = np.array([ko.filter_loglik(y[b], A[b], Q[b], H[b], R[b], warmup=warm)[0] for b in range(B)])
        np.testing.assert_allclose(mle.cpu().numpy(), ref, rtol=0, atol=1e-9)
    out = eng.filter_history(y, A, Q, H, R)
    np.testing.assert_allclose(out["mle"].cpu().numpy(), want, rtol=0, atol=1e-9)
    assert np.array_equal(out["nobs"].cpu().numpy(), wobs)
    sm, sc, st = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], A)
    for b in (0, B - 1):
        obs, idx, cnt = ko.pack_observations(y[b])
        _, _, _, fm, fc, pm, pc = ko.seqkalmanfilter(obs, A[b], Q[b], H[b], R[b], idx, cnt, np.zeros(n), np.eye(n))
        s1, s2 = ko.kalmansmoother(fm, fc, pm, pc, A[b])
        np.testing.assert_allclose(sm[b].cpu().numpy(), s1, atol=1e-9)
        np.testing.assert_allclose(sc[b].cpu().numpy(), s2, atol=1e-9)


def test_all_observations_missing(eng):
    y = np.full((3, 20, 4), np.nan)
    alpha = np.full((3, 6), 10.0)
    G = np.full((3, 4, 2), 0.5)
    mle, nobs, status = eng.dfm_loglik_alpha(y, alpha, G)
    assert np.all(mle.cpu().numpy() == 0.0) and np.all(nobs.cpu().numpy() == 0) and int(status.abs().sum()) == 0
    mats = [ko.dfm_matrices(alpha[b], G[b]) for b in range(3)]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    out = eng.filter_history(y, A, Q, H, R)
    assert int(out["cnt"].sum()) == 0 and float(out["mle"].abs().sum()) == 0.0
    # predict-only recursion: P_t = phi^2 P_{t-1} + q
    phi2 = np.exp(-2.0 / 10.0)
    assert abs(float(out["fc"][0, 0, 0, 0]) - (phi2 + (1 - phi2) * 0.5)) < 1e-15


def test_sentinel_values_are_missing(eng, cases):
    """inf, -inf and exactly -1e10 are treated as missing (kalmanfilter.py:657, 666-667)."""
    c = cases["dfm_n3_m2_specials"]
    assert np.isinf(c["y"]).sum() == 2 and (c["y"] == -1e10).sum() == 1
    G = c["H"][:, 2:]
    mle, nobs, _ = eng.dfm_loglik(c["y"][None], np.diag(c["A"])[None], np.diag(c["Q"])[None], G[None])
    assert abs(float(mle[0]) - float(c["mle"])) < 1e-9
    assert int(nobs[0]) == int(c["ocnt"][1:].sum())


def test_shared_observations_odd_row_length(eng):
    """One data set for many draws (stride 0) with T*m odd -> non-TMA staging of a shared row."""
    rng = np.random.default_rng(3)
    T, m, k = 77, 3, 1
    y = rng.standard_normal((T, m))
    y[rng.random((T, m)) < 0.2] = np.nan
    draws = rng.uniform(2, 60, (45, m + k))
    G = rng.uniform(0.3, 0.8, (m, k))
    mle, _, _ = eng.dfm_loglik_alpha(y, draws, G)
    mats = [ko.dfm_matrices(a, G) for a in draws]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    want, _ = ko.filter_loglik_batch(y, A, Q, H, R)
    np.testing.assert_allclose(mle.cpu().numpy(), want, rtol=0, atol=1e-9)
    # same with an even row (TMA path, single-row box)
    mle2, _, _ = eng.dfm_loglik_alpha(y[:76], draws, G)
    want2, _ = ko.filter_loglik_batch(y[:76], A, Q, H, R)
    np.testing.assert_allclose(mle2.cpu().numpy(), want2, rtol=0, atol=1e-9)
