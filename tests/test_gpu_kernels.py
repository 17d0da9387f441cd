"""Parity of the CUDA kernels (through the C ABI) against the reference's golden outputs
and against the CPU oracle on seeded random batches.  Tolerances (BASELINE.json):
<= 1e-9 on -2 log L, <= 1e-8 on smoothed states; the tests below use tighter ones.
"""
import numpy as np
import pytest
import torch
from conftest import CASE_TAGS

from oracle import kalman_oracle as ko

pytestmark = pytest.mark.gpu

MLE_TOL = 1e-9
STATE_TOL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from metran_b200.batched import KalmanEngine

    return KalmanEngine()


def _np(t):
    return t.detach().cpu().numpy()


def _default_init(c):
    n = c["A"].shape[0]
    return np.array_equal(c["x0"], np.zeros(n)) and np.array_equal(c["P0"], np.eye(n))


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_generic_loglik_golden(eng, cases, tag):
    c = cases[tag]
    for w, key in ((1, "mle"), (0, "mle_w0"), (3, "mle_w3")):
        mle, nobs, status = eng.filter_loglik(c["y"][None], c["A"], c["Q"], c["H"], c["R"],
                                              x0=c["x0"], P0=c["P0"], warmup=w)
        assert abs(float(mle[0]) - float(c[key])) <= MLE_TOL, (tag, w)
        assert int(nobs[0]) == int(np.sum(c["ocnt"][w:]))
        assert int(status[0]) == 0


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_generic_history_and_smoother_golden(eng, cases, tag):
    c = cases[tag]
    out = eng.filter_history(c["y"][None], c["A"], c["Q"], c["H"], c["R"], x0=c["x0"], P0=c["P0"])
    cnt = _np(out["cnt"])[0]
    assert np.array_equal(cnt, c["ocnt"])
    seen = cnt > 0
    np.testing.assert_allclose(_np(out["sigma"])[0][seen], c["sigmas"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(_np(out["detf"])[0][seen], c["detfs"], rtol=0, atol=1e-11)
    assert np.all(_np(out["sigma"])[0][~seen] == 0)
    for key in ("fm", "fc", "pm", "pc"):
        np.testing.assert_allclose(_np(out[key])[0], c[key], rtol=0, atol=STATE_TOL, err_msg=key)
    assert abs(float(out["mle"][0]) - float(c["mle"])) <= MLE_TOL
    sm, sc, status = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], c["A"])
    assert int(status[0]) == 0
    np.testing.assert_allclose(_np(sm)[0], c["sm"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(_np(sc)[0], c["sc"], rtol=0, atol=1e-8)
    # smoother fed with the REFERENCE's filter output isolates the smoother kernel
    sm2, sc2, _ = eng.smoother(c["fm"], c["fc"], c["pm"], c["pc"], c["A"])
    np.testing.assert_allclose(_np(sm2)[0], c["sm"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(_np(sc2)[0], c["sc"], rtol=0, atol=1e-9)


def _dfm_parts(c):
    A, Q, H = c["A"], c["Q"], c["H"]
    m, n = H.shape
    if m > n or not np.array_equal(H[:, :m], np.eye(m)):
        return None
    if np.count_nonzero(A - np.diag(np.diag(A))) or np.count_nonzero(Q - np.diag(np.diag(Q))):
        return None
    return np.diag(A).copy(), np.diag(Q).copy(), H[:, m:].copy()


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_dfm_loglik_golden(eng, cases, tag):
    c = cases[tag]
    parts = _dfm_parts(c)
    if parts is None or not _default_init(c):
        pytest.skip("not Metran-shaped")
    phi, q, G = parts
    m, k = G.shape
    if not eng.dfm_supported(m, k):
        pytest.skip("outside the compiled DFM kernels")
    for w, key in ((1, "mle"), (0, "mle_w0"), (3, "mle_w3")):
        mle, nobs, status = eng.dfm_loglik(c["y"][None], phi[None], q[None], G[None], r=c["R"][None], warmup=w)
        assert abs(float(mle[0]) - float(c[key])) <= MLE_TOL, (tag, w)
        assert int(nobs[0]) == int(np.sum(c["ocnt"][w:]))
        assert int(status[0]) == 0


@pytest.mark.parametrize("name", ["b21", "nbmodel"])
def test_metran_models_alpha_path(eng, request, name):
    """Known answers of the notebook models through the fused alpha -> (phi, q) prologue."""
    g = request.getfixturevalue(name)
    y, F = g["y"], g["factors"]
    alphas = np.concatenate([g["alpha_init"][None], g["alpha_opt"][None], g["alpha_draws"]])
    want = np.concatenate([[g["mle_init"]], [g["mle_opt"]], g["mle_draws"]])
    mle, nobs, status = eng.dfm_loglik_alpha(y, alphas, F)          # shared y, shared loadings
    np.testing.assert_allclose(_np(mle), want, rtol=0, atol=MLE_TOL)
    assert int(status.abs().sum()) == 0
    # generic kernel on the same models
    mats = [ko.dfm_matrices(a, F) for a in alphas]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    mle2, _, _ = eng.filter_loglik(y, A, Q, H, R)
    np.testing.assert_allclose(_np(mle2), want, rtol=0, atol=MLE_TOL)


def test_b21_history_smoother_projection(eng, b21):
    g = b21
    sel = g["sel"]
    out = eng.filter_history(g["y"][None], g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"])
    cnt = _np(out["cnt"])[0]
    assert np.array_equal(cnt, g["observation_count"])
    np.testing.assert_allclose(_np(out["sigma"])[0][cnt > 0], g["sigmas"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(_np(out["detf"])[0][cnt > 0], g["detfs"], rtol=0, atol=1e-11)
    for key in ("fm", "fc", "pm", "pc"):
        np.testing.assert_allclose(_np(out[key])[0][sel], g[key], rtol=0, atol=STATE_TOL)
    sm, sc, status = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], g["A_opt"])
    np.testing.assert_allclose(_np(sm)[0], g["sm_full"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(_np(sc)[0][sel], g["sc"], rtol=0, atol=1e-8)
    Hs = g["H_opt"] * g["oseries_std"][:, None]
    mu, var = eng.project(sm, sc, Hs)
    np.testing.assert_allclose(_np(mu)[0][sel] + g["oseries_mean"], g["sim_means"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(_np(var)[0][sel], g["sim_variances"], rtol=0, atol=1e-8)


def _random_dfm_batch(rng, B, T, m, k, miss):
    n = m + k
    alpha = rng.uniform(5, 50, (B, n))
    phi = np.exp(-1.0 / alpha)
    G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(max(k, 1))
    comm = np.sum(G ** 2, axis=2)
    q = np.concatenate([(1 - phi[:, :m] ** 2) * (1 - comm), 1 - phi[:, m:] ** 2], axis=1)
    x = np.zeros((B, n))
    y = np.zeros((B, T, m))
    for t in range(T):
        x = phi * x + np.sqrt(q) * rng.standard_normal((B, n))
        y[:, t] = x[:, :m] + np.einsum("bmk,bk->bm", G, x[:, m:])
    y[rng.random(y.shape) < miss] = np.nan
    return y, alpha, phi, q, G


@pytest.mark.parametrize("m,k,T,miss", [(5, 3, 501, 0.0), (5, 3, 333, 0.3), (5, 1, 257, 0.5),
                                        (8, 3, 64, 0.2), (3, 0, 90, 0.1), (1, 1, 75, 0.2), (7, 2, 41, 0.0)])
def test_dfm_random_batch_vs_oracle(eng, m, k, T, miss):
    """Seeded batch (ragged last warp, odd row strides -> non-TMA staging) vs the C oracle."""
    rng = np.random.default_rng(1000 + 17 * m + k)
    B = 203
    y, alpha, phi, q, G = _random_dfm_batch(rng, B, T, m, k, miss)
    n = m + k
    A = np.zeros((B, n, n)); Q = np.zeros((B, n, n)); H = np.zeros((B, m, n))
    idx = np.arange(n)
    A[:, idx, idx] = phi
    Q[:, idx, idx] = q
    H[:, np.arange(m), np.arange(m)] = 1.0
    H[:, :, m:] = G
    R = np.zeros((B, m))
    want, wobs = ko.filter_loglik_batch(y, A, Q, H, R)
    mle, nobs, status = eng.dfm_loglik(y, phi, q, G)
    np.testing.assert_allclose(_np(mle), want, rtol=0, atol=MLE_TOL)
    assert np.array_equal(_np(nobs), wobs)
    mle_a, _, _ = eng.dfm_loglik_alpha(y, alpha, G)
    np.testing.assert_allclose(_np(mle_a), want, rtol=0, atol=MLE_TOL)
    mle_g, nobs_g, _ = eng.filter_loglik(y, A, Q, H, R)
    np.testing.assert_allclose(_np(mle_g), want, rtol=0, atol=MLE_TOL)
    assert np.array_equal(_np(nobs_g), wobs)


@pytest.mark.parametrize("m,k,T", [(5, 3, 420), (5, 3, 36), (4, 2, 301), (6, 2, 150), (2, 1, 97), (5, 1, 64),
                                   (8, 3, 80), (6, 3, 77)])
def test_dfm_collapse_transitions_vs_oracle(eng, m, k, T):
    """Runs of fully observed steps use the factor-only step (dfm_collapsed.cuh) when EVERY lane
    of the warp is dense; this batch makes warps collapse, expand and re-collapse:
    warp 0 dense throughout, warp 1 with gaps shared by all its lanes (rows fully missing or one
    series missing), warp 2 dense except single lanes with private gaps, warp 3 ragged."""
    rng = np.random.default_rng(77 + 13 * m + k + T)
    B = 32 * 3 + 9
    y, alpha, phi, q, G = _random_dfm_batch(rng, B, T, m, k, 0.0)
    rows = rng.random(T) < 0.15
    y[32:64, rows] = np.nan                                   # whole-step gaps, warp-wide
    y[32:64, (rng.random(T) < 0.1) & ~rows, 0] = np.nan       # one series missing, warp-wide
    y[40, T // 2, m - 1] = -1e10                              # the reference's drop sentinel
    for lane in (64, 70, 95):
        y[lane, rng.random(T) < 0.2, rng.integers(0, m)] = np.nan
    y[96:, : T // 3] = np.nan                                 # late start in the ragged warp
    y[100, -1] = np.nan
    n = m + k
    A = np.zeros((B, n, n)); Q = np.zeros((B, n, n)); H = np.zeros((B, m, n))
    idx = np.arange(n)
    A[:, idx, idx] = phi
    Q[:, idx, idx] = q
    H[:, np.arange(m), np.arange(m)] = 1.0
    H[:, :, m:] = G
    R = np.zeros((B, m))
    for warmup in (1, 4):
        want, wobs = ko.filter_loglik_batch(y, A, Q, H, R, warmup=warmup)
        mle, nobs, status = eng.dfm_loglik(y, phi, q, G, warmup=warmup)
        np.testing.assert_allclose(_np(mle), want, rtol=0, atol=MLE_TOL)
        assert np.array_equal(_np(nobs), wobs)
        assert not _np(status).any()
        mle_a, _, _ = eng.dfm_loglik_alpha(y, alpha, G, warmup=warmup)
        np.testing.assert_allclose(_np(mle_a), want, rtol=0, atol=MLE_TOL)
    # R != 0 must take the ordinary path (H P != 0 after an update) and still agree
    Rn = np.full((B, m), 0.05)
    want, _ = ko.filter_loglik_batch(y, A, Q, H, Rn)
    mle_r, _, _ = eng.dfm_loglik(y, phi, q, G, r=Rn)
    np.testing.assert_allclose(_np(mle_r), want, rtol=0, atol=MLE_TOL)


def test_reduce_sum_deterministic(eng):
    v = torch.randn(100003, dtype=torch.float64, device=eng.device) * 1e3
    a = float(eng.reduce_sum(v)[0])
    b = float(eng.reduce_sum(v)[0])
    assert a == b
    import math

    assert abs(a - math.fsum(_np(v).tolist())) < 1e-7


def test_nonpositive_f_is_flagged(eng):
    """R = 0 with a duplicated series makes f exactly 0 on the second update."""
    y = np.ones((1, 6, 2))
    A = np.eye(1) * 0.9; Q = np.eye(1) * 0.19; H = np.ones((2, 1)); R = np.zeros(2)
    mle, nobs, status = eng.filter_loglik(y, A, Q, H, R)
    assert int(status[0]) & 1
    assert not np.isfinite(float(mle[0]))


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_dfm_history_golden(eng, cases, tag):
    """Thread-per-model history kernel (mkf_dfm_filter_history) against the reference's arrays."""
    c = cases[tag]
    parts = _dfm_parts(c)
    if parts is None or not _default_init(c):
        pytest.skip("not Metran-shaped")
    phi, q, G = parts
    m, k = G.shape
    if not eng.dfm_supported(m, k):
        pytest.skip("outside the compiled DFM kernels")
    out = eng.dfm_filter_history(c["y"][None], phi[None], q[None], G[None], r=c["R"][None])
    cnt = _np(out["cnt"])[0]
    assert np.array_equal(cnt, c["ocnt"])
    seen = cnt > 0
    np.testing.assert_allclose(_np(out["sigma"])[0][seen], c["sigmas"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(_np(out["detf"])[0][seen], c["detfs"], rtol=0, atol=1e-11)
    for key in ("fm", "fc", "pm", "pc"):
        np.testing.assert_allclose(_np(out[key])[0], c[key], rtol=0, atol=STATE_TOL, err_msg=key)
    assert abs(float(out["mle"][0]) - float(c["mle"])) <= MLE_TOL
    sm, sc, _ = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], c["A"])
    np.testing.assert_allclose(_np(sm)[0], c["sm"], rtol=0, atol=1e-8)


def test_dfm_history_batch_matches_generic_kernel(eng):
    """Ragged batch (B = 203): thread-per-model history == warp-group history to rounding."""
    rng = np.random.default_rng(21)
    B, T, m, k = 203, 97, 5, 3
    y, alpha, phi, q, G = _random_dfm_batch(rng, B, T, m, k, 0.25)
    n = m + k
    A = np.zeros((B, n, n)); Q = np.zeros((B, n, n)); H = np.zeros((B, m, n))
    idx = np.arange(n)
    A[:, idx, idx] = phi; Q[:, idx, idx] = q
    H[:, np.arange(m), np.arange(m)] = 1.0; H[:, :, m:] = G
    ref = eng.filter_history(y, A, Q, H, np.zeros((B, m)), flags=3)
    out = eng.dfm_filter_history(y, phi, q, G)
    assert torch.equal(out["cnt"], ref["cnt"]) and torch.equal(out["nobs"], ref["nobs"])
    for key in ("sigma", "detf", "fm", "fc", "pm", "pc", "mle"):
        np.testing.assert_allclose(_np(out[key]), _np(ref[key]), rtol=0, atol=1e-9, err_msg=key)
    # and against the oracle for a few models
    for b in (0, 101, 202):
        obs, oi, oc = ko.pack_observations(y[b])
        _, _, _, fm, fc, pm, pc = ko.seqkalmanfilter(obs, A[b], Q[b], H[b], np.zeros(m), oi, oc, np.zeros(n), np.eye(n))
        np.testing.assert_allclose(_np(out["fc"])[b], fc, rtol=0, atol=1e-9)
        np.testing.assert_allclose(_np(out["pm"])[b], pm, rtol=0, atol=1e-9)


@pytest.mark.parametrize("tag", ["dfm_n8_m5_miss30", "dfm_n3_m2_miss50", "dense_n16_m7", "dense_n32_m64", "dense_n20_m33"])
def test_smoother_with_fused_projections(eng, cases, tag):
    """mkf_smoother_project: the projections written from the smoother's epilogue equal
    mkf_project applied to the materialised covariances, and the smoothed covariances may be skipped."""
    c = cases[tag]
    out = eng.filter_history(c["y"][None], c["A"], c["Q"], c["H"], c["R"], x0=c["x0"][None], P0=c["P0"][None])
    sm, sc, _ = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], c["A"])
    Hs = c["H"] * np.linspace(0.5, 2.0, c["H"].shape[0])[:, None]
    mu, var = eng.project(sm, sc, Hs)
    res = eng.smoother_project(out["fm"], out["fc"], out["pm"], out["pc"], c["A"], Hs)
    assert torch.equal(res["sm"], sm) and torch.equal(res["sc"], sc)
    np.testing.assert_allclose(_np(res["proj_mean"]), _np(mu), rtol=0, atol=1e-11)
    np.testing.assert_allclose(_np(res["proj_var"]), _np(var), rtol=0, atol=1e-11)
    np.testing.assert_allclose(_np(res["state_var"])[0], np.diagonal(c["sc"], axis1=1, axis2=2), rtol=0, atol=1e-8)
    lean = eng.smoother_project(out["fm"], out["fc"], out["pm"], out["pc"], c["A"], Hs, covariances=False,
                                state_variances=False)
    assert lean["sc"] is None and torch.equal(lean["proj_var"], res["proj_var"]) and torch.equal(lean["sm"], sm)


@pytest.mark.parametrize("n_extra", [0, 14])
def test_smoother_with_a_singular_predicted_covariance(eng, n_extra):
    """A state without process noise (communality 1: q = 0) that starts exactly known (P0_ii = 0) keeps a
    zero row and column in every covariance: P_p is singular.  The reference's pinv (kalmanfilter.py:455)
    gives that direction zero gain; the Cholesky-based smoother must do the same (finite results equal to
    the pinv ones, status flagged) instead of NaN.  n_extra pads the state to exercise the wide kernel."""
    rng = np.random.default_rng(3)
    m, k, T = 4, 2, 150
    G = rng.uniform(0.3, 0.6, (m, k))
    G[1] /= np.linalg.norm(G[1])                    # communality exactly 1 -> q = 0 for series 1
    n = m + k + n_extra
    alpha = rng.uniform(5, 40, n)
    phi = np.exp(-1.0 / alpha)
    comm = np.concatenate([(G ** 2).sum(1), np.zeros(k + n_extra)])
    q = (1 - phi ** 2) * np.where(np.arange(n) < m, 1 - comm[:n], 1.0)
    q[1] = 0.0
    H = np.zeros((m, n)); H[:, :m] = np.eye(m); H[:, m:m + k] = G
    A, Q, R = np.diag(phi), np.diag(q), np.zeros(m)
    y = rng.standard_normal((T, m))
    y[rng.random(y.shape) < 0.2] = np.nan
    P0 = np.eye(n)
    P0[1, 1] = 0.0
    out = eng.filter_history(y[None], A, Q, H, R, x0=np.zeros((1, n)), P0=P0[None])
    sm, sc, st = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], A)
    assert torch.isfinite(sm).all() and torch.isfinite(sc).all()
    assert int(st[0]) & 2                             # MKF_STATUS_CHOL_FAILED: flagged, not fatal
    obs, idx, cnt = ko.pack_observations(y)
    _, _, _, fm, fc, pm, pc = ko.seqkalmanfilter(obs, A, Q, H, R, idx, cnt, np.zeros(n), P0)
    assert min(np.linalg.eigvalsh(pc[t]).min() for t in range(T)) < 1e-15          # really singular
    s1, s2 = ko.kalmansmoother(fm, fc, pm, pc, A)   # pinv-based restatement of the reference
    np.testing.assert_allclose(_np(sm)[0], s1, rtol=0, atol=1e-8)
    np.testing.assert_allclose(_np(sc)[0], s2, rtol=0, atol=1e-8)
