"""Pin oracle/ (the CPU restatement) against golden outputs of the live reference.

Golden vectors: tests/golden/*.npz, written by tests/golden/make_golden.py from
/root/reference/metran (kalmanfilter.py:236-476, 550-567, 646-674; metran.py:246-416).
"""
import numpy as np
import pytest
from conftest import CASE_TAGS

from oracle import kalman_oracle as ko


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_pack_observations_matches_reference(cases, tag):
    c = cases[tag]
    obs, idx, cnt = ko.pack_observations(c["y"])
    assert np.array_equal(obs, c["obs"])
    assert np.array_equal(idx, c["oidx"])
    assert np.array_equal(cnt, c["ocnt"])


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_filter_bit_exact_against_reference(cases, tag):
    """Same IEEE operations in the same order as the numba kernel -> identical bits."""
    c = cases[tag]
    sig, det, ns, fm, fc, pm, pc = ko.seqkalmanfilter(
        c["obs"], c["A"], c["Q"], c["H"], c["R"], c["oidx"], c["ocnt"], c["x0"], c["P0"])
    assert ns == c["sigmas"].shape[0]
    # log() of the C library vs numba's: allow 1 ulp on detf, everything else bitwise
    assert np.array_equal(sig[:ns], c["sigmas"])
    np.testing.assert_allclose(det[:ns], c["detfs"], rtol=0, atol=4e-15)
    for got, key in ((fm, "fm"), (fc, "fc"), (pm, "pm"), (pc, "pc")):
        assert np.array_equal(got, c[key]), key
    for w, key in ((1, "mle"), (0, "mle_w0"), (3, "mle_w3")):
        assert abs(ko.get_mle(sig[:ns], det[:ns], c["ocnt"], w) - float(c[key])) < 1e-11


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_loglik_only_path(cases, tag):
    c = cases[tag]
    if not (np.array_equal(c["x0"], np.zeros_like(c["x0"])) and np.array_equal(c["P0"], np.eye(c["P0"].shape[0]))):
        pytest.skip("default initial state only")
    mle, nobs = ko.filter_loglik(c["y"], c["A"], c["Q"], c["H"], c["R"])
    assert abs(mle - float(c["mle"])) < 1e-10
    got, nb = ko.filter_loglik_batch(c["y"][None], c["A"][None], c["Q"][None], c["H"][None], c["R"][None])
    assert abs(got[0] - float(c["mle"])) < 1e-9 and nb[0] == nobs


@pytest.mark.parametrize("tag", CASE_TAGS)
def test_smoother_against_reference(cases, tag):
    c = cases[tag]
    sm, sc = ko.kalmansmoother(c["fm"], c["fc"], c["pm"], c["pc"], c["A"])
    np.testing.assert_allclose(sm, c["sm"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(sc, c["sc"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("name", ["b21", "nbmodel"])
def test_metran_models(request, name):
    """Known answers of the two notebook models."""
    g = request.getfixturevalue(name)
    y, F = g["y"], g["factors"]
    for alpha, key in ((g["alpha_init"], "mle_init"), (g["alpha_opt"], "mle_opt")):
        A, Q, H, R = ko.dfm_matrices(alpha, F)
        mle, _ = ko.filter_loglik(y, A, Q, H, R)
        assert abs(mle - float(g[key])) < 1e-9
    A, Q, H, R = ko.dfm_matrices(g["alpha_opt"], F)
    for got, key in zip((A, Q, H, R), ("A_opt", "Q_opt", "H_opt", "R_opt")):
        assert np.array_equal(got, g[key]), key
    for a, want in zip(g["alpha_draws"], g["mle_draws"]):
        assert abs(ko.filter_loglik(y, *ko.dfm_matrices(a, F))[0] - want) < 1e-9 * max(1.0, abs(want))


def test_published_numbers(b21, nbmodel):
    """Saved notebook outputs (examples/*.ipynb), SURVEY.md section 6."""
    assert round(float(b21["mle_opt"]), 2) == 2332.33
    assert abs(float(b21["mle_init"]) - 2384.792799342231) < 1e-9
    assert abs(float(nbmodel["mle_init"]) - 2431.3389452203646) < 1e-9
    np.testing.assert_allclose(nbmodel["factors"].ravel(), [0.93540765, 0.93540765], atol=1e-8)
    np.testing.assert_allclose(b21["factors"].ravel(),
                               [0.857982, 0.935874, 0.966197, 0.957794, 0.900857], atol=1e-6)


def test_metran_model_history_and_projections(b21):
    g = b21
    y = g["y"]
    A, Q, H, R = g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"]
    obs, idx, cnt = ko.pack_observations(y)
    n = A.shape[0]
    sig, det, ns, fm, fc, pm, pc = ko.seqkalmanfilter(obs, A, Q, H, R, idx, cnt, np.zeros(n), np.eye(n))
    sel = g["sel"]
    assert np.array_equal(cnt, g["observation_count"])
    assert np.array_equal(sig[:ns], g["sigmas"])
    np.testing.assert_allclose(det[:ns], g["detfs"], rtol=0, atol=4e-15)
    assert np.array_equal(fm[sel], g["fm"]) and np.array_equal(fc[sel], g["fc"])
    assert np.array_equal(pm[sel], g["pm"]) and np.array_equal(pc[sel], g["pc"])
    sm, sc = ko.kalmansmoother(fm, fc, pm, pc, A)
    np.testing.assert_allclose(sm[sel], g["sm"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(sc[sel], g["sc"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(sm, g["sm_full"], rtol=0, atol=1e-12)
    # projections: metran.py:758-829 scale H by the series std and add the mean
    Hs = H * g["oseries_std"][:, None]
    means, var = ko.simulate(sm, sc, Hs)
    np.testing.assert_allclose(means[sel] + g["oseries_mean"], g["sim_means"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(var[sel], g["sim_variances"], rtol=0, atol=1e-10)
    sdf, cdf = ko.decompose(sm, Hs, nsdf=y.shape[1])
    np.testing.assert_allclose(sdf[sel, 0] + g["oseries_mean"][0], g["decompose_first"][:, 0], atol=1e-10)
    np.testing.assert_allclose(cdf[0][sel, 0], g["decompose_first"][:, 1], atol=1e-10)
