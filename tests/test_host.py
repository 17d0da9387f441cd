"""CPU-only tests: the C ABI surface, the host-side packing logic and the mirrors of the
reference's Kalman object / objective, driven through tests/fakelib.py (oracle-backed)."""
import ctypes
import os
import re

import numpy as np
import pytest
from conftest import CASE_TAGS, ROOT

from fakelib import FakeLib
from oracle import kalman_oracle as ko


@pytest.fixture()
def eng():
    from metran_b200.batched import KalmanEngine

    return KalmanEngine(lib=FakeLib(), device="cpu")


# ------------------------------------------------------------------------------- ABI
def test_library_exports_every_declared_symbol():
    from metran_b200 import _lib, build

    header = open(os.path.join(ROOT, "include", "metran_b200.h")).read()
    declared = set(re.findall(r"\b(mkf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    path = build.build()
    L = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(L, name), "libmetran_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.load().mkf_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch

    from metran_b200 import _lib
    from metran_b200.batched import KalmanEngine

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MetranB200Error):
        KalmanEngine()


def test_product_does_not_import_oracle():
    """The product path must never route through the checker (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "metran_b200")
    pat = re.compile(r"^\s*(from|import)\s+(oracle|tests|fakelib)\b", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), "%s imports test infrastructure" % f
                assert "kalman_oracle" not in src, f


# --------------------------------------------------------------------------- packing
@pytest.mark.parametrize("tag", CASE_TAGS)
def test_pack_unpack_observations(cases, tag):
    from metran_b200.kalmanfilter import pack_observations, unpack_observations

    c = cases[tag]
    obs, idx, cnt, observed = pack_observations(c["y"])
    assert np.array_equal(obs, c["obs"]) and np.array_equal(idx, c["oidx"]) and np.array_equal(cnt, c["ocnt"])
    y = unpack_observations(obs, idx, cnt)
    assert np.array_equal(np.isnan(y), ~observed)
    assert np.array_equal(y[observed], c["y"][observed])


def test_unpack_rejects_descending_indices():
    from metran_b200.kalmanfilter import unpack_observations

    obs = np.ones((2, 3))
    idx = np.array([[2.0, 0.0, 0.0], [0.0, 1.0, 2.0]])
    with pytest.raises(ValueError):
        unpack_observations(obs, idx, np.array([2, 3]))


# ------------------------------------------------------------------- Kalman object
def test_filtermethod_contract(eng, cases):
    from metran_b200.kalmanfilter import kalmansmoother_cuda, seqkalmanfilter_cuda

    c = cases["dfm_n8_m5_miss30"]
    res = seqkalmanfilter_cuda(c["obs"], c["A"], c["Q"], c["H"], c["R"], c["oidx"], c["ocnt"],
                               c["x0"], c["P0"], engine=eng)
    sig, det, ns, fm, fc, pm, pc = res
    assert ns == c["sigmas"].shape[0] and sig.shape[0] >= ns
    np.testing.assert_allclose(sig[:ns], c["sigmas"], atol=1e-12)
    np.testing.assert_allclose(det[:ns], c["detfs"], atol=1e-12)
    for got, key in ((fm, "fm"), (fc, "fc"), (pm, "pm"), (pc, "pc")):
        np.testing.assert_allclose(got, c[key], atol=1e-12)
    sm, sc = kalmansmoother_cuda(fm, fc, pm, pc, c["A"], engine=eng)
    np.testing.assert_allclose(sm, c["sm"], atol=1e-12)
    np.testing.assert_allclose(sc, c["sc"], atol=1e-12)


def test_spkalmanfilter_mirror_on_b21(eng, b21):
    from metran_b200.kalmanfilter import SPKalmanFilter

    g = b21
    kf = SPKalmanFilter(device_engine=eng)
    kf.set_observations(g["y"])
    assert np.array_equal(kf.observation_count, g["observation_count"])
    kf.set_matrices(g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"])
    assert kf.nstate == 6
    kf.run_filter()
    np.testing.assert_allclose(kf.sigmas, g["sigmas"], atol=1e-12)
    np.testing.assert_allclose(kf.detfs, g["detfs"], atol=1e-12)
    assert abs(kf.get_mle() - float(g["mle_opt"])) < 1e-9
    kf.run_smoother()
    sel = g["sel"]
    np.testing.assert_allclose(kf.smoothed_state_means, g["sm_full"], atol=1e-11)
    np.testing.assert_allclose(kf.smoothed_state_covariances[sel], g["sc"], atol=1e-11)
    Hs = g["H_opt"] * g["oseries_std"][:, None]
    means, var = kf.simulate(Hs)
    assert isinstance(means, list) and len(means) == g["y"].shape[0]
    np.testing.assert_allclose(np.asarray(means)[sel] + g["oseries_mean"], g["sim_means"], atol=1e-10)
    np.testing.assert_allclose(np.asarray(var)[sel], g["sim_variances"], atol=1e-10)
    means_f, _ = kf.simulate(Hs, method="filter")
    np.testing.assert_allclose(np.asarray(means_f)[sel] + g["oseries_mean"], g["sim_means_filter"], atol=1e-10)
    sdf, cdf = kf.decompose(Hs)
    np.testing.assert_allclose(np.asarray(sdf)[sel, 0] + g["oseries_mean"][0], g["decompose_first"][:, 0], atol=1e-10)
    np.testing.assert_allclose(np.asarray(cdf[0])[sel, 0], g["decompose_first"][:, 1], atol=1e-10)


def test_engine_error_behaviour(eng, b21):
    from metran_b200.kalmanfilter import SPKalmanFilter

    kf = SPKalmanFilter(engine="numba", device_engine=eng)   # accepted, runs the CUDA path
    kf.set_observations(b21["y"][:50])
    kf.set_matrices(b21["A_opt"], b21["Q_opt"], b21["H_opt"], b21["R_opt"])
    with pytest.raises(Exception, match="Unknown engine defined in run_filter."):
        kf.run_filter(engine="fortran")
    kf.run_filter(engine="cuda")
    assert kf.filtered_state_means.shape == (50, 6)


def test_run_smoother_with_plugin_filtermethod(eng, b21):
    """A plug-in filtermethod (the reference's own extension point, kalmanfilter.py:501-504,
    761-771) leaves host arrays only; run_smoother must smooth those like the reference does
    (:685-691).  Both the advertised seqkalmanfilter_cuda and a plain Python callable."""
    from functools import partial

    from metran_b200.kalmanfilter import SPKalmanFilter, seqkalmanfilter_cuda

    g = b21
    for method in (partial(seqkalmanfilter_cuda, engine=eng), ko.seqkalmanfilter):
        kf = SPKalmanFilter(device_engine=eng)
        kf.set_observations(g["y"])
        kf.set_matrices(g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"])
        kf.filtermethod = method
        kf.run_filter()
        np.testing.assert_allclose(kf.sigmas, g["sigmas"], atol=1e-12)
        assert abs(kf.get_mle() - float(g["mle_opt"])) < 1e-9
        kf.run_smoother()
        np.testing.assert_allclose(kf.smoothed_state_means, g["sm_full"], atol=1e-11)
        np.testing.assert_allclose(kf.smoothed_state_covariances[g["sel"]], g["sc"], atol=1e-11)
        means, _ = kf.simulate(g["H_opt"] * g["oseries_std"][:, None])
        np.testing.assert_allclose(np.asarray(means)[g["sel"]] + g["oseries_mean"], g["sim_means"], atol=1e-10)


def test_lazy_history_is_a_snapshot(eng, b21):
    """run_filter() defers the history kernel; what is read later must belong to the matrices and
    observations of THAT call, whatever set_matrices()/init_states() did in between, and
    set_observations() must drop the cached likelihood of the previous data."""
    from metran_b200.kalmanfilter import SPKalmanFilter

    g = b21
    kf = SPKalmanFilter(device_engine=eng)
    kf.set_observations(g["y"])
    kf.set_matrices(g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"])
    kf.run_filter()
    mle = kf.get_mle()
    kf.set_matrices(0.5 * g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"])     # no run_filter()
    np.testing.assert_allclose(kf.sigmas, g["sigmas"], atol=1e-12)
    np.testing.assert_allclose(kf.filtered_state_means[g["sel"]], g["fm"], atol=1e-11)
    assert kf.get_mle() == mle
    # init_states(): state arrays gone, sigmas / detfs / get_mle() survive (kalmanfilter.py:506-518)
    kf.set_matrices(g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"])
    kf.run_filter()
    kf.init_states()
    assert kf.filtered_state_means is None and kf.smoothed_state_covariances is None
    assert kf.get_mle() == mle
    np.testing.assert_allclose(kf.detfs, g["detfs"], atol=1e-12)
    assert abs(kf.get_mle(warmup=3) - (np.sum(g["observation_count"][3:]) * np.log(2 * np.pi)
                                       + g["detfs"][3:].sum() + g["sigmas"][3:].sum())) < 1e-9
    assert kf.filtered_state_means is None
    # new observations: nothing of the old run may be served
    kf.run_filter()
    kf.set_observations(g["y"][:100])
    assert kf._mle_w1 is None and kf._pending is None
    kf.run_filter()
    assert kf.filtered_state_means.shape == (100, 6)


# ---------------------------------------------------------------- Metran-facing mirror
def test_dfm_objective_and_batch(eng, b21):
    from metran_b200.dfm import DynamicFactorModel

    g = b21
    mdl = DynamicFactorModel(g["y"], g["factors"], engine=eng)
    assert abs(mdl.get_mle(g["alpha_init"]) - 2384.792799342231) < 1e-9
    assert abs(mdl.objfunction(g["alpha_opt"]) - float(g["mle_opt"])) < 1e-9
    got = mdl.get_mle_batch(g["alpha_draws"])
    np.testing.assert_allclose(got, g["mle_draws"], rtol=1e-12, atol=1e-9)
    A, Q, H, R = mdl._get_matrices(g["alpha_opt"])
    for a, key in ((A, "A_opt"), (Q, "Q_opt"), (H, "H_opt"), (R, "R_opt")):
        assert np.array_equal(a, g[key])
    if hasattr(eng.lib, "calls"):
        # 94.5 % of B21B0214's daily steps are empty: the batch walks the observed steps only
        assert eng.lib.calls[-1][:3] == ("mkf_dfm_loglik_compact", 16, int(g["observation_count"].astype(bool).sum()))


def test_dfm_solve_reaches_reference_optimum(eng, nbmodel):
    """L-BFGS-B with all finite-difference points of a gradient in one batched call."""
    from metran_b200.dfm import DynamicFactorModel

    g = nbmodel
    mdl = DynamicFactorModel(g["y"], g["factors"], engine=eng)
    ok, x, stderr = mdl.solve()
    assert ok
    assert abs(mdl.parameters["obj_func"] - float(g["obj_opt"])) < 1e-5
    np.testing.assert_allclose(x, g["alpha_opt"], rtol=2e-3)
    assert mdl.parameters["batched_launches"] >= 1
    # every batched call carried nparam finite-difference points
    if hasattr(eng.lib, "calls"):
        fd = [c for c in eng.lib.calls if c[0] == "mkf_dfm_loglik_alpha" and c[1] == mdl.nstate]
        assert len(fd) == mdl.parameters["batched_launches"]


def test_dfm_mask_unmask(eng, b21):
    from metran_b200.dfm import DynamicFactorModel

    g = b21
    mdl = DynamicFactorModel(g["y"], g["factors"], engine=eng)
    Hs = g["H_opt"] * g["oseries_std"][:, None]
    # quirk Q7: the reference's masked getter reuses the matrices of the LAST objective call
    p = g["alpha_opt"]
    mask = np.zeros_like(g["y"], dtype=bool)
    row, col = int(g["mask_row"]), int(g["mask_col"])
    mask[row, col] = True
    before = mdl.get_simulated_means(p, Hs)[row, col]
    mdl.mask_observations(mask)
    after = mdl.get_simulated_means(p, Hs)
    var = mdl.get_simulated_variances(p, Hs)
    mdl.unmask_observations()
    assert before != after[row, col]
    want = g["masked_sim_row"]            # [mean, lower, upper] of the 95 % interval
    assert abs(after[row, col] + g["oseries_mean"][col] - want[0]) < 1e-6
    half = 1.959963984540054 * np.sqrt(var[row, col])
    assert abs((want[2] - want[1]) / 2 - half) < 1e-6


# ------------------------------------------------------------------------ sharding
def test_shard_bounds_cover_batch():
    from metran_b200.distributed import shard_bounds

    for total in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_dfm_filter_history_argument_packing(eng, cases):
    """mkf_dfm_filter_history: compact (phi, q, gamma) layout -> reference result contract."""
    c = cases["dfm_n8_m5_miss30"]
    m = c["H"].shape[0]
    out = eng.dfm_filter_history(c["y"][None], np.diag(c["A"])[None], np.diag(c["Q"])[None], c["H"][:, m:][None])
    cnt = out["cnt"][0].numpy()
    assert np.array_equal(cnt, c["ocnt"])
    np.testing.assert_allclose(out["sigma"][0].numpy()[cnt > 0], c["sigmas"], atol=1e-12)
    for key in ("fm", "fc", "pm", "pc"):
        np.testing.assert_allclose(out[key][0].numpy(), c[key], atol=1e-12)
    assert abs(float(out["mle"][0]) - float(c["mle"])) < 1e-10


def test_batched_fit_reaches_the_reference_optimum(eng, nbmodel):
    """fit_dfm_batch: several starts of the notebook model all reach the reference's optimum
    (examples/dynamic_factor_model.ipynb data; -2logL at the optimum from the live reference)."""
    from metran_b200.fit import fit_dfm_batch

    g = nbmodel
    y = g["y"][:600]
    B = 3
    starts = np.array([[10.0, 10.0, 10.0], [4.0, 30.0, 15.0], [25.0, 6.0, 40.0]])
    res = fit_dfm_batch(eng, np.broadcast_to(y, (B,) + y.shape).copy(),
                        np.broadcast_to(g["factors"], (B,) + g["factors"].shape).copy(), alpha0=starts)
    assert res.converged.all() and set(res.reason) <= {"gtol", "ftol"}
    # a model whose objective is not finite at the start is reported as such, not as converged
    ybad = np.broadcast_to(y, (B,) + y.shape).copy()
    fbad = np.broadcast_to(g["factors"], (B,) + g["factors"].shape).copy()
    fbad[1] = np.nan
    res_bad = fit_dfm_batch(eng, ybad, fbad, alpha0=starts, maxiter=3)
    assert not res_bad.converged[1] and res_bad.reason[1] == "nonfinite_start"
    # reference optimum on the same truncated data through the single-model path
    from metran_b200.dfm import DynamicFactorModel

    mdl = DynamicFactorModel(y, g["factors"], engine=eng)
    mdl.solve()
    assert np.all(np.abs(res.obj - mdl.parameters["obj_func"]) < 1e-4)
    np.testing.assert_allclose(res.alpha, np.tile(mdl.parameters["optimal"], (B, 1)), rtol=5e-3)


@pytest.mark.parametrize("m,k", [(3, 2), (5, 3)])
def test_steady_state_cases_are_sound(eng, m, k):
    """The body of tests/test_gpu_kernels.py::test_dfm_steady_state_vs_oracle, run through the
    oracle-backed stand-in: proves shapes, packing and tolerances of the check itself on the CPU."""
    import steady_state_cases as ssc

    ssc.check(eng, m, k, lambda t: t.cpu().numpy())


# ---------------------------------------------------------- batched FD Hessian (solver.py:65-140)
def _reference_style_covariance(f, x0, epsilon=None, diff="forward"):
    """The loop of BaseSolver._get_covariance (solver.py:96-140) restated with one objective call at
    a time -- what the batched version must reproduce."""
    from scipy.optimize import approx_fprime

    from metran_b200.dfm import near_psd

    n = x0.shape[0]
    if epsilon is None:
        epsilon = np.finfo(float).eps ** (1.0 / 4)
    for _ in range(8):
        hessian = np.zeros((n, n))
        xx = x0.copy()
        if diff == "forward":
            f0 = approx_fprime(x0, f, epsilon)
        for j in range(n):
            xx0 = xx[j]
            d = epsilon * max(np.abs(xx0), 0.1)
            xx[j] = xx0 + d
            ff = approx_fprime(xx, f, epsilon)
            if diff == "central":
                xx[j] = xx0 - d
                fb = approx_fprime(xx, f, epsilon)
                hessian[: j + 1, j] = (ff[: j + 1] - fb[: j + 1]) / (2 * d)
            else:
                hessian[: j + 1, j] = (ff[: j + 1] - f0[: j + 1]) / d
            hessian[j, : j + 1] = hessian[: j + 1, j]
            xx[j] = xx0
        cov = np.linalg.pinv(hessian)
        if np.amin(np.diag(cov)) <= 0:
            cov = np.linalg.pinv(near_psd(hessian))
        if np.amin(np.diag(cov)) > 0:
            return cov
        epsilon *= 10.0
    return cov


@pytest.mark.parametrize("diff", ["forward", "central"])
def test_batched_covariance_matches_the_reference_loop(eng, nbmodel, diff):
    """All (n+1)^2 objective values of a Hessian from ONE batched launch, identical to the
    one-call-at-a-time loop of the reference."""
    from metran_b200.dfm import DynamicFactorModel

    g = nbmodel
    dfm = DynamicFactorModel(g["y"], g["factors"], engine=eng)
    x0 = g["alpha_opt"].astype(float)
    calls = []
    orig = dfm.get_mle_batch

    def counting(P):
        calls.append(np.atleast_2d(P).shape[0])
        return orig(P)

    dfm.get_mle_batch = counting
    cov = dfm.get_covariance(x0, diff=diff)
    n = x0.size
    assert len(calls) == 1 and calls[0] <= ((n + 1) ** 2 if diff == "forward" else (2 * n + 1) * (n + 1))
    want = _reference_style_covariance(lambda p: float(orig(np.asarray(p)[None])[0]), x0, diff=diff)
    np.testing.assert_allclose(cov, want, rtol=1e-9, atol=1e-12)
    assert np.all(np.diag(cov) > 0)
    # the reference's own stderr at this optimum came from L-BFGS-B's hess_inv; the FD Hessian must be
    # of the same order of magnitude (loose: different estimators of the same curvature)
    ratio = np.sqrt(np.diag(cov)) / g["stderr"]
    assert np.all((ratio > 0.05) & (ratio < 20.0)), ratio


# ------------------------------------------------------------------ compacted time grid (f4)
def _sparse_batch(seed=0, B=7, T=300, m=4, k=2):
    rng = np.random.default_rng(seed)
    n = m + k
    alpha = rng.uniform(3.0, 60.0, (B, n))
    G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(k)
    y = rng.standard_normal((B, T, m))
    keep = rng.random((B, T)) < rng.uniform(0.03, 0.3, (B, 1))      # ragged: different counts per model
    y[~keep] = np.nan
    y[rng.random(y.shape) < 0.2] = np.nan                              # partially observed days
    y[1] = np.nan                                                      # a model without any observation
    y[2, 0] = np.nan                                                   # first step empty
    y[3, 17, 0] = -1e10                                                # the reference's sentinel = missing
    y[3, 18, 1] = np.inf
    return y, alpha, G


def test_compact_observations_properties(eng):
    y, alpha, G = _sparse_batch()
    yc, tidx, J, frac = eng.compact_observations(y)
    yc, tidx = yc.numpy(), tidx.numpy()
    obs = np.isfinite(y) & (y != -1e10)
    anyobs = obs.any(-1)
    assert J == anyobs.sum(1).max() and yc.shape == (y.shape[0], J, y.shape[2]) and tidx.dtype == np.int32
    for b in range(y.shape[0]):
        steps = np.flatnonzero(anyobs[b])
        c = steps.size
        assert np.array_equal(tidx[b, :c], steps)
        assert np.array_equal(np.isnan(yc[b, :c]), np.isnan(y[b, steps]))
        assert np.array_equal(np.nan_to_num(yc[b, :c], nan=0.0, posinf=1.0, neginf=-1.0),
                              np.nan_to_num(y[b, steps], nan=0.0, posinf=1.0, neginf=-1.0))
        assert np.all(np.isnan(yc[b, c:])) and np.all(tidx[b, c:] == (steps[-1] if c else -1))
    # shared block
    yc1, t1, J1, _ = eng.compact_observations(y[0])
    assert yc1.shape == (J1, y.shape[2]) and t1.shape == (J1,) and J1 == anyobs[0].sum()


@pytest.mark.parametrize("warmup", [0, 1, 3])
def test_compact_path_equals_full_grid(eng, warmup):
    """Host side of the compacted grid: indices, padding, warmup / nobs on the ORIGINAL time axis."""
    y, alpha, G = _sparse_batch(seed=1)
    yc, tidx, J, _ = eng.compact_observations(y)
    full = eng.dfm_loglik_alpha(y, alpha, G, warmup=warmup)
    comp = eng.dfm_loglik_compact(yc, tidx, G, alpha=alpha, warmup=warmup)
    keep = np.ones(y.shape[0], dtype=bool)
    keep[1] = False                       # all-missing model: the oracle's mle is 0 - both agree on nobs
    np.testing.assert_allclose(comp[0].numpy()[keep], full[0].numpy()[keep], rtol=0, atol=1e-9)
    assert np.array_equal(comp[1].numpy(), full[1].numpy())
    phi = np.exp(-1.0 / alpha)
    q = np.concatenate([(1 - phi[:, :4] ** 2) * (1 - (G ** 2).sum(2)), 1 - phi[:, 4:] ** 2], axis=1)
    comp2 = eng.dfm_loglik_compact(yc, tidx, G, phi=phi, q=q, warmup=warmup)
    np.testing.assert_allclose(comp2[0].numpy()[keep], full[0].numpy()[keep], rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        eng.dfm_loglik_compact(yc, tidx, G, alpha=alpha, phi=phi, q=q)


def test_spkalmanfilter_uses_compact_grid_for_sparse_data(eng, b21):
    from metran_b200.kalmanfilter import SPKalmanFilter

    g = b21
    kf = SPKalmanFilter(device_engine=eng)
    kf.set_observations(g["y"])
    assert kf._compact is not None and kf._compact[0].shape[0] == int((g["observation_count"] > 0).sum())
    kf.set_matrices(g["A_opt"], g["Q_opt"], g["H_opt"], g["R_opt"])
    kf.run_filter()
    assert eng.lib.calls[-1][0] == "mkf_dfm_loglik_compact"
    assert abs(kf.get_mle() - float(g["mle_opt"])) < 1e-9
    np.testing.assert_allclose(kf.sigmas, g["sigmas"], atol=1e-12)      # history: the full grid, as ever
    kf2 = SPKalmanFilter(device_engine=eng)
    kf2.set_observations(np.random.default_rng(0).standard_normal((50, 5)))   # dense data: no compaction
    assert kf2._compact is None
