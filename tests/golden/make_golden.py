"""Generate the golden fixtures in tests/golden/*.npz from the LIVE reference.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

Every number stored is produced by the unmodified reference code
(/root/reference/metran, loaded through the import shims of _refload.py):
``Metran`` (metran/metran.py) for the two models of the example notebooks and
``SPKalmanFilter`` + ``seqkalmanfilter`` + ``kalmansmoother``
(metran/kalmanfilter.py:236-476, 479-778) for the synthetic state-space cases.
The fixtures pin oracle/ (tests/test_oracle.py) and the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refload  # noqa: E402


def _sel(T):
    """Timesteps at which the long histories are stored (keeps fixtures small)."""
    idx = sorted(set(list(range(0, 25)) + list(range(25, T, 97)) + list(range(T - 25, T))))
    return np.array([i for i in idx if 0 <= i < T], dtype=np.int64)


def metran_model_fixture(mt, name, extra=None, mask_spec=None):
    """Everything the reference facade computes for one solved Metran model."""
    kf = mt.kf
    popt = mt.parameters["optimal"].values.astype(float)
    pinit = mt.parameters["initial"].values.astype(float)
    out = dict(
        y=mt.oseries.values.astype(float),               # standardized, daily grid, NaN = missing
        index_ns=mt.oseries.index.values.astype("datetime64[ns]").astype(np.int64),
        series_names=np.array([str(c) for c in mt.oseries.columns]),
        oseries_std=mt.oseries_std, oseries_mean=mt.oseries_mean,
        factors=mt.factors, eigval=mt.eigval, fep=np.float64(mt.fep),
        alpha_init=pinit, alpha_opt=popt,
        obj_opt=np.float64(mt.fit.obj_func), nfev=np.int64(mt.fit.nfev),
        stderr=mt.parameters["stderr"].values.astype(float),
        pcov=mt.fit.pcov.values.astype(float),
    )
    # objective at given parameter vectors (pin -2logL at GIVEN alpha, not the optimiser path)
    out["mle_init"] = np.float64(mt.get_mle(pinit))
    out["mle_opt"] = np.float64(mt.get_mle(popt))
    rng = np.random.default_rng(7)
    draws = rng.uniform(1.0, 100.0, size=(16, popt.size))
    out["alpha_draws"] = draws
    out["mle_draws"] = np.array([mt.get_mle(d) for d in draws])

    # matrices + full filter/smoother results at the optimum
    p = pd.Series(popt, index=mt.parameters.index)
    A, Q, H, R = mt._get_matrices(p)
    out.update(A_opt=A, Q_opt=Q, H_opt=H, R_opt=R)
    mt._run_kalman("smoother", p=p)
    T = kf.filtered_state_means.shape[0]
    sel = _sel(T)
    out.update(
        sel=sel, sigmas=kf.sigmas.copy(), detfs=kf.detfs.copy(),
        observation_count=kf.observation_count.copy(),
        fm=kf.filtered_state_means[sel], fc=kf.filtered_state_covariances[sel],
        pm=kf.predicted_state_means[sel], pc=kf.predicted_state_covariances[sel],
        sm=kf.smoothed_state_means[sel], sc=kf.smoothed_state_covariances[sel],
        sm_full=kf.smoothed_state_means.copy(),
        svar_full=np.array([np.diag(c) for c in kf.smoothed_state_covariances]),
    )
    # facade-level projections (metran.py:655-942) at the same parameters
    out["state_means"] = mt.get_state_means(p=p).values[sel]
    out["state_variances"] = mt.get_state_variances(p=p).values[sel]
    out["sim_means"] = mt.get_simulated_means(p=p).values[sel]
    out["sim_variances"] = mt.get_simulated_variances(p=p).values[sel]
    out["sim_means_filter"] = mt.get_simulated_means(p=p, method="filter").values[sel]
    last = str(mt.oseries.columns[-1])
    first = str(mt.oseries.columns[0])
    out["simulation_last_ci"] = mt.get_simulation(last, p=p, alpha=0.05).values[sel]
    out["decompose_first"] = mt.decompose_simulation(first, p=p).values[sel]
    out["state0_ci"] = mt.get_state(0, p=p, alpha=0.05).values[sel]
    if mask_spec is not None:
        date, col = mask_spec
        oseries = mt.get_observations()
        mask = (0 * oseries).astype(bool)
        mask.loc[date, col] = True
        row = int(np.where(mask.index == pd.Timestamp(date))[0][0])
        mt.mask_observations(mask)
        sim = mt.get_simulation(col, alpha=0.05)
        out["mask_row"] = np.int64(row)
        out["mask_col"] = np.int64(list(oseries.columns).index(col))
        out["masked_sim_row"] = sim.values[row]
        out["masked_sim_sel"] = sim.values[sel]
        mt.unmask_observations()
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "T=%d" % T, "obj_opt=%r mle_init=%r" % (out["obj_opt"], out["mle_init"]))


def synthetic_cases(metran):
    """State-space cases run through SPKalmanFilter (numba engine) + kalmansmoother."""
    kfmod = metran.kalmanfilter
    rng = np.random.default_rng(20240922)
    cases = {}

    def run(tag, y, A, Q, H, R, x0=None, P0=None):
        kf = kfmod.SPKalmanFilter(engine="numba")
        kf.set_observations(pd.DataFrame(y))
        kf.set_matrices(A, Q, H, R)
        kf.run_filter(initial_state_mean=None if x0 is None else x0.copy(),
                      initial_state_covariance=None if P0 is None else P0.copy())
        sm, sc = kfmod.kalmansmoother(kf.filtered_state_means, kf.filtered_state_covariances,
                                      kf.predicted_state_means, kf.predicted_state_covariances, A)
        n = A.shape[0]
        d = dict(y=y, A=A, Q=Q, H=H, R=R,
                 x0=np.zeros(n) if x0 is None else x0, P0=np.eye(n) if P0 is None else P0,
                 obs=kf.observations, oidx=kf.observation_indices, ocnt=kf.observation_count,
                 sigmas=kf.sigmas, detfs=kf.detfs,
                 fm=kf.filtered_state_means, fc=kf.filtered_state_covariances,
                 pm=kf.predicted_state_means, pc=kf.predicted_state_covariances,
                 sm=sm, sc=sc,
                 mle=np.float64(kf.get_mle()), mle_w0=np.float64(kf.get_mle(warmup=0)),
                 mle_w3=np.float64(kf.get_mle(warmup=3)))
        for k, v in d.items():
            cases[tag + "/" + k] = np.asarray(v)
        print(tag, "mle=%r" % d["mle"])

    def dfm(n, m, T, miss, tag, x0=None, P0=None, gaps=False):
        k = n - m
        alpha = rng.uniform(5, 50, n)
        phi = np.exp(-1.0 / alpha)
        if k > 0:
            G = rng.uniform(0.3, 0.9, (m, k)) / np.sqrt(k)
        else:
            G = np.zeros((m, 0))
        comm = np.sum(G ** 2, axis=1)
        q = np.concatenate([(1 - phi[:m] ** 2) * (1 - comm), 1 - phi[m:] ** 2])
        H = np.concatenate([np.eye(m), G], axis=1)
        x = np.zeros(n)
        y = np.zeros((T, m))
        for t in range(T):
            x = phi * x + np.sqrt(q) * rng.standard_normal(n)
            y[t] = H @ x
        y[rng.random((T, m)) < miss] = np.nan
        if gaps:
            keep = rng.random(T) < 0.15
            keep[0] = True
            y[~keep] = np.nan
        run(tag, y, np.diag(phi), np.diag(q), H, np.zeros(m), x0, P0)

    dfm(8, 5, 240, 0.0, "dfm_n8_m5_dense")
    dfm(8, 5, 240, 0.3, "dfm_n8_m5_miss30")
    dfm(6, 5, 400, 0.1, "dfm_n6_m5_gaps", gaps=True)
    dfm(3, 2, 120, 0.5, "dfm_n3_m2_miss50")
    dfm(5, 5, 90, 0.2, "dfm_n5_m5_nofactor")
    dfm(11, 8, 80, 0.2, "dfm_n11_m8")
    Lx = rng.standard_normal((8, 8)) * 0.3
    dfm(8, 5, 150, 0.2, "dfm_n8_m5_init", x0=rng.standard_normal(8),
        P0=Lx @ Lx.T + 0.5 * np.eye(8))

    # first step unobserved (quirk Q4: nobs/warmup bookkeeping diverge)
    k = "dfm_n8_m5_miss30"
    y = cases[k + "/y"].copy()
    y[0] = np.nan
    y[1, 1:] = np.nan
    run("dfm_n8_m5_t0missing", y, cases[k + "/A"], cases[k + "/Q"], cases[k + "/H"], cases[k + "/R"])

    # values the reference treats as missing: inf and exactly -1e10 (Q2)
    y = cases["dfm_n3_m2_miss50/y"].copy()
    seen = np.argwhere(np.isfinite(y))          # overwrite OBSERVED slots only
    for (t, j), val in zip(seen[[5, 17, 40]], (np.inf, -1e10, -np.inf)):
        y[t, j] = val
    kk = "dfm_n3_m2_miss50"
    run("dfm_n3_m2_specials", y, cases[kk + "/A"], cases[kk + "/Q"], cases[kk + "/H"], cases[kk + "/R"])

    # general dense H, R > 0, m > n (C3-shaped, small T) and m < n
    def dense(n, m, T, miss, tag, fullA=False, fullQ=False):
        alpha = rng.uniform(5, 50, n)
        phi = np.exp(-1.0 / alpha)
        A = np.diag(phi)
        if fullA:   # quirk Q1: mean uses the full A, covariance only its diagonal
            A = A + 0.05 * rng.standard_normal((n, n))
        Qm = np.diag(1 - phi ** 2)
        if fullQ:
            L = 0.2 * rng.standard_normal((n, n))
            Qm = Qm + L @ L.T
        H = rng.standard_normal((m, n)) / np.sqrt(n)
        R = np.full(m, 0.1) * rng.uniform(0.5, 1.5, m)
        x = np.zeros(n)
        y = np.zeros((T, m))
        for t in range(T):
            x = np.diag(A) * x + np.sqrt(1 - phi ** 2) * rng.standard_normal(n)
            y[t] = H @ x + np.sqrt(R) * rng.standard_normal(m)
        y[rng.random((T, m)) < miss] = np.nan
        run(tag, y, A, Qm, H, R)

    dense(32, 64, 30, 0.0, "dense_n32_m64")
    dense(32, 64, 30, 0.25, "dense_n32_m64_miss25")
    dense(16, 7, 60, 0.1, "dense_n16_m7")
    dense(4, 3, 80, 0.1, "dense_n4_m3_fullA", fullA=True)
    dense(5, 3, 80, 0.1, "dense_n5_m3_fullQ", fullQ=True)
    dense(1, 1, 50, 0.2, "dense_n1_m1")
    dense(20, 33, 25, 0.1, "dense_n20_m33")
    np.savez_compressed(os.path.join(HERE, "synthetic_cases.npz"), **cases)


def main():
    metran = _refload.load_reference()

    # (1) examples/example_script.py == tests/conftest.py fixture `mt` (BASELINE config C0)
    mt = metran.Metran(_refload.load_example_series(), name="B21B0214")
    mt.solve(report=False)
    metran_model_fixture(mt, "b21b0214", mask_spec=("1997-8-28", "B21B0214005"))

    # (2) examples/dynamic_factor_model.ipynb cell 7 (seed 20210505)
    np.random.seed(20210505)
    noise = np.random.multivariate_normal(np.zeros(3), np.diag(np.square([1, 0.6, 2])), 2001)
    phi = np.array([0.80, 0.95, 0.90])
    a = np.zeros_like(noise)
    for i in range(1, noise.shape[0]):
        a[i] = noise[i] + np.multiply(a[i - 1], phi)
    s = pd.DataFrame(np.array([a[1:, 0] + a[1:, 2], a[1:, 1] + a[1:, 2]]).T,
                     index=pd.date_range(start="1-1-2000", periods=2000),
                     columns=["series 1", "series 2"])
    mt2 = metran.Metran(s)
    mt2.solve(report=False)
    metran_model_fixture(mt2, "dfm_notebook")

    # (3) synthetic state-space cases through the Kalman object itself
    synthetic_cases(metran)


if __name__ == "__main__":
    main()
