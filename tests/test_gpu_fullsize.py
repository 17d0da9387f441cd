"""Parity at BASELINE.json's full sizes (configs C1, C2, C3, one C4 shard).

Where the oracle finishes in seconds the whole batch is compared (C1, C2); the larger
configurations are checked through size-independent properties: random sub-batches against the
oracle, bit-identical results when the batch is split or permuted, the pooled sum, and the
equivalence of an all-missing series with a model that does not have that series.
"""
import numpy as np
import pytest
import torch

from oracle import kalman_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from metran_b200.batched import KalmanEngine

    return KalmanEngine()


def _params(rng, B, m, k):
    alpha = rng.uniform(5.0, 50.0, (B, m + k))
    gamma = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(max(k, 1))
    return alpha, gamma


def _simulate(dev, seed, alpha, gamma, T, miss=0.0):
    import bench

    return bench.simulate_gpu(torch, dev, seed, alpha, gamma, T, miss)


def _oracle_subset(y, alpha, gamma, idx):
    mats = [ko.dfm_matrices(alpha[b], gamma[b]) for b in idx]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    return ko.filter_loglik_batch(y, A, Q, H, R)


def test_c1_full_batch_against_oracle(eng):
    """configs[1]: 1024 models, state_dim 8, obs_dim 5, T = 2000: every model vs the oracle."""
    rng = np.random.default_rng(11)
    B, T, m, k = 1024, 2000, 5, 3
    alpha, gamma = _params(rng, B, m, k)
    a_d, g_d = eng.tensor(alpha), eng.tensor(gamma)
    y = _simulate(eng.device, 3, a_d, g_d, T, miss=0.1)
    mle, nobs, status = eng.dfm_loglik_alpha(y, a_d, g_d)
    want, wobs = _oracle_subset(y.cpu().numpy(), alpha, gamma, range(B))
    got = mle.cpu().numpy()
    assert np.max(np.abs(got - want)) <= 1e-9      # north_star: <= 1e-9 absolute on -2 log L
    assert np.array_equal(nobs.cpu().numpy(), wobs) and int(status.abs().sum()) == 0
    pooled = float(eng.reduce_sum(mle)[0])
    assert abs(pooled - np.sum(want)) <= 1e-12 * abs(np.sum(want))


def test_c2_solver_sweep_8192_draws(eng, b21):
    """configs[2]: 8192 parameter draws of ONE model (shared y, shared loadings)."""
    rng = np.random.default_rng(12)
    draws = rng.uniform(1.0, 100.0, (8192, 6))
    mle, _, status = eng.dfm_loglik_alpha(b21["y"], draws, b21["factors"])
    got = mle.cpu().numpy()
    assert int(status.abs().sum()) == 0 and np.all(np.isfinite(got))
    pick = rng.choice(8192, 96, replace=False)
    mats = [ko.dfm_matrices(draws[i], b21["factors"]) for i in pick]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    want, _ = ko.filter_loglik_batch(b21["y"], A, Q, H, R)
    np.testing.assert_allclose(got[pick], want, rtol=0, atol=1e-9)
    # the golden draws of the reference are reproduced inside the big batch too
    mle2, _, _ = eng.dfm_loglik_alpha(b21["y"], np.concatenate([b21["alpha_draws"], draws[:100]]), b21["factors"])
    np.testing.assert_allclose(mle2.cpu().numpy()[:16], b21["mle_draws"], rtol=0, atol=1e-9)


def test_c4_shard_properties(eng):
    """configs[4] per-GPU shard: 32768 models, T = 10000 (13 GB of observations)."""
    rng = np.random.default_rng(13)
    B, T, m, k = 32768, 10000, 5, 3
    alpha, gamma = _params(rng, B, m, k)
    a_d, g_d = eng.tensor(alpha), eng.tensor(gamma)
    y = _simulate(eng.device, 4, a_d, g_d, T)
    mle, nobs, status = eng.dfm_loglik_alpha(y, a_d, g_d)
    got = mle.cpu().numpy()
    assert int(status.abs().sum()) == 0 and np.all(np.isfinite(got))
    assert np.all(nobs.cpu().numpy() == (T - 1) * m)
    # (1) random sub-batch against the oracle
    pick = np.sort(rng.choice(B, 48, replace=False))
    want, _ = _oracle_subset(y[torch.as_tensor(pick, device=y.device)].cpu().numpy(), alpha, gamma, pick)
    np.testing.assert_allclose(got[pick], want, rtol=0, atol=1e-9)
    # (2) splitting the batch gives bit-identical per-model results
    lo, _, _ = eng.dfm_loglik_alpha(y[: B // 2 + 7], a_d[: B // 2 + 7], g_d[: B // 2 + 7])
    hi, _, _ = eng.dfm_loglik_alpha(y[B // 2 + 7:], a_d[B // 2 + 7:], g_d[B // 2 + 7:])
    assert torch.equal(torch.cat([lo, hi]), mle)
    # (3) pooled objective: deterministic reduction == exact sum of the per-model values
    import math

    pooled = float(eng.reduce_sum(mle)[0])
    assert pooled == float(eng.reduce_sum(mle)[0])
    assert abs(pooled - math.fsum(got.tolist())) <= 1e-13 * abs(pooled)
    # (4) a prefix in time equals the filter run on the prefix (causality of the recursion)
    Tp = 2502
    pre, _, _ = eng.dfm_loglik_alpha(y[:256, :Tp].contiguous(), a_d[:256], g_d[:256])
    want_p, _ = _oracle_subset(y[:8, :Tp].cpu().numpy(), alpha, gamma, range(8))
    np.testing.assert_allclose(pre.cpu().numpy()[:8], want_p, rtol=0, atol=1e-9)


def test_all_missing_series_equals_dropped_series(eng):
    """A series that is never observed must not change -2logL: (m, k) vs (m-1, k) kernels."""
    rng = np.random.default_rng(14)
    B, T, m, k = 300, 400, 5, 2
    alpha, gamma = _params(rng, B, m, k)
    y = _simulate(eng.device, 5, eng.tensor(alpha), eng.tensor(gamma), T, miss=0.2)
    y[:, :, 2] = float("nan")
    full, nf, _ = eng.dfm_loglik_alpha(y, alpha, gamma)
    keep = [0, 1, 3, 4]
    sub, ns, _ = eng.dfm_loglik_alpha(y[:, :, keep].contiguous(), alpha[:, keep + [5, 6]], gamma[:, keep])
    np.testing.assert_allclose(full.cpu().numpy(), sub.cpu().numpy(), rtol=1e-13, atol=1e-10)
    assert torch.equal(nf, ns)


def test_c3_filter_smoother_full_size(eng):
    """configs[3]: 256 models, state_dim 32, obs_dim 64, T = 5000, filter + RTS smoother."""
    rng = np.random.default_rng(15)
    B, T, n, m = 256, 5000, 32, 64
    dev = eng.device
    g = torch.Generator(device=dev)
    g.manual_seed(6)
    phi = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 0.18 + 0.8
    H = torch.randn(B, m, n, dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
    R = torch.full((B, m), 0.1, dtype=torch.float64, device=dev)
    A = torch.diag_embed(phi)
    Q = torch.diag_embed(1 - phi ** 2)
    x = torch.zeros(B, n, dtype=torch.float64, device=dev)
    y = torch.empty(B, T, m, dtype=torch.float64, device=dev)
    sq = torch.sqrt(1 - phi ** 2)
    for t in range(T):
        x = phi * x + sq * torch.randn(B, n, dtype=torch.float64, device=dev, generator=g)
        y[:, t] = torch.einsum("bmn,bn->bm", H, x) + np.sqrt(0.1) * torch.randn(B, m, dtype=torch.float64, device=dev, generator=g)
    out = eng.filter_history(y, A, Q, H, R, flags=3)
    sm, sc, status = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], A, flags=1)
    assert int(status.abs().sum()) == 0 and int(out["status"].abs().sum()) == 0
    ll, _, _ = eng.filter_loglik(y, A, Q, H, R, flags=3)
    assert torch.equal(ll, out["mle"])                       # history and loglik kernels agree bitwise
    for b in (0, 255):
        yb = y[b].cpu().numpy()
        Ab, Qb, Hb, Rb = (v[b].cpu().numpy() for v in (A, Q, H, R))
        obs, idx, cnt = ko.pack_observations(yb)
        sig, det, ns, fm, fc, pm, pc = ko.seqkalmanfilter(obs, Ab, Qb, Hb, Rb, idx, cnt, np.zeros(n), np.eye(n))
        want = ko.get_mle(sig[:ns], det[:ns], cnt)
        assert abs(float(out["mle"][b]) - want) <= 1e-12 * abs(want)     # |mle| ~ 3e5
        np.testing.assert_allclose(out["fm"][b].cpu().numpy(), fm, rtol=0, atol=1e-9)
        np.testing.assert_allclose(out["fc"][b].cpu().numpy(), fc, rtol=0, atol=1e-9)
        tail = 600                                           # numpy pinv smoother: last 600 steps
        s1, s2 = ko.kalmansmoother(fm[-tail:], fc[-tail:], pm[-tail:], pc[-tail:], Ab)
        np.testing.assert_allclose(sm[b, -tail:].cpu().numpy(), s1, rtol=0, atol=1e-8)
        np.testing.assert_allclose(sc[b, -tail:].cpu().numpy(), s2, rtol=0, atol=1e-8)
    # smoothing never increases the variance; last step equals the filter
    assert torch.all(torch.diagonal(sc, dim1=2, dim2=3) <= torch.diagonal(out["fc"], dim1=2, dim2=3) + 1e-12)
    assert torch.equal(sm[:, -1], out["fm"][:, -1])
