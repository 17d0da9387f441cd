"""Shared body of the steady-state parity check (used by a CPU test through the oracle-backed
stand-in library, to prove the test itself, and by the GPU test through the CUDA library).

Long fully observed series of several (m, k) shapes, so that the thread-per-model kernel reaches
its steady-state step, with gaps that force it to leave and re-enter it;
two warps converge quickly (alpha 2..30 days: the whole warp switches), the ragged third one mixes
fast and slow models (alpha up to 400 days: lanes hold their covariance, the warp never switches).
"""
import numpy as np

from oracle import kalman_oracle as ko

SHAPES = [(3, 2), (6, 2), (2, 1), (4, 3), (5, 1), (1, 3), (5, 3)]


def make_batch(m, k, B=70, T=2000, seed=0):
    rng = np.random.default_rng(1000 * m + 10 * k + seed)
    n = m + k
    G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(k)
    alpha = np.exp(rng.uniform(np.log(2.0), np.log(400.0), (B, n)))
    alpha[:64] = rng.uniform(2.0, 30.0, (64, n))       # warps 0 and 1: every lane converges quickly
    phi = np.exp(-1.0 / alpha)
    comm = (G ** 2).sum(2)
    q = np.concatenate([(1 - phi[:, :m] ** 2) * (1 - comm), 1 - phi[:, m:] ** 2], axis=1)
    x = np.zeros((B, n))
    y = np.empty((B, T, m))
    for t in range(T):
        x = phi * x + np.sqrt(q) * rng.standard_normal((B, n))
        y[:, t] = x[:, :m] + np.einsum("bmk,bk->bm", G, x[:, m:])
    y[32:64, 1000:1003] = np.nan          # a warp-wide gap after the switch: leave and re-enter
    y[40, 1900, 0] = np.nan               # a single lane pulls its warp out once more
    y[64:, :7] = np.nan                   # ragged warp starts late
    return y, alpha, phi, q, G


def oracle_values(y, phi, q, G, warmup=1):
    B, T, m = y.shape
    n = phi.shape[1]
    A = np.zeros((B, n, n)); Q = np.zeros((B, n, n)); H = np.zeros((B, m, n))
    i = np.arange(n)
    A[:, i, i] = phi
    Q[:, i, i] = q
    H[:, np.arange(m), np.arange(m)] = 1.0
    H[:, :, m:] = G
    return ko.filter_loglik_batch(y, A, Q, H, np.zeros((B, m)), warmup=warmup)


def check(eng, m, k, to_numpy):
    y, alpha, phi, q, G = make_batch(m, k)
    want, wobs = oracle_values(y, phi, q, G)
    mle, nobs, status = eng.dfm_loglik_alpha(y, alpha, G)
    got = to_numpy(mle)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)
    assert np.array_equal(to_numpy(nobs), wobs)
    assert not to_numpy(status).any()
    # a model's value must not depend on its neighbours in the warp: reversed batch order
    mle_r, _, _ = eng.dfm_loglik_alpha(y[::-1].copy(), alpha[::-1].copy(), G[::-1].copy())
    np.testing.assert_allclose(to_numpy(mle_r)[::-1], got, rtol=0, atol=1e-9)
    return got
