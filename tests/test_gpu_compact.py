"""Compacted time grid (mkf_dfm_loglik_compact) on the GPU.

The reference forces a daily grid (metran.py:571): on its own example data 94.5 % of the steps
carry no observation and only predict (kalmanfilter.py:335 not taken).  The compact kernel walks
the observed steps and applies the k predicts in between in closed form.  Checked against the 18
golden -2 log L values of the reference on B21B0214, against the full-grid kernel and the oracle
on random ragged batches (per-model patterns, all-missing models, sentinels, every warmup), and
for bit-identity with the full-grid kernel when no step is empty.
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


def test_b21_golden_values_on_the_compact_grid(eng, b21):
    from metran_b200.dfm import DynamicFactorModel

    g = b21
    mdl = DynamicFactorModel(g["y"], g["factors"], engine=eng)
    assert mdl.kf._compact is not None and mdl.kf._compact[0].shape[0] == 343
    assert abs(mdl.get_mle(g["alpha_init"]) - 2384.792799342231) <= 1e-9
    assert abs(mdl.get_mle(g["alpha_opt"]) - 2332.3270693771483) <= 1e-9
    got = mdl.get_mle_batch(g["alpha_draws"])
    assert np.max(np.abs(got - g["mle_draws"])) <= 1e-9
    # the same 18 values on the full grid, for the record of the closed form's rounding
    full = eng.dfm_loglik_alpha(g["y"], np.vstack([g["alpha_init"], g["alpha_opt"], g["alpha_draws"]]), g["factors"])
    comp = eng.dfm_loglik_compact(mdl.kf._compact[0], mdl.kf._compact[1], g["factors"],
                                  alpha=np.vstack([g["alpha_init"], g["alpha_opt"], g["alpha_draws"]]))
    assert float((full[0] - comp[0]).abs().max()) <= 1e-9 and torch.equal(full[1], comp[1])


def _sparse_batch(rng, B, T, m, k, dense_frac):
    n = m + k
    alpha = rng.uniform(3.0, 60.0, (B, n))
    G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(max(k, 1))
    phi = np.exp(-1.0 / alpha)
    q = np.concatenate([(1 - phi[:, :m] ** 2) * (1 - (G ** 2).sum(2)), 1 - phi[:, m:] ** 2], axis=1)
    x = np.zeros((B, n))
    y = np.empty((B, T, m))
    for t in range(T):
        x = phi * x + np.sqrt(q) * rng.standard_normal((B, n))
        y[:, t] = x[:, :m] + np.einsum("bmk,bk->bm", G, x[:, m:])
    keep = rng.random((B, T)) < rng.uniform(0.02, dense_frac, (B, 1))
    y[~keep] = np.nan
    y[rng.random(y.shape) < 0.25] = np.nan
    return y, alpha, G


@pytest.mark.parametrize("m,k", [(5, 1), (5, 3), (3, 2), (8, 3), (1, 0)])
@pytest.mark.parametrize("warmup", [0, 1, 3])
def test_ragged_sparse_batches(eng, m, k, warmup):
    rng = np.random.default_rng(100 * m + 10 * k + warmup)
    B, T = 75, 900
    y, alpha, G = _sparse_batch(rng, B, T, m, k, 0.3)
    y[4] = np.nan                                   # no observation at all
    y[5, :40] = np.nan                              # late start
    y[6, 100, 0] = -1e10
    y[6, 101, m - 1] = np.inf
    yc, tidx, J, frac = eng.compact_observations(y)
    assert frac < 0.35
    full = eng.dfm_loglik_alpha(y, alpha, G, warmup=warmup)
    comp = eng.dfm_loglik_compact(yc, tidx, G, alpha=alpha, warmup=warmup)
    d = (full[0] - comp[0]).abs()
    d[4] = 0.0
    assert float(d.max()) <= 1e-9, float(d.max())
    assert torch.equal(full[1], comp[1]) and int(comp[2].abs().sum()) == 0
    idx = [0, 5, 6, 33, 74]
    mats = [ko.dfm_matrices(alpha[b], G[b]) for b in idx]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    want, wobs = ko.filter_loglik_batch(y[idx], A, Q, H, R, warmup=warmup)
    np.testing.assert_allclose(comp[0].cpu().numpy()[idx], want, rtol=0, atol=1e-9)
    assert np.array_equal(comp[1].cpu().numpy()[idx], wobs)


def test_phi_q_r_form_and_shared_block(eng):
    rng = np.random.default_rng(5)
    y, alpha, G = _sparse_batch(rng, 1, 1500, 5, 3, 0.1)
    draws = rng.uniform(2.0, 80.0, (200, 8))
    yc, tidx, J, _ = eng.compact_observations(y[0])
    full = eng.dfm_loglik_alpha(y[0], draws, G[0])
    comp = eng.dfm_loglik_compact(yc, tidx, G[0], alpha=draws)
    assert float((full[0] - comp[0]).abs().max()) <= 1e-9 and torch.equal(full[1], comp[1])
    phi = np.exp(-1.0 / draws)
    q = np.concatenate([(1 - phi[:, :5] ** 2) * (1 - (G[0] ** 2).sum(1)), 1 - phi[:, 5:] ** 2], axis=1)
    r = np.full((200, 5), 0.05)
    fullr = eng.dfm_loglik(y[0], phi, q, G[0], r=r)
    compr = eng.dfm_loglik_compact(yc, tidx, G[0], phi=phi, q=q, r=r)
    assert float((fullr[0] - compr[0]).abs().max()) <= 1e-9


def test_no_empty_steps_is_bit_identical_to_the_full_grid_kernel(eng):
    """k = 1 everywhere: the compact kernel takes the ordinary predict (same build as the R != 0
    path), so with a tiny R it must reproduce mkf_dfm_loglik bit for bit."""
    rng = np.random.default_rng(6)
    y, alpha, G = _sparse_batch(rng, 40, 300, 5, 3, 0.3)
    y = np.where(np.isnan(y).all(-1, keepdims=True), 0.1, y)          # every step has an observation
    phi = np.exp(-1.0 / alpha)
    q = np.concatenate([(1 - phi[:, :5] ** 2) * (1 - (G ** 2).sum(2)), 1 - phi[:, 5:] ** 2], axis=1)
    r = np.full((40, 5), 1e-3)
    yc, tidx, J, frac = eng.compact_observations(y)
    assert J == 300
    a = eng.dfm_loglik(y, phi, q, G, r=r)
    b = eng.dfm_loglik_compact(yc, tidx, G, phi=phi, q=q, r=r)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
