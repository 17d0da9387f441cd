"""Three-launch path of the thread-per-model kernel (transient -> steady-state kernel -> resume,
metran_b200/csrc/dfm_steady.cuh) against the single-kernel path of the same library.

Both paths apply the same functions of dfm_collapsed.cuh in the same order, so the per-model
-2 log L must agree BIT FOR BIT: dense data (the steady kernel runs to the end of the series),
gaps after the switch (the steady kernel hands the warp back at exactly the incomplete step and
the resume launch re-converges), shared observations, ragged batches, every warmup, and T that is
not a multiple of any tile length.  The oracle comparison of the same inputs is in
test_gpu_zz_steady_state.py / test_gpu_fullsize.py (they run with the hand-off enabled).
"""
import numpy as np
import pytest
import torch

import steady_state_cases as ssc
from oracle import kalman_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from metran_b200.batched import KalmanEngine

    return KalmanEngine()


def _both(eng, *args, **kw):
    """(handoff path, single-kernel path) of dfm_loglik_alpha on the same inputs."""
    prev = eng.lib.mkf_dfm_set_handoff(1)
    try:
        on = [t.clone() for t in eng.dfm_loglik_alpha(*args, **kw)]
        eng.lib.mkf_dfm_set_handoff(0)
        off = [t.clone() for t in eng.dfm_loglik_alpha(*args, **kw)]
    finally:
        eng.lib.mkf_dfm_set_handoff(prev)
    torch.cuda.synchronize()
    return on, off


def _dense_batch(m, k, B, T, seed, amax=30.0):
    rng = np.random.default_rng(seed)
    n = m + k
    G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(k)
    alpha = rng.uniform(2.0, amax, (B, n))
    phi = np.exp(-1.0 / alpha)
    q = np.concatenate([(1 - phi[:, :m] ** 2) * (1 - (G ** 2).sum(2)), 1 - phi[:, m:] ** 2], axis=1)
    x = np.zeros((B, n))
    y = np.empty((B, T, m))
    for t in range(T):
        x = phi * x + np.sqrt(q) * rng.standard_normal((B, n))
        y[:, t] = x[:, :m] + np.einsum("bmk,bk->bm", G, x[:, m:])
    return y, alpha, G


@pytest.mark.parametrize("m,k", ssc.SHAPES)
@pytest.mark.parametrize("T", [1201, 1500])
def test_dense_bit_identical(eng, m, k, T):
    y, alpha, G = _dense_batch(m, k, 70, T, seed=7 * m + k)
    on, off = _both(eng, y, alpha, G)
    assert torch.equal(on[0], off[0]) and torch.equal(on[1], off[1]) and torch.equal(on[2], off[2])
    assert torch.isfinite(on[0]).all()


@pytest.mark.parametrize("m,k", ssc.SHAPES)
def test_gaps_after_switch_bit_identical(eng, m, k):
    """The gaps of steady_state_cases: a warp-wide one at t = 1000 and a single lane at t = 1900
    (odd and even positions relative to the 2-step tiles), plus the oracle on the same batch."""
    y, alpha, phi, q, G = ssc.make_batch(m, k)
    y[5, 1501, m - 1] = np.nan
    on, off = _both(eng, y, alpha, G)
    assert torch.equal(on[0], off[0]) and torch.equal(on[1], off[1])
    want, wobs = ssc.oracle_values(y, phi, q, G)
    np.testing.assert_allclose(on[0].cpu().numpy(), want, rtol=0, atol=1e-9)
    assert np.array_equal(on[1].cpu().numpy(), wobs)


@pytest.mark.parametrize("warmup", [0, 1, 3, 700, 5000])
def test_warmup_bit_identical(eng, warmup):
    y, alpha, G = _dense_batch(5, 3, 40, 1400, seed=3)
    on, off = _both(eng, y, alpha, G, warmup=warmup)
    assert torch.equal(on[0], off[0]) and torch.equal(on[1], off[1])
    mats = [ko.dfm_matrices(alpha[b], G[b]) for b in range(8)]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    want, wobs = ko.filter_loglik_batch(y[:8], A, Q, H, R, warmup=warmup)
    np.testing.assert_allclose(on[0][:8].cpu().numpy(), want, rtol=0, atol=1e-9)
    assert np.array_equal(on[1][:8].cpu().numpy(), wobs)


def test_shared_observations_bit_identical(eng):
    """One observation block for all parameter draws (y_stride = 0): the steady kernel stages one
    row per tile and every lane reads it."""
    y, alpha, G = _dense_batch(5, 3, 1, 1300, seed=5)
    rng = np.random.default_rng(6)
    draws = rng.uniform(2.0, 30.0, (100, 8))
    on, off = _both(eng, y[0], draws, G[0])
    assert torch.equal(on[0], off[0])
    mats = [ko.dfm_matrices(draws[b], G[0]) for b in range(6)]
    A, Q, H, R = (np.stack([mm[i] for mm in mats]) for i in range(4))
    want, _ = ko.filter_loglik_batch(np.repeat(y, 6, axis=0), A, Q, H, R)
    np.testing.assert_allclose(on[0][:6].cpu().numpy(), want, rtol=0, atol=1e-9)


def test_handoff_actually_taken(eng):
    """The three-launch path must be the one that ran: with the hand-off on, a second stream-ordered
    scratch buffer exists and the steady kernel's share shows up as a speed-up on a long dense batch."""
    y, alpha, G = _dense_batch(5, 3, 4096, 3000, seed=9, amax=20.0)
    y_d, a_d, g_d = eng.tensor(y), eng.tensor(alpha), eng.tensor(G)

    def timed(flag):
        prev = eng.lib.mkf_dfm_set_handoff(flag)
        try:
            eng.dfm_loglik_alpha(y_d, a_d, g_d)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.dfm_loglik_alpha(y_d, a_d, g_d)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1), out[0].clone()
        finally:
            eng.lib.mkf_dfm_set_handoff(prev)

    t_on, m_on = timed(1)
    t_off, m_off = timed(0)
    assert torch.equal(m_on, m_off)
    assert t_on < t_off, "hand-off path (%.3f ms) not faster than the single kernel (%.3f ms)" % (t_on, t_off)


def test_batch_permutation_dense_bit_identical_and_gaps_within_1e_10(eng):
    """for fully observed data a model's value depends on its own inputs only
    (any permutation / split of the batch gives the same bits); with gaps the warp-level switch points
    move results at the 1e-13 relative level."""
    y, alpha, G = _dense_batch(5, 3, 96, 1300, seed=11)
    perm = np.random.default_rng(0).permutation(96)
    a = eng.dfm_loglik_alpha(y, alpha, G)[0]
    b = eng.dfm_loglik_alpha(y[perm].copy(), alpha[perm].copy(), G[perm].copy())[0]
    assert torch.equal(a[torch.as_tensor(perm, device=a.device)], b)
    c = eng.dfm_loglik_alpha(y[:40].copy(), alpha[:40].copy(), G[:40].copy())[0]
    assert torch.equal(a[:40], c)
    yg = y.copy()
    yg[7, 900:903] = np.nan
    yg[50, 1100, 2] = np.nan
    ag = eng.dfm_loglik_alpha(yg, alpha, G)[0]
    bg = eng.dfm_loglik_alpha(yg[perm].copy(), alpha[perm].copy(), G[perm].copy())[0]
    assert float((ag[torch.as_tensor(perm, device=a.device)] - bg).abs().max()) <= 1e-10
