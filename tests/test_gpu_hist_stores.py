"""Store paths of the thread-per-model history kernel (mkf_dfm_filter_history, row a1 of SURVEY.md 8:
the result contract of seqkalmanfilter, metran/kalmanfilter.py:392-400).

The kernel has two ways of writing fm/fc/pm/pc (TMA tensor stores from a swizzled staging box for
n = 4 and n = 8; shared-memory transposition otherwise) and two ways of writing sigma/detf/cnt (whole
32-byte sectors assembled per lane; scalar stores for unaligned bases).  All of them must write the
same bits to the same places: ragged batches, T that is not a multiple of the sector length, rows
that start in the middle of a sector, missing output arrays, and the oracle on a few models.
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


def _batch(rng, B, T, m, k, miss):
    n = m + k
    alpha = rng.uniform(5, 50, (B, n))
    G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(k)
    phi = np.exp(-1.0 / alpha)
    comm = (G ** 2).sum(2)
    q = np.concatenate([(1 - phi[:, :m] ** 2) * (1 - comm), 1 - phi[:, m:] ** 2], axis=1)
    y = rng.standard_normal((B, T, m))
    y[rng.random(y.shape) < miss] = np.nan
    y[:, T // 2] = np.nan            # a step nobody observes (cnt = 0, detf = 0)
    return y, phi, q, G


def _run(eng, tma, *args, **kw):
    prev = eng.lib.mkf_dfm_set_hist_tma(1 if tma else 0)
    try:
        out = eng.dfm_filter_history(*args, **kw)
        torch.cuda.synchronize()
    finally:
        eng.lib.mkf_dfm_set_hist_tma(prev)
    return out


KEYS = ("sigma", "detf", "cnt", "fm", "fc", "pm", "pc", "mle", "nobs", "status")


@pytest.mark.parametrize("m,k", [(5, 3), (3, 1), (2, 2), (6, 2), (4, 3)])
@pytest.mark.parametrize("B,T", [(203, 97), (64, 128), (31, 5), (33, 3), (1, 1), (70, 100), (40, 98), (37, 16), (5, 17)])
def test_tma_stores_equal_transposed_stores(eng, m, k, B, T):
    """Both store paths, bit for bit, on every array (n = 8, 4: TMA available; n = 7: not)."""
    rng = np.random.default_rng(1000 * m + 10 * k + B)
    y, phi, q, G = _batch(rng, B, T, m, k, 0.2)
    a = _run(eng, True, y, phi, q, G)
    b = _run(eng, False, y, phi, q, G)
    for key in KEYS:
        assert torch.equal(a[key], b[key]), key
    # nothing outside the arrays was touched and every element was written: compare with the
    # warp-group kernel's history as an independent writer
    n = m + k
    A = np.zeros((B, n, n)); Q = np.zeros((B, n, n)); H = np.zeros((B, m, n))
    idx = np.arange(n)
    A[:, idx, idx] = phi; Q[:, idx, idx] = q
    H[:, np.arange(m), np.arange(m)] = 1.0; H[:, :, m:] = G
    ref = eng.filter_history(y, A, Q, H, np.zeros((B, m)), flags=3)
    assert torch.equal(a["cnt"], ref["cnt"])
    for key in ("sigma", "detf", "fm", "fc", "pm", "pc"):
        np.testing.assert_allclose(a[key].cpu().numpy(), ref[key].cpu().numpy(), rtol=0, atol=1e-9, err_msg=key)


def test_history_against_oracle_after_store_rewrite(eng):
    rng = np.random.default_rng(5)
    B, T, m, k = 70, 61, 5, 3
    n = m + k
    y, phi, q, G = _batch(rng, B, T, m, k, 0.3)
    out = _run(eng, True, y, phi, q, G)
    for b in (0, 33, 69):
        A, Q = np.diag(phi[b]), np.diag(q[b])
        H = np.concatenate([np.eye(m), G[b]], axis=1)
        obs, oi, oc = ko.pack_observations(y[b])
        sig, det, _, fm, fc, pm, pc = ko.seqkalmanfilter(obs, A, Q, H, np.zeros(m), oi, oc, np.zeros(n), np.eye(n))
        cnt = out["cnt"][b].cpu().numpy()
        assert np.array_equal(cnt, oc)
        seen = cnt > 0
        ns = int(seen.sum())
        np.testing.assert_allclose(out["sigma"][b].cpu().numpy()[seen], sig[:ns], rtol=0, atol=1e-10)
        np.testing.assert_allclose(out["detf"][b].cpu().numpy()[seen], det[:ns], rtol=0, atol=1e-11)
        assert np.all(out["detf"][b].cpu().numpy()[~seen] == 0.0) and np.all(out["sigma"][b].cpu().numpy()[~seen] == 0.0)
        for key, want in (("fm", fm), ("fc", fc), ("pm", pm), ("pc", pc)):
            np.testing.assert_allclose(out[key][b].cpu().numpy(), want, rtol=0, atol=1e-9, err_msg=key)


def _call_raw(eng, y, phi, q, G, bufs):
    """mkf_dfm_filter_history on caller-provided (possibly offset / missing) output buffers."""
    from metran_b200 import _lib
    from metran_b200.batched import _ptr
    B, T, m = y.shape
    k = G.shape[2]
    mle, nobs, st = eng.empty(B), eng.empty(B, dtype=torch.int64), eng.empty(B, dtype=torch.int32)
    rc = eng.lib.mkf_dfm_filter_history(
        B, T, m, k, _ptr(y), T * m, _ptr(phi), m + k, _ptr(q), m + k, _ptr(G), m * k, _ptr(None), 0, 1,
        _ptr(bufs.get("sigma")), _ptr(bufs.get("detf")), _ptr(bufs.get("cnt")), _ptr(bufs.get("fm")),
        _ptr(bufs.get("fc")), _ptr(bufs.get("pm")), _ptr(bufs.get("pc")), _ptr(mle), _ptr(nobs), _ptr(st), eng._stream())
    _lib.check(eng.lib, rc, "mkf_dfm_filter_history")
    torch.cuda.synchronize()
    return mle


@pytest.mark.parametrize("off", [1, 2, 3])
def test_unaligned_and_missing_outputs(eng, off):
    """Output bases that are only 8-byte aligned (no TMA, scalar stores of the step scalars), and
    calls that ask for a subset of the arrays; guard elements around every buffer stay untouched."""
    rng = np.random.default_rng(77 + off)
    B, T, m, k = 45, 50, 5, 3
    n = m + k
    yh, phih, qh, Gh = _batch(rng, B, T, m, k, 0.15)
    y, phi, q, G = (eng.tensor(v) for v in (yh, phih, qh, Gh))
    ref = _run(eng, True, y, phi, q, G)
    sizes = dict(sigma=B * T, detf=B * T, fm=B * T * n, pm=B * T * n, fc=B * T * n * n, pc=B * T * n * n)
    GUARD = -7.25

    def alloc(keys):
        raw, view = {}, {}
        for key in keys:
            if key == "cnt":
                raw[key] = torch.full((B * T + 16,), -7, dtype=torch.int32, device=eng.device)
                view[key] = raw[key][off:off + B * T]
            else:
                raw[key] = torch.full((sizes[key] + 16,), GUARD, dtype=torch.float64, device=eng.device)
                view[key] = raw[key][off:off + sizes[key]]
        return raw, view

    for keys in (("sigma", "detf", "cnt", "fm", "fc", "pm", "pc"), ("fc",), ("pm", "cnt"), ("sigma", "fm")):
        raw, view = alloc(keys)
        mle = _call_raw(eng, y, phi, q, G, view)
        assert torch.equal(mle, ref["mle"])
        for key in keys:
            assert torch.equal(view[key], ref[key].reshape(-1)), (keys, key)
            guard = raw[key][:off].tolist() + raw[key][off + view[key].numel():].tolist()
            assert all(g == (-7 if key == "cnt" else GUARD) for g in guard), (keys, key)
    # aligned subsets take the TMA path with some arrays missing
    for keys in (("fc",), ("pm", "cnt"), ("fm", "pc", "detf")):
        view = {key: (torch.empty(B * T, dtype=torch.int32, device=eng.device) if key == "cnt"
                      else eng.empty(sizes[key])) for key in keys}
        _call_raw(eng, y, phi, q, G, view)
        for key in keys:
            assert torch.equal(view[key], ref[key].reshape(-1)), (keys, key)


def test_tensor_stores_on_a_batch_of_many_ctas(eng):
    """A batch that fills the GPU several CTAs deep (three per SM resident, more waiting): every element of
    every array equals the transposition path's."""
    rng = np.random.default_rng(404)
    B, T, m, k = 20000, 20, 5, 3
    y, phi, q, G = _batch(rng, B, T, m, k, 0.1)
    a = _run(eng, True, y, phi, q, G)
    b = _run(eng, False, y, phi, q, G)
    for key in KEYS:
        assert torch.equal(a[key], b[key]), key
