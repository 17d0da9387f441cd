"""Python face of the CPU oracle -- TEST INFRASTRUCTURE, not part of the product.

Restates the reference Kalman path of pastas/metran (file:line are relative to
/root/reference/):

* ``seqkalmanfilter``  -> C, oracle/kalman_oracle.c  (metran/kalmanfilter.py:236-400)
* ``pack_observations``-> numpy                      (metran/kalmanfilter.py:646-674)
* ``get_mle``          -> numpy                      (metran/kalmanfilter.py:550-567)
* ``kalmansmoother``   -> numpy (pinv, like the reference) (metran/kalmanfilter.py:403-476)
* ``simulate`` / ``decompose`` -> numpy              (metran/kalmanfilter.py:569-644)
* ``dfm_matrices``     -> numpy                      (metran/metran.py:246-416)

Pinned against the live reference through tests/golden/*.npz (tests/test_oracle.py) and,
when /root/reference is present, directly (tests/test_oracle_vs_reference.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libkalman_oracle.so")
_lib = None

_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile oracle/kalman_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "kalman_oracle.c")
    if (not force and os.path.isfile(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libkalman_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_seqkalmanfilter.restype = ctypes.c_int64
        L.oracle_seqkalmanfilter.argtypes = (
            [ctypes.c_int64] * 3 + [_f64p] * 6 + [_i64p] + [_f64p] * 2 + [_f64p] * 6)
        L.oracle_filter_loglik.restype = ctypes.c_double
        L.oracle_filter_loglik.argtypes = (
            [ctypes.c_int64] * 3 + [_f64p] * 7 + [ctypes.c_int64, _i64p])
        L.oracle_filter_loglik_batch.restype = None
        L.oracle_filter_loglik_batch.argtypes = (
            [ctypes.c_int64] * 4
            + [_f64p, ctypes.c_int64] * 5
            + [ctypes.c_int64, ctypes.c_int, _f64p, _i64p])
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_f64p) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------- packing
def pack_observations(y):
    """NaN-coded [T,m] -> (observations, observation_indices, observation_count).

    Follows SPKalmanFilter.set_observations (kalmanfilter.py:646-674): a slot is
    observed when it is finite and ``value + 1e10 != 0``; indices are float64,
    ascending, zero padded; observations hold 0.0 at missing slots.
    """
    y = np.asarray(y, dtype=np.float64)
    T, m = y.shape
    finite = np.isfinite(y)
    with np.errstate(invalid="ignore"):
        observed = finite & ((np.where(finite, y, 0.0) + 1e10) != 0.0)
    observations = np.where(observed, y, 0.0)
    count = observed.sum(axis=1).astype(np.int64)
    indices = np.zeros((T, m), dtype=np.float64)
    for t in range(T):
        w = np.nonzero(observed[t])[0]
        indices[t, : w.size] = w
    return observations, indices, count


# ------------------------------------------------------------------------------ filter
def seqkalmanfilter(observations, A, Q, H, R, indices, count, x0, P0, history=True):
    """Reference 9-in / 7-out contract (kalmanfilter.py:243-253, 392-400)."""
    obs, A, Q, H, R = _c(observations), _c(A), _c(Q), _c(H), _c(R)
    idx, x0, P0 = _c(indices), _c(x0), _c(P0)
    cnt = np.ascontiguousarray(count, dtype=np.int64)
    T, m = obs.shape
    n = x0.shape[0]
    sig = np.zeros(T)
    det = np.zeros(T)
    if history:
        fm, fc = np.zeros((T, n)), np.zeros((T, n, n))
        pm, pc = np.zeros((T, n)), np.zeros((T, n, n))
    else:
        fm = fc = pm = pc = None
    ns = lib().oracle_seqkalmanfilter(T, n, m, _p(obs), _p(A), _p(Q), _p(H), _p(R), _p(idx),
                                      cnt.ctypes.data_as(_i64p), _p(x0), _p(P0),
                                      _p(sig), _p(det), _p(fm), _p(fc), _p(pm), _p(pc))
    if ns < 0:
        raise ValueError("oracle_seqkalmanfilter: bad sizes")
    return sig, det, int(ns), fm, fc, pm, pc


def get_mle(sigmas, detfs, count, warmup=1):
    """-2 log L (kalmanfilter.py:550-567), numpy pairwise sums like the reference."""
    nobs = np.sum(count[warmup:])
    return nobs * np.log(2 * np.pi) + np.sum(detfs[warmup:]) + np.sum(sigmas[warmup:])


def filter_loglik(y, A, Q, H, R, x0=None, P0=None, warmup=1):
    """pack -> filter -> get_mle for one NaN-coded model; returns (mle, nobs)."""
    obs, idx, cnt = pack_observations(y)
    n = np.asarray(A).shape[0]
    x0 = np.zeros(n) if x0 is None else x0          # kalmanfilter.py:747-750
    P0 = np.eye(n) if P0 is None else P0
    sig, det, ns, *_ = seqkalmanfilter(obs, A, Q, H, R, idx, cnt, x0, P0, history=False)
    return float(get_mle(sig[:ns], det[:ns], cnt, warmup)), int(np.sum(cnt[warmup:]))


def filter_loglik_batch(y, A, Q, H, R, warmup=1, threads=0):
    """All-cores batch: y[B,T,m] (or [T,m] shared), A/Q[B,n,n], H[B,m,n], R[B,m].

    Sequential-sum reduction in C (differs from np.sum pairwise by <= 1e-10).
    """
    A, Q, H, R = _c(A), _c(Q), _c(H), _c(R)
    y = _c(y)
    B = A.shape[0] if A.ndim == 3 else (y.shape[0] if y.ndim == 3 else 1)
    n = A.shape[-1]
    m = H.shape[-2]
    T = y.shape[-2]

    def st(a, base):
        return 0 if a.ndim == base else int(np.prod(a.shape[1:]))

    mle = np.zeros(B)
    nobs = np.zeros(B, dtype=np.int64)
    lib().oracle_filter_loglik_batch(B, T, n, m, _p(y), st(y, 2), _p(A), st(A, 2), _p(Q), st(Q, 2),
                                     _p(H), st(H, 2), _p(R), st(R, 1), warmup, int(threads),
                                     _p(mle), nobs.ctypes.data_as(_i64p))
    return mle, nobs


def max_threads():
    return int(lib().oracle_max_threads())


# ---------------------------------------------------------------------------- smoother
def kalmansmoother(fm, fc, pm, pc, A):
    """RTS backward pass exactly as kalmanfilter.py:453-474 (pinv of P_pred[t+1])."""
    T, n = fm.shape
    sm = np.zeros((T, n))
    sc = np.zeros((T, n, n))
    sm[-1] = fm[-1]
    sc[-1] = fc[-1]
    At = np.asarray(A).T
    for t in range(T - 2, -1, -1):
        G = fc[t] @ (At @ np.linalg.pinv(pc[t + 1]))
        sm[t] = fm[t] + G @ (sm[t + 1] - pm[t + 1])
        sc[t] = fc[t] + G @ ((sc[t + 1] - pc[t + 1]) @ G.T)
    return sm, sc


# ------------------------------------------------------------------------- projections
def simulate(means, covariances, observation_matrix):
    """H x_t and max(diag(H P_t H^T), 0) (kalmanfilter.py:569-603, quirk Q11)."""
    Hs = np.asarray(observation_matrix)
    sim_means = means @ Hs.T
    var = np.einsum("ij,tjk,ik->ti", Hs, covariances, Hs)
    return sim_means, np.maximum(var, 0)


def decompose(means, observation_matrix, nsdf):
    """sdf / cdf contributions (kalmanfilter.py:605-644)."""
    Hs = np.asarray(observation_matrix)
    sdf = means[:, :nsdf] @ Hs[:, :nsdf].T
    cdf = [np.outer(means[:, j], Hs[:, j]) for j in range(nsdf, Hs.shape[1])]
    return sdf, cdf


# ---------------------------------------------------------------------------- matrices
def dfm_matrices(alpha, factors, dt=1.0):
    """alpha[n] -> (A, Q, H, R) of Metran's DFM (metran.py:246-416).

    phi = exp(-dt/alpha) (262-263); A = diag(phi) (283-290); Q = diag((1-phi^2)(1-comm))
    for the specific factors and (1-phi^2) for the common ones (310-322); H = [I | factors]
    (365-370); R = 0 (382-384).  State order: series first, then common factors (Q10).
    """
    factors = np.asarray(factors, dtype=np.float64)
    m, k = factors.shape
    alpha = np.asarray(alpha, dtype=np.float64)
    phi = np.exp(-dt / alpha)
    comm = np.sum(np.square(factors), axis=1)
    q = np.empty(m + k)
    q[:m] = (1 - phi[:m] ** 2) * (1 - comm)
    q[m:] = 1 - phi[m:] ** 2
    H = np.zeros((m, m + k))
    H[:, :m] = np.eye(m)
    H[:, m:] = factors
    return np.diag(phi), np.diag(q), H, np.zeros(m)
