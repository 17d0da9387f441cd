"""numba face of the CPU oracle -- TEST INFRASTRUCTURE / CPU baseline, not part of the product.

The reference's hot loop is a numba ``@njit`` function (``seqkalmanfilter``,
/root/reference/metran/kalmanfilter.py:236-400) driven one model at a time by
``SPKalmanFilter.run_filter`` (:696-778) and reduced by ``get_mle`` (:550-567).  This module
restates that path with the SAME technology -- numba-compiled scalar loops, IEEE fp64, no
fast-math -- so that ``bench.py`` can time "the numba CPU path" BASELINE.json names next to the C
port of oracle/kalman_oracle.c (which gcc makes about twice as fast per core).  The arithmetic
follows kalman_oracle.c statement by statement (same association of every sum), so the two agree
to rounding of the final long-double reduction; tests/test_oracle_numba.py pins this module to the
golden vectors of the unmodified reference and to the C oracle.

Only ``tests/`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
"""
import math

import numpy as np

try:
    import numba
    from numba import njit, prange

    HAVE_NUMBA = True
except Exception:  # pragma: no cover - numba is part of the image; the C port is used without it
    HAVE_NUMBA = False

    def njit(*a, **k):
        def deco(f):
            return f
        return deco if not (a and callable(a[0])) else a[0]

    prange = range


@njit(cache=True, fastmath=False)
def _loglik_one(y, A, Q, H, R, warmup):
    """-2 log L of one model from its NaN-coded observations y[T, m].

    kalmanfilter.py:315-390 (predict with the full A for the mean and diag(A) for the covariance,
    one scalar update per observed series), missing-data rule of set_observations (:657,
    :666-667), reduction of get_mle (:563-566, quirk Q4: the first `warmup` OBSERVED steps are
    dropped from the sums, nobs counts raw steps t >= warmup).  x0 = 0, P0 = I (:747-750).
    """
    T, m = y.shape
    n = A.shape[0]
    x = np.zeros(n)
    xp = np.zeros(n)
    d = np.zeros(n)
    g = np.zeros(n)
    P = np.eye(n)
    nsig = 0
    nobs = 0
    sum_sigma = 0.0
    sum_detf = 0.0
    for t in range(T):
        for r in range(n):                       # :318-322
            s = 0.0
            for c in range(n):
                s += A[r, c] * x[c]
            xp[r] = s
        for r in range(n):                       # :324-331
            for c in range(n):
                P[r, c] = A[r, r] * P[r, c] * A[c, c] + Q[r, c]
        cnt = 0
        sigma = 0.0
        detf = 0.0
        for j in range(m):
            yj = y[t, j]
            if not math.isfinite(yj) or yj + 1e10 == 0.0:
                continue
            cnt += 1
            s = 0.0
            for r in range(n):
                s += H[j, r] * xp[r]
            v = yj - s                           # :344-347
            for r in range(n):                   # :349-357
                a = 0.0
                for c in range(n):
                    a += P[r, c] * H[j, c]
                d[r] = a
            s = 0.0
            for r in range(n):
                s += H[j, r] * d[r]
            f = R[j] + s                         # :359-362
            for r in range(n):                   # :364-366
                g[r] = d[r] / f
            for r in range(n):                   # :368-372
                for c in range(n):
                    P[r, c] += -g[r] * g[c] * f
            for r in range(n):                   # :374-375
                xp[r] += g[r] * v
            sigma += v * v / f                   # :377
            detf += math.log(f)                  # :378
        if cnt > 0:
            if nsig >= warmup:
                sum_sigma += sigma
                sum_detf += detf
            nsig += 1
        if t >= warmup:
            nobs += cnt
        for r in range(n):
            x[r] = xp[r]
    return nobs * math.log(2.0 * math.pi) + sum_detf + sum_sigma, nobs


@njit(cache=True, parallel=True, fastmath=False)
def _loglik_batch(y, A, Q, H, R, warmup, mle, nobs):
    for b in prange(y.shape[0]):
        v, k = _loglik_one(y[b], A[b], Q[b], H[b], R[b], warmup)
        mle[b] = v
        nobs[b] = k


def filter_loglik_batch(y, A, Q, H, R, warmup=1, threads=0):
    """-2 log L and nobs of B independent models; y[B,T,m], A,Q[B,n,n], H[B,m,n], R[B,m].

    The reference evaluates models one at a time (solver.py:62); prange over the models stands for
    "as many reference processes as the box has cores".  ``threads`` > 0 caps numba's pool."""
    if not HAVE_NUMBA:
        raise RuntimeError("numba is not importable")
    y, A, Q, H, R = (np.ascontiguousarray(v, dtype=np.float64) for v in (y, A, Q, H, R))
    B = y.shape[0]
    mle = np.empty(B)
    nobs = np.empty(B, dtype=np.int64)
    if threads and threads > 0:
        numba.set_num_threads(int(min(threads, numba.config.NUMBA_NUM_THREADS)))
    _loglik_batch(y, A, Q, H, R, int(warmup), mle, nobs)
    return mle, nobs


def max_threads():
    return int(numba.config.NUMBA_NUM_THREADS) if HAVE_NUMBA else 1
