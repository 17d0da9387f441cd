"""Host-side mirror of the reference's Kalman object, backed by the CUDA library.

``SPKalmanFilter`` keeps the public surface of the reference class
(metran/kalmanfilter.py:479-778: same method names, arguments, result attributes and
error behaviour) so that ``Metran`` -- which only ever touches the attributes listed in
SURVEY.md section 8b -- runs unmodified on top of it (see install.py).  The two module
functions are the reference's plug-in points:

* ``seqkalmanfilter_cuda``  : the 9-in / 7-out ``filtermethod`` (kalmanfilter.py:236-400)
* ``kalmansmoother_cuda``   : ``kalmansmoother`` (kalmanfilter.py:403-476)

All arithmetic happens in libmetran_b200.so; this file only packs arguments.  There is
no CPU engine here: the product fails loudly without the CUDA library / a CUDA device.
"""
from logging import getLogger

import numpy as np
import torch

from . import _lib

logger = getLogger(__name__)

_ENGINE_NAMES = ("cuda", "numba", "numpy")
_engine_singleton = None


def _engine(engine=None):
    global _engine_singleton
    if engine is not None:
        return engine
    if _engine_singleton is None:
        from .batched import KalmanEngine

        _engine_singleton = KalmanEngine()
    return _engine_singleton


def set_default_engine(engine):
    """Install a KalmanEngine used by every SPKalmanFilter created afterwards."""
    global _engine_singleton
    _engine_singleton = engine


def pack_observations(values):
    """NaN-coded [T,m] array -> (observations, observation_indices, observation_count).

    Vectorised restatement of SPKalmanFilter.set_observations (kalmanfilter.py:646-674):
    a slot is observed when finite and ``value + 1e10 != 0`` (:657, :666-667);
    ``observation_indices`` is float64, ascending, zero padded (:659, :674);
    ``observations`` carries 0.0 at missing slots (:661, :673).
    """
    y = np.asarray(values, dtype=np.float64)
    finite = np.isfinite(y)
    observed = finite & ((np.where(finite, y, 0.0) + 1e10) != 0.0)
    observations = np.where(observed, y, 0.0)
    count = observed.sum(axis=1).astype(np.int64)
    T, m = y.shape
    # stable argsort puts observed columns first, in ascending column order
    order = np.argsort(~observed, axis=1, kind="stable")
    keep = np.arange(m)[None, :] < count[:, None]
    indices = np.where(keep, order, 0).astype(np.float64)
    return observations, indices, count, observed


def unpack_observations(observations, observation_indices, observation_count):
    """Inverse of pack_observations: back to the NaN-coded matrix the kernels consume.

    The kernels process the observed series of a time step in ascending index order (the
    order set_observations produces); any other order is rejected rather than silently
    reordered.
    """
    obs = np.asarray(observations, dtype=np.float64)
    idx = np.asarray(observation_indices)
    cnt = np.asarray(observation_count).astype(np.int64)
    T, m = obs.shape
    if idx.shape != (T, m) or cnt.shape != (T,):
        raise ValueError("observation_indices / observation_count do not match observations")
    if cnt.min(initial=0) < 0 or cnt.max(initial=0) > m:
        raise ValueError("observation_count out of range")
    cols = np.arange(m)[None, :]
    valid = cols < cnt[:, None]
    ii = idx.astype(np.int64)
    if np.any(valid & ((ii < 0) | (ii >= m))):
        raise ValueError("observation_indices out of range")
    asc = (ii[:, 1:] > ii[:, :-1]) | ~valid[:, 1:]
    if not np.all(asc):
        raise ValueError("observation_indices must be strictly ascending within a time step "
                         "(the order SPKalmanFilter.set_observations produces)")
    observed = np.zeros((T, m), dtype=bool)
    rows = np.broadcast_to(np.arange(T)[:, None], (T, m))
    observed[rows[valid], ii[valid]] = True
    y = np.where(observed, obs, np.nan)
    if np.any(observed & ~(np.isfinite(obs) & (obs + 1e10 != 0.0))):
        raise ValueError("an indexed observation is not finite")
    return y


def _structure_flags(A, Q):
    A = np.asarray(A)
    Q = np.asarray(Q)
    f = 0
    if np.count_nonzero(A - np.diag(np.diag(A))) == 0:
        f |= _lib.FLAG_A_DIAGONAL
    if np.count_nonzero(Q - np.diag(np.diag(Q))) == 0:
        f |= _lib.FLAG_Q_DIAGONAL
    return f


def _filter_device(engine, y_dev, A, Q, H, R, x0, P0, covariances=True):
    A, Q, H, R = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, Q, H, R))
    n = A.shape[0]
    if A.shape != (n, n) or Q.shape != (n, n) or H.shape[1] != n or R.shape != (H.shape[0],):
        raise ValueError("inconsistent state-space matrices")
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    P0 = np.ascontiguousarray(P0, dtype=np.float64)
    return engine.filter_history(y_dev, A, Q, H, R, x0=x0[None], P0=P0[None],
                                 flags=_structure_flags(A, Q), covariances=covariances)


class _LazyHistory:
    """Device-resident filter/smoother history; arrays are copied to the host on first read.

    An objective evaluation through the reference facade (Metran.get_mle, metran.py:605-622)
    only ever reads sigmas/detfs; the [T,n,n] covariances cross PCIe only when a getter asks.
    """

    def __init__(self, out):
        self.out = out
        self._host = {}

    def get(self, key):
        if key not in self._host:
            t = self.out.get(key)
            self._host[key] = None if t is None else t[0].cpu().numpy()
        return self._host[key]


def _history_to_numpy(out):
    cnt = out["cnt"][0].cpu().numpy()
    seen = cnt > 0
    sig = out["sigma"][0].cpu().numpy()
    det = out["detf"][0].cpu().numpy()
    sigmacount = int(seen.sum())
    T = cnt.shape[0]
    # reference contract: arrays of length T whose first `sigmacount` entries are valid
    sigmas = np.zeros(T)
    detfs = np.zeros(T)
    sigmas[:sigmacount] = sig[seen]
    detfs[:sigmacount] = det[seen]
    return (sigmas, detfs, sigmacount,
            out["fm"][0].cpu().numpy(), out["fc"][0].cpu().numpy(),
            out["pm"][0].cpu().numpy(), out["pc"][0].cpu().numpy())


def seqkalmanfilter_cuda(observations, transition_matrix, transition_covariance, observation_matrix,
                         observation_variance, observation_indices, observation_count,
                         filtered_state_mean, filtered_state_covariance, engine=None):
    """CUDA ``filtermethod``: same 9 inputs / 7 outputs as seqkalmanfilter
    (metran/kalmanfilter.py:243-253, 392-400).  Runs mkf_filter_history with B = 1."""
    eng = _engine(engine)
    y = unpack_observations(observations, observation_indices, observation_count)
    out = _filter_device(eng, eng.tensor(y[None]), transition_matrix, transition_covariance,
                         observation_matrix, observation_variance,
                         filtered_state_mean, filtered_state_covariance)
    return _history_to_numpy(out)


def kalmansmoother_cuda(filtered_state_means, filtered_state_covariances, predicted_state_means,
                        predicted_state_covariances, transition_matrix, engine=None):
    """CUDA ``kalmansmoother`` (metran/kalmanfilter.py:403-476): returns (means, covariances)."""
    eng = _engine(engine)
    A = np.ascontiguousarray(transition_matrix, dtype=np.float64)
    flags = _lib.FLAG_A_DIAGONAL if np.count_nonzero(A - np.diag(np.diag(A))) == 0 else 0
    sm, sc, status = eng.smoother(np.asarray(filtered_state_means)[None], np.asarray(filtered_state_covariances)[None],
                                  np.asarray(predicted_state_means)[None], np.asarray(predicted_state_covariances)[None],
                                  A, flags=flags)
    if int(status[0]) & _lib.STATUS_CHOL_FAILED:
        logger.warning("Kalman smoother: a predicted covariance is not positive definite.")
    return sm[0].cpu().numpy(), sc[0].cpu().numpy()


def torch_int32(eng, a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(eng.device)


class SPKalmanFilter:
    """Kalman filter class for Metran, CUDA engine (mirror of kalmanfilter.py:479-778).

    Parameters
    ----------
    engine : str, optional
        "cuda" (default).  "numba" / "numpy" are accepted for call compatibility with
        ``Metran.solve(engine=...)`` (metran.py:991, 1025) and run the same CUDA kernels;
        this package has no CPU engine.
    """

    # use the compacted time grid when at most this fraction of the steps carries an observation
    COMPACT_BELOW = 0.5

    def __init__(self, engine="cuda", device_engine=None):
        self._compact = None
        self._arrays = {}
        self._lazy = None
        self._pending = None
        self._mle_w1 = None
        self._hist = None
        self.init_states()
        self.detfs = None
        self.sigmas = None
        self.nobs = None
        self.mask = False
        self.engine = engine
        self._eng = device_engine
        self._y_dev = None
        self._hist = None
        self.filtermethod = self._filtermethod

    # ---------------------------------------------------------------- reference surface
    def init_states(self):
        """Reset state means and covariances (kalmanfilter.py:506-518)."""
        # the reference resets the six state arrays only: sigmas / detfs (and with them get_mle())
        # survive.  A not yet materialised history therefore stays pending for sigmas / detfs,
        # while the state arrays read as None until the next run.
        self._arrays = {k: v for k, v in getattr(self, "_arrays", {}).items() if k in ("sigmas_", "detfs_")}
        self._lazy = None
        self._states_reset = True
        self.filtered_state_means = None
        self.filtered_state_covariances = None
        self.predicted_state_means = None
        self.predicted_state_covariances = None
        self.smoothed_state_means = None
        self.smoothed_state_covariances = None
        self._hist = None

    # The six result arrays of the reference are plain attributes there; here they are
    # properties backed by the device-resident history so that an objective evaluation does
    # not pay for copying [T,n,n] covariances nobody reads (they are fetched on first access).
    def _get_array(self, name, key):
        if name in self._arrays:
            return self._arrays[name]
        if getattr(self, "_states_reset", False):
            return None
        if key in ("fm", "fc", "pm", "pc"):
            self._ensure_history()
        if self._lazy is not None and self._lazy.out.get(key) is not None:
            self._arrays[name] = self._lazy.get(key)
            return self._arrays[name]
        return None

    def _set_array(self, name, value):
        self._arrays[name] = value

    def _ensure_history(self):
        """Run the history kernel for the run_filter() call still pending, on the SNAPSHOT of the
        matrices, initial state and observations taken by that call."""
        if self._hist is None and getattr(self, "_pending", None) is not None:
            out = self._filter_on_device(self._pending)
            cnt = out["cnt"][0].cpu().numpy()
            seen = cnt > 0
            self._arrays.setdefault("sigmas_", out["sigma"][0].cpu().numpy()[seen])
            self._arrays.setdefault("detfs_", out["detf"][0].cpu().numpy()[seen])
            if not getattr(self, "_states_reset", False):
                self._lazy = _LazyHistory(out)

    def _get_series(self, name):
        if name not in self._arrays:
            self._ensure_history()
        return self._arrays.get(name)

    sigmas = property(lambda self: self._get_series("sigmas_"), lambda self, v: self._set_array("sigmas_", v))
    detfs = property(lambda self: self._get_series("detfs_"), lambda self, v: self._set_array("detfs_", v))

    filtered_state_means = property(lambda self: self._get_array("fm_", "fm"),
                                    lambda self, v: self._set_array("fm_", v))
    filtered_state_covariances = property(lambda self: self._get_array("fc_", "fc"),
                                          lambda self, v: self._set_array("fc_", v))
    predicted_state_means = property(lambda self: self._get_array("pm_", "pm"),
                                     lambda self, v: self._set_array("pm_", v))
    predicted_state_covariances = property(lambda self: self._get_array("pc_", "pc"),
                                           lambda self, v: self._set_array("pc_", v))
    smoothed_state_means = property(lambda self: self._get_array("sm_", "sm"),
                                    lambda self, v: self._set_array("sm_", v))
    smoothed_state_covariances = property(lambda self: self._get_array("sc_", "sc"),
                                          lambda self, v: self._set_array("sc_", v))

    def set_matrices(self, transition_matrix, transition_covariance, observation_matrix, observation_variance):
        """Set the state-space matrices (kalmanfilter.py:520-548)."""
        self.transition_matrix = transition_matrix
        self.transition_covariance = transition_covariance
        self.observation_matrix = observation_matrix
        self.observation_variance = observation_variance
        self.nstate = np.int64(self.transition_matrix.shape[0])

    def get_mle(self, warmup=1):
        """-2 log L (kalmanfilter.py:550-567, quirk Q4): the value the log-likelihood kernel
        produced in run_filter when warmup is the default, else from sigmas/detfs."""
        if warmup == 1 and getattr(self, "_mle_w1", None) is not None:
            return self._mle_w1
        detfs = self.detfs[warmup:]
        sigmas = self.sigmas[warmup:]
        nobs = np.sum(self.observation_count[warmup:])
        return nobs * np.log(2 * np.pi) + np.sum(detfs) + np.sum(sigmas)

    def set_observations(self, oseries):
        """Pack observations (kalmanfilter.py:646-674) and upload them once.

        ``oseries``: pandas.DataFrame (or 2-D array), NaN = missing.  Besides the
        reference's three host arrays the NaN-coded matrix is kept resident on the device,
        so objective evaluations and mask toggles do not re-upload it.
        """
        self.oseries_index = getattr(oseries, "index", None)
        values = np.asarray(oseries, dtype=np.float64)
        if values.ndim != 2:
            raise ValueError("oseries must be two-dimensional")
        (self.observations, self.observation_indices, self.observation_count,
         observed) = pack_observations(values)
        y = np.where(observed, values, np.nan)
        eng = self._device_engine()
        self._y_dev = eng.tensor(y[None])
        # Compacted grid for the log-likelihood kernel: on Metran's daily grid most steps carry no
        # observation at all; when fewer than COMPACT_BELOW of them do, objective evaluations walk
        # the observed steps only (mkf_dfm_loglik_compact).  Host-side, once per data set.
        self._compact = None
        seen = np.flatnonzero(observed.any(axis=1))
        if 0 < seen.size <= self.COMPACT_BELOW * values.shape[0]:
            self._compact = (eng.tensor(y[seen]), torch_int32(eng, seen))
        # results of a filter run on the previous observations must not be served for these
        self._hist = None
        self._pending = None
        self._mle_w1 = None
        self._lazy = None

    def run_filter(self, initial_state_mean=None, initial_state_covariance=None, engine=None):
        """Run the Kalman filter (kalmanfilter.py:696-778) on the device."""
        if self.mask:
            logger.info("Running Kalman filter with masked observations.")
        if engine is not None:
            if engine not in _ENGINE_NAMES:
                msg = "Unknown engine defined in run_filter."
                logger.error(msg)
                raise Exception(msg)
            self.engine = engine
        if initial_state_mean is None:
            initial_state_mean = np.zeros(self.nstate)
        if initial_state_covariance is None:
            initial_state_covariance = np.eye(self.nstate)

        self._arrays = {}
        self._lazy = None
        self._hist = None
        self._mle_w1 = None
        self._pending = None
        self._states_reset = False
        if self.filtermethod == self._filtermethod:
            # Device path.  What an objective evaluation needs (-2 log L, warmup = 1) is
            # computed now by the log-likelihood kernel; the history kernel that fills
            # sigmas/detfs and the state arrays runs only if somebody reads them -- on the
            # snapshot taken here, whatever set_matrices()/set_observations() do in between.
            self._pending = self._snapshot(initial_state_mean, initial_state_covariance)
            self._mle_w1 = self._loglik_on_device(self._pending)
            return
        (sigmas, detfs, sigmacount, filtered_state_means, filtered_state_covariances,
         predicted_state_means, predicted_state_covariances) = self.filtermethod(
            self.observations, self.transition_matrix, self.transition_covariance,
            self.observation_matrix, self.observation_variance, self.observation_indices,
            self.observation_count, initial_state_mean, initial_state_covariance)

        self.sigmas = sigmas[:sigmacount]
        self.detfs = detfs[:sigmacount]
        self.filtered_state_means = filtered_state_means
        self.filtered_state_covariances = filtered_state_covariances
        self.predicted_state_means = predicted_state_means
        self.predicted_state_covariances = predicted_state_covariances

    def run_smoother(self):
        """Filter, then RTS smoother (kalmanfilter.py:676-694); history stays on the device
        between the two kernels."""
        self.run_filter()
        self._ensure_history()
        eng = self._device_engine()
        h = self._hist
        A = np.ascontiguousarray(self.transition_matrix, dtype=np.float64)
        if h is None:
            # a plug-in filtermethod (kalmanfilter.py:761-771) left host arrays only: smooth those,
            # as the reference's run_smoother does with whatever filter produced them (:685-691)
            sm, sc = kalmansmoother_cuda(self.filtered_state_means, self.filtered_state_covariances,
                                         self.predicted_state_means, self.predicted_state_covariances,
                                         A, engine=eng)
            self.smoothed_state_means = sm
            self.smoothed_state_covariances = sc
            return
        flags = h["flags"] & _lib.FLAG_A_DIAGONAL
        # the projections every consumer of the smoother asks for (simulate(), kalmanfilter.py:596-602,
        # with a row-scaled copy of the observation matrix: metran.py:758-829) are fused into the
        # smoother's epilogue: H x_s and diag(H P_s H^T) leave the kernel, [T,n,n] is not read back
        Hobs = np.ascontiguousarray(self.observation_matrix, dtype=np.float64)
        res = eng.smoother_project(h["fm"], h["fc"], h["pm"], h["pc"], A, Hobs, flags=flags)
        sm, sc, status = res["sm"], res["sc"], res["status"]
        if int(status[0]) & _lib.STATUS_CHOL_FAILED:
            logger.warning("Kalman smoother: a predicted covariance is not positive definite.")
        self._hist["sm"], self._hist["sc"] = sm, sc
        self._hist["proj"] = (Hobs, res["proj_mean"], res["proj_var"])
        self._arrays.pop("sm_", None)
        self._arrays.pop("sc_", None)
        if self._lazy is None:
            self._lazy = _LazyHistory(self._hist)
        self._lazy.out = self._hist

    def simulate(self, observation_matrix, method="smoother"):
        """Projected means and variances (kalmanfilter.py:569-603), one kernel for all t."""
        Hs = np.ascontiguousarray(observation_matrix, dtype=np.float64)
        fused = self._fused_projection(Hs) if method != "filter" else None
        if fused is not None:
            return fused
        means, covs = self._moments(method)
        eng = self._device_engine()
        mu, var = eng.project(means, covs, Hs)
        return list(mu[0].cpu().numpy()), list(var[0].cpu().numpy())

    def _fused_projection(self, Hs):
        """simulate() from the smoother's fused epilogue when Hs is the observation matrix with its
        rows scaled (what Metran passes: H * oseries_std, metran.py:775-829): H_s x = s * (H x) and
        diag(H_s P H_s^T) = s^2 * diag(H P H^T)."""
        h = self._hist
        if h is None or "proj" not in h:
            return None
        Hobs, pmean, pvar = h["proj"]
        if Hs.shape != Hobs.shape:
            return None
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(Hobs != 0.0, Hs / Hobs, np.nan)
        if np.any((Hobs == 0.0) & (Hs != 0.0)) or np.all(np.isnan(ratio), axis=1).any():
            return None
        scale = np.nanmax(ratio, axis=1)
        if not np.array_equal(Hobs * scale[:, None], Hs):
            return None
        mu = pmean[0].cpu().numpy() * scale[None, :]
        var = pvar[0].cpu().numpy() * (scale * scale)[None, :]
        return list(mu), list(var)

    def decompose(self, observation_matrix, method="smoother"):
        """Specific / common dynamic factor contributions (kalmanfilter.py:605-644)."""
        if method == "filter":
            means = self.filtered_state_means
        else:
            means = self.smoothed_state_means
        Hs = np.asarray(observation_matrix, dtype=np.float64)
        nsdf = self.observation_matrix.shape[0]
        ncdf = self.observation_matrix.shape[1] - nsdf
        eng = self._device_engine()
        # one projection launch for all 1 + ncdf components: the projection matrices (specific block only;
        # one common factor each) are stacked row-wise and the result is split again
        p = Hs.shape[0]
        stack = np.zeros(((1 + ncdf) * p, Hs.shape[1]))
        stack[:p, :nsdf] = Hs[:, :nsdf]
        for k in range(ncdf):
            stack[(1 + k) * p:(2 + k) * p, nsdf + k] = Hs[:, nsdf + k]
        if stack.shape[0] <= 256:
            allc, _ = eng.project(means[None], None, stack, variances=False)
            allc = allc[0].cpu().numpy()                     # [T, (1 + ncdf) * p]
            parts = [allc[:, i * p:(i + 1) * p] for i in range(1 + ncdf)]
        else:                                                # beyond mkf_project's width: one launch per component
            parts = [eng.project(means[None], None, stack[i * p:(i + 1) * p], variances=False)[0][0].cpu().numpy()
                     for i in range(1 + ncdf)]
        sdf_means = list(parts[0])
        cdf_means = [list(c) for c in parts[1:]]
        return (sdf_means, cdf_means)

    # ---------------------------------------------------------------------- internals
    def _device_engine(self):
        if self._eng is None:
            self._eng = _engine()
        return self._eng

    def _moments(self, method):
        h = self._hist
        if method == "filter":
            if h is not None:
                return h["fm"], h["fc"]
            return self.filtered_state_means[None], self.filtered_state_covariances[None]
        if h is not None and "sm" in h:
            return h["sm"], h["sc"]
        return self.smoothed_state_means[None], self.smoothed_state_covariances[None]

    def _snapshot(self, x0, P0):
        """Everything one filter run depends on, copied (host arrays) or referenced (the device
        observations: set_observations() replaces the tensor, it never writes into it)."""
        eng = self._device_engine()
        if self._y_dev is None:
            self._y_dev = eng.tensor(unpack_observations(self.observations, self.observation_indices,
                                                         self.observation_count)[None])
        return dict(x0=np.array(x0, dtype=np.float64), P0=np.array(P0, dtype=np.float64),
                    A=np.array(self.transition_matrix, dtype=np.float64),
                    Q=np.array(self.transition_covariance, dtype=np.float64),
                    H=np.array(self.observation_matrix, dtype=np.float64),
                    R=np.array(self.observation_variance, dtype=np.float64), y=self._y_dev,
                    compact=self._compact)

    def _loglik_on_device(self, snap):
        """-2 log L (warmup = 1) of the snapshot: thread-per-model DFM kernel when the model is
        Metran-shaped (metran.py:283-384) with the default initial state, else the generic
        kernel."""
        eng = self._device_engine()
        A, Q, H, R, x0, P0 = (snap[k] for k in ("A", "Q", "H", "R", "x0", "P0"))
        m, n = H.shape
        flags = _structure_flags(A, Q)
        dfm = (flags == (_lib.FLAG_A_DIAGONAL | _lib.FLAG_Q_DIAGONAL) and m <= n
               and eng.dfm_supported(m, n - m) and np.array_equal(H[:, :m], np.eye(m))
               and not np.any(x0) and np.array_equal(P0, np.eye(n)))
        if dfm:
            # one upload for all parameters of the evaluation (phi, q, loadings, r), result buffers reused:
            # a single model's objective call is bound by host overhead, not by the kernel
            k = n - m
            buf = np.concatenate([np.diag(A), np.diag(Q), np.ascontiguousarray(H[:, m:]).ravel(), R])
            par = eng.tensor(buf)
            phi_d, q_d = par[:n][None], par[n:2 * n][None]
            g_d = par[2 * n:2 * n + m * k].view(1, m, k)
            r_d = par[2 * n + m * k:][None] if np.any(R) else None
            if getattr(self, "_out1", None) is None or self._out1[0].device != par.device:
                self._out1 = (eng.empty(1), eng.empty(1, dtype=torch.int64), eng.empty(1, dtype=torch.int32))
        if dfm and snap.get("compact") is not None:
            yc, tidx = snap["compact"]
            mle, _, _ = eng.dfm_loglik_compact(yc, tidx, g_d, phi=phi_d, q=q_d, r=r_d, out=self._out1)
        elif dfm:
            mle, _, _ = eng.dfm_loglik(snap["y"], phi_d, q_d, g_d, r=r_d, out=self._out1)
        else:
            mle, _, _ = eng.filter_loglik(snap["y"], A, Q, H, R, x0=x0[None], P0=P0[None], flags=flags)
        return float(mle[0])

    def loglik_batch(self, A, Q, H, R, warmup=1):
        """-2 log L of the resident observations for a BATCH of matrix sets A,Q[B,n,n], H[B,m,n],
        R[B,m] in one launch (default initial state): the thread-per-model DFM kernel when every
        set is Metran-shaped, else the generic kernel.  Used by the batched Hessian."""
        eng = self._device_engine()
        if self._y_dev is None:
            self._y_dev = eng.tensor(unpack_observations(self.observations, self.observation_indices,
                                                         self.observation_count)[None])
        A, Q, H, R = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, Q, H, R))
        B, m, n = H.shape
        y = self._y_dev[0]
        eye = np.eye(n, dtype=bool)
        diag = not np.any(A[:, ~eye]) and not np.any(Q[:, ~eye])
        dfm = (diag and m <= n and eng.dfm_supported(m, n - m)
               and np.array_equal(H[:, :, :m], np.broadcast_to(np.eye(m), (B, m, m))))
        if dfm:
            phi = np.ascontiguousarray(np.diagonal(A, axis1=1, axis2=2))
            q = np.ascontiguousarray(np.diagonal(Q, axis1=1, axis2=2))
            if self._compact is not None:
                mle, _, _ = eng.dfm_loglik_compact(self._compact[0], self._compact[1], np.ascontiguousarray(H[:, :, m:]),
                                                   phi=phi, q=q, r=R if np.any(R) else None, warmup=warmup)
            else:
                mle, _, _ = eng.dfm_loglik(y, phi, q, np.ascontiguousarray(H[:, :, m:]),
                                           r=R if np.any(R) else None, warmup=warmup)
        else:
            flags = (_lib.FLAG_A_DIAGONAL | _lib.FLAG_Q_DIAGONAL) if diag else 0
            mle, _, _ = eng.filter_loglik(y, A, Q, H, R, warmup=warmup, flags=flags)
        return mle.cpu().numpy()

    def _filter_on_device(self, snap):
        eng = self._device_engine()
        out = _filter_device(eng, snap["y"], snap["A"], snap["Q"], snap["H"], snap["R"], snap["x0"], snap["P0"])
        self._hist = out
        return out

    def _filtermethod(self, observations, A, Q, H, R, observation_indices, observation_count, x0, P0):
        """Bound CUDA filtermethod with the reference's 9-in / 7-out contract: uses the
        device-resident observations when the host arrays are the ones set_observations made."""
        eng = self._device_engine()
        if (self._y_dev is not None and observations is self.observations
                and observation_indices is self.observation_indices
                and observation_count is self.observation_count):
            y_dev = self._y_dev
        else:
            y_dev = eng.tensor(unpack_observations(observations, observation_indices, observation_count)[None])
        out = _filter_device(eng, y_dev, A, Q, H, R, x0, P0)
        self._hist = out
        return _history_to_numpy(out)
