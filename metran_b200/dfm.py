"""Callers of the Kalman hot path, mirrored from the reference and batched.

Only the rows of SURVEY.md section 8 are mirrored here, not the ``Metran`` facade:

* ``_phi`` / matrix builders / ``_get_matrices``  <- metran/metran.py:246-416      (row a10)
* ``get_mle``                                       <- metran/metran.py:605-622      (row a9)
* ``objfunction``                                   <- metran/solver.py:42-63        (row a11)
* ``_run_kalman`` + the state/projection getters    <- metran/metran.py:655-989      (row a12)
* ``get_mle_batch`` / ``BatchedObjective`` / ``solve`` are the batched-objective hook of
  SURVEY.md section 8f rank 1: all finite-difference points scipy's L-BFGS-B asks for
  (metran/solver.py:248-255 passes no ``jac``) are evaluated by ONE kernel launch.

The pandas-facing ``Metran`` class of the reference stays what it is; to run it on this
engine use ``metran_b200.install()``.  ``DynamicFactorModel`` is the array-level entry
for users who already have standardized series and factor loadings (both produced by the
reference's out-of-scope ``Metran.__init__`` / ``FactorAnalysis``).
"""
from logging import getLogger

import numpy as np

from .kalmanfilter import SPKalmanFilter, _engine

logger = getLogger(__name__)


class BatchedObjective:
    """Map-like for ``scipy.optimize.minimize(..., options={"workers": obj})``.

    scipy calls ``workers(fun, iterable_of_x)`` once per finite-difference gradient
    (scipy/optimize/_numdiff.py, _dense_difference); instead of mapping ``fun`` over the
    points this evaluates all of them in one batched kernel launch.
    """

    def __init__(self, model):
        self.model = model
        self.launches = 0
        self.points = 0

    def __call__(self, fun, xs):
        xs = [np.asarray(x, dtype=np.float64) for x in xs]
        if not xs:
            return []
        out = self.model.get_mle_batch(np.stack(xs))
        self.launches += 1
        self.points += len(xs)
        return list(out)


def near_psd(A, epsilon=0):
    """Nearest positive semi-definite matrix, the construction of BaseSolver._nearPSD
    (solver.py:167-191): eigenvalues clipped at epsilon, rows rescaled to the original diagonal."""
    n = A.shape[0]
    eigval, eigvec = np.linalg.eig(A)
    val = np.array(np.maximum(eigval, epsilon))
    vec = np.array(eigvec)
    with np.errstate(divide="ignore", invalid="ignore"):
        T = 1 / (vec ** 2 @ val.T)
        T = np.array(np.sqrt(np.diag(np.array(T).reshape((n)))))
        B = T @ vec * np.diag(np.array(np.sqrt(val)).reshape((n)))
    return B @ B.T


def batched_covariance(fun_batch, x0, epsilon=None, diff="forward", max_rounds=8):
    """Covariance of the parameter estimates from a finite-difference Hessian, batched.

    Mirror of BaseSolver._get_covariance (solver.py:65-140): the Hessian column j is the
    difference of two forward-difference gradients (scipy ``approx_fprime`` with absolute step
    epsilon) taken at x0 and at x0 + d_j e_j, d_j = epsilon * max(|x0_j|, 0.1) (:106-123); the
    covariance is pinv(Hessian), repaired through the nearest PSD matrix when a variance is not
    positive (:126-136), and epsilon grows tenfold until every variance is positive (:138).

    The reference evaluates the (n+1)^2 (forward) or (2n+1)(n+1) (central) objective values one
    at a time.  Here scipy's own ``approx_fprime`` is first run against a recording stand-in to
    collect exactly the points it would evaluate, ``fun_batch(points[npoints, n])`` evaluates
    them in ONE launch, and ``approx_fprime`` is run again on the cached values -- same points,
    same difference formulas, same result as the reference's loop.
    """
    from scipy.optimize import approx_fprime

    x0 = np.array(x0, dtype=np.float64)
    n = x0.shape[0]
    if epsilon is None:
        epsilon = np.finfo(float).eps ** (1.0 / 4)

    def gradient_sites():
        sites = [x0.copy()] if diff == "forward" else []
        for j in range(n):
            d = epsilon * max(np.abs(x0[j]), 0.1)
            steps = (d,) if diff == "forward" else (d, -d)
            for step in steps:
                xx = x0.copy()
                xx[j] = x0[j] + step
                sites.append(xx)
        return sites

    cov = None
    for _ in range(max_rounds):
        sites = gradient_sites()
        recorded = []

        def record(x):
            recorded.append(np.array(x, dtype=np.float64))
            return 0.0

        for site in sites:
            approx_fprime(site, record, epsilon)
        points = np.stack(recorded)
        uniq, inverse = np.unique(points, axis=0, return_inverse=True)
        values = np.asarray(fun_batch(uniq), dtype=np.float64)
        cache = {uniq[i].tobytes(): values[i] for i in range(uniq.shape[0])}

        def replay(x):
            return cache[np.array(x, dtype=np.float64).tobytes()]

        grads = [approx_fprime(site, replay, epsilon) for site in sites]
        hessian = np.zeros((n, n))
        pos = 0
        f0 = None
        if diff == "forward":
            f0 = grads[0]
            pos = 1
        for j in range(n):
            d = epsilon * max(np.abs(x0[j]), 0.1)
            if diff == "forward":
                ff = grads[pos]
                pos += 1
                hessian[: j + 1, j] = (ff[: j + 1] - f0[: j + 1]) / d
            else:
                ff, fb = grads[pos], grads[pos + 1]
                pos += 2
                hessian[: j + 1, j] = (ff[: j + 1] - fb[: j + 1]) / (2 * d)
            hessian[j, : j + 1] = hessian[: j + 1, j]
        if not np.isnan(hessian).any():
            cov = np.linalg.pinv(hessian)
            if np.amin(np.diag(cov)) <= 0:
                try:
                    cov = np.linalg.pinv(near_psd(hessian))
                except Exception as e:           # as the reference: keep going with a larger step
                    logger.debug("Could not calculate 'cov': %s", e)
            if np.amin(np.diag(cov)) > 0:
                return cov
        epsilon *= 10.0
    return cov


class DynamicFactorModel:
    """Metran's dynamic factor model at array level, evaluated on the GPU.

    Parameters
    ----------
    oseries : array_like or pandas.DataFrame, shape [T, nseries]
        Standardized series on the regular (daily) grid, NaN = missing -- what
        ``Metran.set_observations`` leaves in ``mt.oseries`` (metran.py:571-574).
    factors : array_like [nseries, nfactors]
        Factor loadings (``mt.factors``; metran.py:199-226).
    dt : float
        Grid step in days divided by one day: ``a`` of ``Metran._phi`` (metran.py:262).
    """

    def __init__(self, oseries, factors, dt=1.0, engine=None, warmup=1):
        self.oseries_index = getattr(oseries, "index", None)
        self.oseries = np.asarray(oseries, dtype=np.float64)
        self.factors = np.asarray(factors, dtype=np.float64)
        if self.factors.ndim != 2 or self.factors.shape[0] != self.oseries.shape[1]:
            raise ValueError("factors must be [nseries, nfactors]")
        self.nseries, self.nfactors = self.factors.shape
        self.nstate = self.nseries + self.nfactors
        self.dt = float(dt)
        self.warmup = int(warmup)
        self._eng = engine
        self.kf = SPKalmanFilter(engine="cuda", device_engine=self.engine)
        self.kf.set_observations(oseries)
        self.parameters = None
        self.fit = None
        self.nfev = 0

    @property
    def engine(self):
        if self._eng is None:
            self._eng = _engine()
        return self._eng

    # ------------------------------------------------------- matrices (metran.py:246-416)
    def _phi(self, alpha):
        """phi = exp(-dt / alpha) (metran.py:262-263)."""
        return np.exp(-self.dt / np.asarray(alpha, dtype=np.float64))

    def get_transition_matrix(self, p):
        """diag(phi), specific factors first, then common ones (metran.py:283-290)."""
        return np.diag(self._phi(p))

    def get_transition_covariance(self, p):
        """diag((1-phi^2)(1-communality)) / diag(1-phi^2) (metran.py:310-322)."""
        phi = self._phi(p)
        comm = np.sum(np.square(self.factors), axis=1)
        q = np.empty(self.nstate)
        q[: self.nseries] = (1 - phi[: self.nseries] ** 2) * (1 - comm)
        q[self.nseries:] = 1 - phi[self.nseries:] ** 2
        return np.diag(q)

    def get_observation_matrix(self):
        """[I | factors] (metran.py:365-370)."""
        H = np.zeros((self.nseries, self.nstate))
        H[:, : self.nseries] = np.eye(self.nseries)
        H[:, self.nseries:] = self.factors
        return H

    def get_observation_variance(self):
        """Zero observation variance (metran.py:382-384)."""
        return np.zeros(self.nseries)

    def _get_matrices(self, p):
        return (self.get_transition_matrix(p), self.get_transition_covariance(p),
                self.get_observation_matrix(), self.get_observation_variance())

    # --------------------------------------------------------------- objective (a9, a11)
    def _use_dfm_kernel(self):
        return self.engine.dfm_supported(self.nseries, self.nfactors) and self.kf._y_dev is not None

    def get_mle_batch(self, P):
        """-2 log L for every row of P[npoints, nstate] in one launch.

        Thread-per-model DFM kernel with the alpha -> (phi, q) map fused
        (mkf_dfm_loglik_alpha) when the shape is compiled, else the generic kernel on
        host-built matrices.  The observations stay resident on the device and are shared
        by all points (stride 0).
        """
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        if P.shape[1] != self.nstate:
            raise ValueError("parameter vectors must have %d entries" % self.nstate)
        eng = self.engine
        y = self.kf._y_dev[0]
        if self._use_dfm_kernel() and self.kf._compact is not None:
            # sparse daily grid: observed steps only, k-step predicts in closed form
            mle, _, _ = eng.dfm_loglik_compact(self.kf._compact[0], self.kf._compact[1], self.factors, alpha=P,
                                               dt=self.dt, warmup=self.warmup)
        elif self._use_dfm_kernel():
            mle, _, _ = eng.dfm_loglik_alpha(y, P, self.factors, dt=self.dt, warmup=self.warmup)
        else:
            mats = [self._get_matrices(p) for p in P]
            A, Q, H, R = (np.stack([m[i] for m in mats]) for i in range(4))
            mle, _, _ = eng.filter_loglik(y, A, Q, H, R, warmup=self.warmup, flags=3)
        self.nfev += P.shape[0]
        return mle.cpu().numpy()

    def get_mle(self, p):
        """-2 log L at parameters p (metran.py:605-622)."""
        return float(self.get_mle_batch(np.asarray(p, dtype=np.float64)[None])[0])

    def objfunction(self, p, callback=None):
        """Objective used by the solver (solver.py:42-63)."""
        if callback is not None:
            p = callback(p)
        return self.get_mle(p)

    def solve(self, p0=None, pmin=1e-5, pmax=None, method="l-bfgs-b", batched=True, **kwargs):
        """Minimise -2 log L like ScipySolve.solve (solver.py:222-288).

        Initial alpha = 10, bounds (1e-5, None) as Metran.set_init_parameters
        (metran.py:439-462).  With ``batched`` the finite-difference gradient points are
        evaluated in one launch per iteration through scipy's ``workers`` hook; the
        optimiser trajectory is scipy's own either way.
        """
        from scipy.optimize import minimize

        x0 = np.full(self.nstate, 10.0) if p0 is None else np.asarray(p0, dtype=np.float64)
        bounds = [(pmin, pmax)] * self.nstate
        options = dict(kwargs.pop("options", {}) or {})
        hook = None
        if batched:
            hook = BatchedObjective(self)
            options["workers"] = hook
        self.nfev = 0
        result = minimize(fun=self.objfunction, method=method, x0=x0, bounds=bounds,
                          options=options or None, **kwargs)
        stderr = np.full(self.nstate, np.nan)
        pcov = None
        if hasattr(result, "hess_inv"):
            pcov = result.hess_inv.todense()
            stderr = np.sqrt(np.diag(pcov))
        if np.isnan(stderr).any():
            # solver.py:264-269: finite-difference Hessian; all its points in one launch
            pcov = self.get_covariance(result.x)
            stderr = np.sqrt(np.diag(pcov))
        self.fit = result
        self.parameters = {"optimal": np.array(result.x), "stderr": stderr, "pcov": pcov,
                           "obj_func": float(result.fun), "nfev": int(result.nfev),
                           "aic": 2 * self.nstate + float(result.fun),
                           "batched_launches": None if hook is None else hook.launches}
        return bool(getattr(result, "success", True)), np.array(result.x), stderr

    def get_covariance(self, p, epsilon=None, diff="forward"):
        """Parameter covariance from the finite-difference Hessian of -2 log L at p
        (BaseSolver._get_covariance, solver.py:65-140), every objective value of one Hessian
        from a single batched launch."""
        return batched_covariance(self.get_mle_batch, p, epsilon=epsilon, diff=diff)

    # ------------------------------------------------------------------ getters (a12)
    def _run_kalman(self, method, p):
        """(Re)run filter or smoother at parameters p (metran.py:963-989, always explicit p)."""
        self.kf.set_matrices(*self._get_matrices(p))
        if method == "filter":
            self.kf.run_filter()
        else:
            self.kf.run_smoother()

    def get_state_means(self, p, method="smoother"):
        """Filtered / smoothed state means [T, nstate] (metran.py:655-680)."""
        self._run_kalman(method, p)
        return self.kf.filtered_state_means if method == "filter" else self.kf.smoothed_state_means

    def get_state_variances(self, p, method="smoother"):
        """diag of the filtered / smoothed covariances [T, nstate] (metran.py:682-713)."""
        self._run_kalman(method, p)
        cov = self.kf.filtered_state_covariances if method == "filter" else self.kf.smoothed_state_covariances
        return np.diagonal(cov, axis1=1, axis2=2).copy()

    def get_simulated_means(self, p, observation_matrix=None, method="smoother"):
        """Projected means H x_t [T, nseries] (metran.py:758-790; scaling is the caller's)."""
        self._run_kalman(method, p)
        H = self.get_observation_matrix() if observation_matrix is None else observation_matrix
        means, _ = self.kf.simulate(H, method=method)
        return np.asarray(means)

    def get_simulated_variances(self, p, observation_matrix=None, method="smoother"):
        """max(diag(H P_t H^T), 0) [T, nseries] (metran.py:792-829)."""
        self._run_kalman(method, p)
        H = self.get_observation_matrix() if observation_matrix is None else observation_matrix
        _, var = self.kf.simulate(H, method=method)
        return np.asarray(var)

    def decompose_simulation(self, p, observation_matrix=None, method="smoother"):
        """(sdf[T, nseries], [cdf_k[T, nseries]]) (metran.py:885-942)."""
        self._run_kalman(method, p)
        H = self.get_observation_matrix() if observation_matrix is None else observation_matrix
        sdf, cdf = self.kf.decompose(H, method=method)
        return np.asarray(sdf), [np.asarray(c) for c in cdf]

    def mask_observations(self, mask):
        """Hide observations from the filter (metran.py:464-495)."""
        mask = np.asarray(mask).astype(bool)
        if mask.shape != self.oseries.shape:
            logger.error("Dimensions of mask %s do not equal dimensions of series %s. "
                         "Mask cannot be applied.", mask.shape, self.oseries.shape)
            return
        self.kf.init_states()
        self.kf.set_observations(np.where(mask, np.nan, self.oseries))
        self.kf.mask = True

    def unmask_observations(self):
        """Restore the observations (metran.py:497-507)."""
        self.kf.init_states()
        self.kf.set_observations(self.oseries)
        self.kf.mask = False
