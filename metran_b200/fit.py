"""Fit many Metran-shaped models at once: batched objective, batched quasi-Newton.

The reference fits ONE model per ``Metran.solve()`` call: scipy L-BFGS-B around
``BaseSolver.objfunction`` -> ``Metran.get_mle`` with 2-point finite differences
(metran/solver.py:42-63, 222-288; metran/metran.py:605-622).  Here the same objective,
-2 log L as a function of the alpha vector of every model, is evaluated for a whole batch
per kernel launch (mkf_dfm_loglik_alpha), so the optimiser below advances all B models in
lock step: per iteration n+1 launches for the forward-difference gradients plus one launch
per line-search trial, each over all B models.  

This is host-side driver logic (numpy, vectorised over the batch); every likelihood
evaluation happens in the CUDA library.  The optimiser works in theta = log(alpha), which
keeps alpha above the reference's lower bound of 1e-5 (metran.py:446-462) without a
projected step; its iterates differ from scipy's, its optimum does not.
"""
import numpy as np

__all__ = ["fit_dfm_batch", "BatchFitResult", "REASONS"]

REASONS = ("gtol", "ftol", "maxiter", "linesearch_failed", "nonfinite_start")


class BatchFitResult(dict):
    """alpha[B,n], obj[B] (-2 log L at the optimum), nit, nfev (per model), converged[B] (the gradient
    or function-decrease criterion was met after a successful step) and reason[B]: one of REASONS."""

    __getattr__ = dict.__getitem__


def _objective(engine, y_dev, alpha, factors_dev, dt, warmup):
    mle, _, _ = engine.dfm_loglik_alpha(y_dev, alpha, factors_dev, dt=dt, warmup=warmup)
    out = mle.cpu().numpy()
    return np.where(np.isfinite(out), out, np.inf)


def fit_dfm_batch(engine, y, factors, alpha0=10.0, dt=1.0, warmup=1, maxiter=60, gtol=1e-5, ftol=1e-10,
                  fd_step=1e-6, max_backtracks=12):
    """Maximum-likelihood alpha for B models with observations y[B,T,m] and loadings factors[B,m,k].

    Parameters mirror the single-model path: initial alpha = 10 for every state
    (metran.py:439-462), dt = grid step in days (metran.py:262), warmup = 1.
    Returns a BatchFitResult.
    """
    y_dev = engine.tensor(y)
    f_dev = engine.tensor(factors)
    B, T, m = (int(s) for s in y_dev.shape)
    k = int(f_dev.shape[-1])
    n = m + k
    theta = np.full((B, n), np.log(float(alpha0))) if np.isscalar(alpha0) else np.log(np.asarray(alpha0, dtype=np.float64))
    theta = np.broadcast_to(theta, (B, n)).copy()
    nfev = 0

    def fun(th):
        nonlocal nfev
        nfev += 1
        return _objective(engine, y_dev, np.exp(th), f_dev, dt, warmup)

    def grad(th, f0):
        g = np.empty_like(th)
        for i in range(n):                      # one launch per coordinate, all models at once
            tp = th.copy()
            h = fd_step * np.maximum(1.0, np.abs(tp[:, i]))
            tp[:, i] += h
            g[:, i] = (fun(tp) - f0) / h
        return g

    f = fun(theta)
    g = grad(theta, f)
    Hinv = np.broadcast_to(np.eye(n), (B, n, n)).copy()
    active = np.isfinite(f)
    nit = np.zeros(B, dtype=np.int64)
    reason = np.where(active, REASONS.index("maxiter"), REASONS.index("nonfinite_start"))
    for _ in range(maxiter):
        if not active.any():
            break
        p = -np.einsum("bij,bj->bi", Hinv, g)
        slope = np.einsum("bi,bi->b", p, g)
        bad_dir = slope >= 0                     # not a descent direction: reset to steepest descent
        if bad_dir.any():
            Hinv[bad_dir] = np.eye(n)
            p[bad_dir] = -g[bad_dir]
            slope[bad_dir] = -np.einsum("bi,bi->b", g[bad_dir], g[bad_dir])
        # cap the step in log-alpha so exp() stays sane
        scale = np.minimum(1.0, 2.0 / np.maximum(np.abs(p).max(axis=1), 1e-300))
        step = np.where(active, scale, 0.0)
        f_new = f.copy()
        theta_new = theta.copy()
        todo = active.copy()
        for _bt in range(max_backtracks):        # Armijo backtracking, one launch per trial
            trial = theta + step[:, None] * p
            ft = fun(trial)
            ok = todo & (ft <= f + 1e-4 * step * slope)
            theta_new[ok] = trial[ok]
            f_new[ok] = ft[ok]
            todo &= ~ok
            if not todo.any():
                break
            step = np.where(todo, 0.5 * step, step)
        moved = active & ~todo
        g_new = grad(theta_new, f_new)
        s = theta_new - theta
        yk = g_new - g
        sy = np.einsum("bi,bi->b", s, yk)
        upd = moved & (sy > 1e-12)
        if upd.any():                            # BFGS update of the inverse Hessian, batched
            rho = 1.0 / sy[upd]
            su, yu, Hu = s[upd], yk[upd], Hinv[upd]
            Hy = np.einsum("bij,bj->bi", Hu, yu)
            yHy = np.einsum("bi,bi->b", yu, Hy)
            Hinv[upd] = (Hu - rho[:, None, None] * (su[:, :, None] * Hy[:, None, :] + Hy[:, :, None] * su[:, None, :])
                         + (rho ** 2 * yHy + rho)[:, None, None] * (su[:, :, None] * su[:, None, :]))
        df = f - f_new
        nit += moved
        theta, f, g = theta_new, f_new, g_new
        hit_g = moved & (np.abs(g).max(axis=1) < gtol)
        hit_f = moved & ~hit_g & (df <= ftol * np.maximum(1.0, np.abs(f)))
        failed = active & ~moved
        reason = np.where(active & hit_g, REASONS.index("gtol"), reason)
        reason = np.where(active & hit_f, REASONS.index("ftol"), reason)
        reason = np.where(failed, REASONS.index("linesearch_failed"), reason)
        active &= ~(hit_g | hit_f | failed)
    return BatchFitResult(alpha=np.exp(theta), obj=f, nit=nit, nfev=nfev, launches=nfev,
                          converged=(reason == REASONS.index("gtol")) | (reason == REASONS.index("ftol")),
                          reason=np.array(REASONS)[reason], grad=g)
