"""A HOST stand-in for libmetran_b200.so, for CPU-only tests of the host-side logic.

TEST INFRASTRUCTURE ONLY.  It implements the C ABI of include/metran_b200.h on host
pointers by calling the oracle (oracle/), so that argument packing, shapes, strides,
compaction, sharding and the Metran-facing mirrors can be exercised by
``pytest -m "not gpu"`` on a box without a GPU.  The product never imports this module and
never selects a host path by itself: ``KalmanEngine()`` without arguments requires the CUDA
library and a CUDA device.
"""
import ctypes

import numpy as np

from oracle import kalman_oracle as ko


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return int(p)


def _arr(p, shape, dtype=np.float64):
    a = _addr(p)
    if a == 0:
        return None
    n = int(np.prod(shape)) if len(shape) else 1
    ct = {np.float64: ctypes.c_double, np.int64: ctypes.c_int64, np.int32: ctypes.c_int32}[dtype]
    buf = (ct * n).from_address(a)
    return np.ctypeslib.as_array(buf).reshape(shape)


def _model(p, b, stride, shape):
    a = _addr(p)
    if a == 0:
        return None
    return _arr(a + 8 * b * stride, shape)


class FakeLib:
    def __init__(self):
        self.calls = []

    # ---- misc ---------------------------------------------------------------------
    def mkf_version(self):
        return 100

    def mkf_error_string(self, code):
        return {0: b"ok", 1: b"bad argument", 2: b"unsupported", 3: b"cuda"}.get(code, b"?")

    def mkf_last_cuda_error(self):
        return b""

    def mkf_limits(self, a, b, c):
        for ref, v in ((a, 32), (b, 8), (c, 3)):
            ref._obj.value = v

    # ---- generic filter -----------------------------------------------------------
    def _filter(self, B, T, n, m, y, ys, A, As, Q, Qs, H, Hs, R, Rs, x0, x0s, P0, P0s, warmup, flags,
                sigma, detf, cnt, fm, fc, pm, pc, mle, nobs, status):
        if n > 32:
            return 2
        sig_o = _arr(sigma, (B, T)); det_o = _arr(detf, (B, T)); cnt_o = _arr(cnt, (B, T), np.int32)
        fm_o = _arr(fm, (B, T, n)); fc_o = _arr(fc, (B, T, n, n))
        pm_o = _arr(pm, (B, T, n)); pc_o = _arr(pc, (B, T, n, n))
        mle_o = _arr(mle, (B,)); nobs_o = _arr(nobs, (B,), np.int64); st_o = _arr(status, (B,), np.int32)
        for b in range(B):
            yb = _model(y, b, ys, (T, m))
            Ab, Qb = _model(A, b, As, (n, n)), _model(Q, b, Qs, (n, n))
            Hb, Rb = _model(H, b, Hs, (m, n)), _model(R, b, Rs, (m,))
            x0b = _model(x0, b, x0s, (n,))
            P0b = _model(P0, b, P0s, (n, n))
            x0b = np.zeros(n) if x0b is None else x0b
            P0b = np.eye(n) if P0b is None else P0b
            obs, idx, c = ko.pack_observations(yb)
            with np.errstate(all="ignore"):
                s, d, ns, f1, f2, p1, p2 = ko.seqkalmanfilter(obs, Ab, Qb, Hb, Rb, idx, c, x0b, P0b)
            seen = c > 0
            if sig_o is not None:
                sig_o[b] = 0.0; sig_o[b][seen] = s[:ns]
            if det_o is not None:
                det_o[b] = 0.0; det_o[b][seen] = d[:ns]
            if cnt_o is not None:
                cnt_o[b] = c
            for o, v in ((fm_o, f1), (fc_o, f2), (pm_o, p1), (pc_o, p2)):
                if o is not None:
                    o[b] = v
            if mle_o is not None:
                mle_o[b] = ko.get_mle(s[:ns], d[:ns], c, warmup)
            if nobs_o is not None:
                nobs_o[b] = int(np.sum(c[warmup:]))
            if st_o is not None:
                st_o[b] = 0 if np.isfinite(mle_o[b] if mle_o is not None else 0.0) else 1
        return 0

    def mkf_filter_history(self, B, T, n, m, y, ys, A, As, Q, Qs, H, Hs, R, Rs, x0, x0s, P0, P0s, warmup, flags,
                           sigma, detf, cnt, fm, fc, pm, pc, mle, nobs, status, stream):
        self.calls.append(("mkf_filter_history", B, T, n, m))
        return self._filter(B, T, n, m, y, ys, A, As, Q, Qs, H, Hs, R, Rs, x0, x0s, P0, P0s, warmup, flags,
                            sigma, detf, cnt, fm, fc, pm, pc, mle, nobs, status)

    def mkf_filter_loglik(self, B, T, n, m, y, ys, A, As, Q, Qs, H, Hs, R, Rs, x0, x0s, P0, P0s, warmup, flags,
                          mle, nobs, status, stream):
        self.calls.append(("mkf_filter_loglik", B, T, n, m))
        return self._filter(B, T, n, m, y, ys, A, As, Q, Qs, H, Hs, R, Rs, x0, x0s, P0, P0s, warmup, flags,
                            None, None, None, None, None, None, None, mle, nobs, status)

    # ---- DFM -----------------------------------------------------------------------
    def _dfm(self, B, T, m, k, y, ys, phi_of, q_of, gamma, gs, r, rs, warmup, mle, nobs, status):
        if m > 8 or k > 3:
            return 2
        n = m + k
        mle_o = _arr(mle, (B,)); nobs_o = _arr(nobs, (B,), np.int64); st_o = _arr(status, (B,), np.int32)
        for b in range(B):
            yb = _model(y, b, ys, (T, m))
            G = _model(gamma, b, gs, (m, k)) if k else np.zeros((m, 0))
            phi, q = phi_of(b, G), q_of(b, G)
            H = np.concatenate([np.eye(m), G], axis=1)
            Rb = _model(r, b, rs, (m,))
            Rb = np.zeros(m) if Rb is None else Rb
            val, nb = ko.filter_loglik(yb, np.diag(phi), np.diag(q), H, Rb, warmup=warmup)
            mle_o[b] = val
            if nobs_o is not None:
                nobs_o[b] = nb
            if st_o is not None:
                st_o[b] = 0 if np.isfinite(val) else 1
        return 0

    def mkf_dfm_loglik(self, B, T, m, k, y, ys, phi, ps, q, qs, gamma, gs, r, rs, warmup, mle, nobs, status, stream):
        self.calls.append(("mkf_dfm_loglik", B, T, m, k))
        n = m + k
        return self._dfm(B, T, m, k, y, ys, lambda b, G: _model(phi, b, ps, (n,)),
                         lambda b, G: _model(q, b, qs, (n,)), gamma, gs, r, rs, warmup, mle, nobs, status)

    def mkf_dfm_filter_history(self, B, T, m, k, y, ys, phi, ps, q, qs, gamma, gs, r, rs, warmup,
                               sigma, detf, cnt, fm, fc, pm, pc, mle, nobs, status, stream):
        self.calls.append(("mkf_dfm_filter_history", B, T, m, k))
        if m > 8 or k > 3:
            return 2
        n = m + k
        # expand to the generic layout in host memory, then reuse _filter
        A = np.zeros((B, n, n)); Q = np.zeros((B, n, n)); H = np.zeros((B, m, n)); R = np.zeros((B, m))
        for b in range(B):
            A[b] = np.diag(_model(phi, b, ps, (n,)))
            Q[b] = np.diag(_model(q, b, qs, (n,)))
            H[b, :, :m] = np.eye(m)
            if k:
                H[b, :, m:] = _model(gamma, b, gs, (m, k))
            if _addr(r):
                R[b] = _model(r, b, rs, (m,))
        P = lambda a_: a_.ctypes.data
        return self._filter(B, T, n, m, y, ys, P(A), n * n, P(Q), n * n, P(H), m * n, P(R), m, None, 0, None, 0,
                            warmup, 3, sigma, detf, cnt, fm, fc, pm, pc, mle, nobs, status)

    def mkf_dfm_loglik_alpha(self, B, T, m, k, y, ys, alpha, als, dt, gamma, gs, warmup, mle, nobs, status, stream):
        self.calls.append(("mkf_dfm_loglik_alpha", B, T, m, k))
        n = m + k

        def phi_of(b, G):
            return np.exp(-dt / _model(alpha, b, als, (n,)))

        def q_of(b, G):
            phi = phi_of(b, G)
            comm = np.sum(G ** 2, axis=1)
            return np.concatenate([(1 - phi[:m] ** 2) * (1 - comm), 1 - phi[m:] ** 2])

        return self._dfm(B, T, m, k, y, ys, phi_of, q_of, gamma, gs, None, 0, warmup, mle, nobs, status)

    def mkf_dfm_loglik_compact(self, B, J, m, k, yc, ycs, tidx, ts, alpha, als, dt, phi, ps, q, qs, gamma, gs, r, rs,
                               warmup, mle, nobs, status, stream):
        """Compacted grid: re-expanded to the full daily grid and filtered by the oracle, i.e. the
        reference's own step-by-step semantics."""
        self.calls.append(("mkf_dfm_loglik_compact", B, J, m, k))
        n = m + k
        ti_all = _arr(tidx, (B if ts else 1, J), np.int32)
        mle_o = _arr(mle, (B,)); nobs_o = _arr(nobs, (B,), np.int64); st_o = _arr(status, (B,), np.int32)
        for b in range(B):
            ycb = _model(yc, b, ycs, (J, m))
            ti = ti_all[b if ts else 0].astype(np.int64)
            T = int(ti.max()) + 1 if ti.size and ti.max() >= 0 else 1
            yb = np.full((T, m), np.nan)
            keep = np.concatenate([[True], np.diff(ti) > 0]) & (ti >= 0)
            yb[ti[keep]] = ycb[keep]
            G = _model(gamma, b, gs, (m, k)) if k else np.zeros((m, 0))
            if _addr(alpha):
                ph = np.exp(-dt / _model(alpha, b, als, (n,)))
                comm = np.sum(G ** 2, axis=1)
                qq = np.concatenate([(1 - ph[:m] ** 2) * (1 - comm), 1 - ph[m:] ** 2])
            else:
                ph, qq = _model(phi, b, ps, (n,)), _model(q, b, qs, (n,))
            H = np.concatenate([np.eye(m), G], axis=1)
            Rb = _model(r, b, rs, (m,))
            Rb = np.zeros(m) if Rb is None else Rb
            val, nb = ko.filter_loglik(yb, np.diag(ph), np.diag(qq), H, Rb, warmup=warmup)
            mle_o[b] = val
            if nobs_o is not None:
                nobs_o[b] = nb
            if st_o is not None:
                st_o[b] = 0 if np.isfinite(val) else 1
        return 0

    def mkf_dfm_set_handoff(self, enabled):
        return 1

    def mkf_dfm_set_hist_tma(self, enabled):
        return 1

    def mkf_smoother_set_thread_min_batch(self, min_batch):
        return 6144

    # ---- smoother / projections / reduction ---------------------------------------------
    def mkf_smoother(self, B, T, n, fm, fc, pm, pc, A, As, flags, sm, sc, status, stream):
        self.calls.append(("mkf_smoother", B, T, n))
        fm_, fc_ = _arr(fm, (B, T, n)), _arr(fc, (B, T, n, n))
        pm_, pc_ = _arr(pm, (B, T, n)), _arr(pc, (B, T, n, n))
        sm_, sc_ = _arr(sm, (B, T, n)), _arr(sc, (B, T, n, n))
        st_ = _arr(status, (B,), np.int32)
        for b in range(B):
            s1, s2 = ko.kalmansmoother(fm_[b], fc_[b], pm_[b], pc_[b], _model(A, b, As, (n, n)))
            sm_[b], sc_[b] = s1, s2
            if st_ is not None:
                st_[b] = 0
        return 0

    def mkf_smoother_project(self, B, T, n, fm, fc, pm, pc, A, As, flags, p, Hp, Hs, sm, sc, pmean, pvar, svar,
                             status, stream):
        self.calls.append(("mkf_smoother_project", B, T, n, p))
        fm_, fc_ = _arr(fm, (B, T, n)), _arr(fc, (B, T, n, n))
        pm_, pc_ = _arr(pm, (B, T, n)), _arr(pc, (B, T, n, n))
        sm_, sc_ = _arr(sm, (B, T, n)), _arr(sc, (B, T, n, n))
        mo, vo, so = _arr(pmean, (B, T, p)), _arr(pvar, (B, T, p)), _arr(svar, (B, T, n))
        st_ = _arr(status, (B,), np.int32)
        for b in range(B):
            s1, s2 = ko.kalmansmoother(fm_[b], fc_[b], pm_[b], pc_[b], _model(A, b, As, (n, n)))
            sm_[b] = s1
            if sc_ is not None:
                sc_[b] = s2
            H = _model(Hp, b, Hs, (p, n))
            mo[b] = s1 @ H.T
            vo[b] = np.maximum(np.einsum("ij,tjk,ik->ti", H, s2, H), 0)
            if so is not None:
                so[b] = np.diagonal(s2, axis1=1, axis2=2)
            if st_ is not None:
                st_[b] = 0
        return 0

    def mkf_project(self, B, T, n, p, means, covs, Hp, Hs, mout, vout, stream):
        self.calls.append(("mkf_project", B, T, n, p))
        mu, cv = _arr(means, (B, T, n)), _arr(covs, (B, T, n, n))
        mo, vo = _arr(mout, (B, T, p)), _arr(vout, (B, T, p))
        for b in range(B):
            H = _model(Hp, b, Hs, (p, n))
            if mo is not None:
                mo[b] = mu[b] @ H.T
            if vo is not None:
                vo[b] = np.maximum(np.einsum("ij,tjk,ik->ti", H, cv[b], H), 0)
        return 0

    def mkf_reduce_sum(self, v, count, out, stream):
        import math

        _arr(out, (1,))[0] = math.fsum(_arr(v, (count,)).tolist()) if count else 0.0
        return 0

    def mkf_probe_fp64(self, blocks, iters, scratch, stream):
        return 0
