"""Batched Kalman filter / smoother on device tensors -- thin layer over the C ABI.

PyTorch is used for device memory and streams only; every computation below is a call into
libmetran_b200.so (include/metran_b200.h).  Shapes follow the reference's arrays
(metran/kalmanfilter.py:243-253) with a leading batch axis; an array given WITHOUT the
batch axis is shared by all models of the batch (stride 0).
"""
import ctypes

import numpy as np
import torch

from . import _lib

__all__ = ["KalmanEngine", "default_engine"]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class KalmanEngine:
    """Owns the library handle and the device; all methods enqueue on the current stream.

    ``lib`` / ``device`` exist so host-side logic can be unit-tested without a GPU by
    injecting a fake library (tests/fakelib.py); the product always uses the CUDA library
    and a CUDA device and raises if either is missing.
    """

    def __init__(self, device=None, lib=None):
        if lib is None:
            lib = _lib.load()
            if not torch.cuda.is_available():
                raise _lib.MetranB200Error(
                    "metran_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        elif device is None:
            device = torch.device("cpu")
        self.lib = lib
        self.device = torch.device(device)
        self.limits = _lib.limits(lib)

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        if self.device.type == "cuda":
            return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def tensor(self, a, dtype=torch.float64):
        """Host array / tensor -> contiguous device tensor (no copy if already there)."""
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    @staticmethod
    def _stride(t, base_ndim):
        """Element stride between models: 0 when the batch axis is absent (shared)."""
        if t is None or t.dim() == base_ndim:
            return 0
        return int(np.prod(t.shape[1:]))

    @staticmethod
    def _batch(*pairs):
        B = None
        for t, base in pairs:
            if t is not None and t.dim() == base + 1:
                if B is not None and t.shape[0] != B:
                    raise ValueError("inconsistent batch sizes: %d vs %d" % (B, t.shape[0]))
                B = t.shape[0]
        return 1 if B is None else int(B)

    def structure_flags(self, A, Q):
        """MKF_FLAG_* from the actual matrices (one device reduction + sync)."""
        n = A.shape[-1]
        eye = torch.eye(n, dtype=torch.bool, device=A.device)
        offA = torch.count_nonzero(A.masked_fill(eye, 0.0))
        offQ = torch.count_nonzero(Q.masked_fill(eye, 0.0))
        offs = torch.stack([offA, offQ]).cpu()
        return (_lib.FLAG_A_DIAGONAL if offs[0] == 0 else 0) | (_lib.FLAG_Q_DIAGONAL if offs[1] == 0 else 0)

    # ------------------------------------------------------------------ generic SSM
    def _ssm_args(self, y, A, Q, H, R, x0, P0):
        y, A, Q, H, R = (self.tensor(v) for v in (y, A, Q, H, R))
        x0, P0 = self.tensor(x0), self.tensor(P0)
        B = self._batch((y, 2), (A, 2), (Q, 2), (H, 2), (R, 1), (x0, 1), (P0, 2))
        T, m = int(y.shape[-2]), int(y.shape[-1])
        n = int(A.shape[-1])
        if H.shape[-2:] != (m, n) or Q.shape[-2:] != (n, n) or R.shape[-1] != m:
            raise ValueError("inconsistent state-space shapes")
        return y, A, Q, H, R, x0, P0, B, T, n, m

    def filter_loglik(self, y, A, Q, H, R, x0=None, P0=None, warmup=1, flags=None):
        """-2 log L per model: run_filter + get_mle (kalmanfilter.py:696-778, 550-567).

        Returns (mle[B] f64, nobs[B] i64, status[B] i32) device tensors.
        """
        y, A, Q, H, R, x0, P0, B, T, n, m = self._ssm_args(y, A, Q, H, R, x0, P0)
        if flags is None:
            flags = self.structure_flags(A, Q)
        mle = self.empty(B)
        nobs = self.empty(B, dtype=torch.int64)
        status = self.empty(B, dtype=torch.int32)
        rc = self.lib.mkf_filter_loglik(
            B, T, n, m, _ptr(y), self._stride(y, 2), _ptr(A), self._stride(A, 2),
            _ptr(Q), self._stride(Q, 2), _ptr(H), self._stride(H, 2), _ptr(R), self._stride(R, 1),
            _ptr(x0), self._stride(x0, 1), _ptr(P0), self._stride(P0, 2),
            int(warmup), int(flags), _ptr(mle), _ptr(nobs), _ptr(status), self._stream())
        _lib.check(self.lib, rc, "mkf_filter_loglik")
        return mle, nobs, status

    def filter_history(self, y, A, Q, H, R, x0=None, P0=None, warmup=1, flags=None, covariances=True):
        """Forward filter with the reference's full result contract (kalmanfilter.py:392-400).

        Returns a dict of device tensors: sigma[B,T], detf[B,T], cnt[B,T] (uncompacted, see
        the header), fm[B,T,n], fc[B,T,n,n], pm, pc, mle[B], nobs[B], status[B].
        """
        y, A, Q, H, R, x0, P0, B, T, n, m = self._ssm_args(y, A, Q, H, R, x0, P0)
        if flags is None:
            flags = self.structure_flags(A, Q)
        out = dict(
            sigma=self.empty(B, T), detf=self.empty(B, T), cnt=self.empty(B, T, dtype=torch.int32),
            fm=self.empty(B, T, n), pm=self.empty(B, T, n),
            fc=self.empty(B, T, n, n) if covariances else None,
            pc=self.empty(B, T, n, n) if covariances else None,
            mle=self.empty(B), nobs=self.empty(B, dtype=torch.int64),
            status=self.empty(B, dtype=torch.int32))
        rc = self.lib.mkf_filter_history(
            B, T, n, m, _ptr(y), self._stride(y, 2), _ptr(A), self._stride(A, 2),
            _ptr(Q), self._stride(Q, 2), _ptr(H), self._stride(H, 2), _ptr(R), self._stride(R, 1),
            _ptr(x0), self._stride(x0, 1), _ptr(P0), self._stride(P0, 2),
            int(warmup), int(flags),
            _ptr(out["sigma"]), _ptr(out["detf"]), _ptr(out["cnt"]),
            _ptr(out["fm"]), _ptr(out["fc"]), _ptr(out["pm"]), _ptr(out["pc"]),
            _ptr(out["mle"]), _ptr(out["nobs"]), _ptr(out["status"]), self._stream())
        _lib.check(self.lib, rc, "mkf_filter_history")
        out["flags"] = flags
        return out

    def smoother(self, fm, fc, pm, pc, A, flags=None):
        """RTS smoother (kalmanfilter.py:403-476): returns (sm[B,T,n], sc[B,T,n,n], status[B])."""
        fm, fc, pm, pc, A = (self.tensor(v) for v in (fm, fc, pm, pc, A))
        if fm.dim() == 2:
            fm, fc, pm, pc = fm[None], fc[None], pm[None], pc[None]
        B, T, n = (int(s) for s in fm.shape)
        if flags is None:
            flags = self.structure_flags(A, A)
        sm = self.empty(B, T, n)
        sc = self.empty(B, T, n, n)
        status = self.empty(B, dtype=torch.int32)
        rc = self.lib.mkf_smoother(B, T, n, _ptr(fm), _ptr(fc), _ptr(pm), _ptr(pc),
                                   _ptr(A), self._stride(A, 2), int(flags),
                                   _ptr(sm), _ptr(sc), _ptr(status), self._stream())
        _lib.check(self.lib, rc, "mkf_smoother")
        return sm, sc, status

    def smoother_project(self, fm, fc, pm, pc, A, Hp, flags=None, covariances=True, state_variances=True):
        """RTS smoother with the projections fused into its epilogue (mkf_smoother_project): returns a
        dict sm[B,T,n], sc[B,T,n,n] | None, proj_mean[B,T,p], proj_var[B,T,p], state_var[B,T,n] | None,
        status[B].  Nobody has to read the [T,n,n] covariances back to project them."""
        fm, fc, pm, pc, A, Hp = (self.tensor(v) for v in (fm, fc, pm, pc, A, Hp))
        if fm.dim() == 2:
            fm, fc, pm, pc = fm[None], fc[None], pm[None], pc[None]
        B, T, n = (int(s) for s in fm.shape)
        p = int(Hp.shape[-2])
        if flags is None:
            flags = self.structure_flags(A, A)
        out = dict(sm=self.empty(B, T, n), sc=self.empty(B, T, n, n) if covariances else None,
                   proj_mean=self.empty(B, T, p), proj_var=self.empty(B, T, p),
                   state_var=self.empty(B, T, n) if state_variances else None,
                   status=self.empty(B, dtype=torch.int32))
        rc = self.lib.mkf_smoother_project(B, T, n, _ptr(fm), _ptr(fc), _ptr(pm), _ptr(pc),
                                           _ptr(A), self._stride(A, 2), int(flags), p, _ptr(Hp), self._stride(Hp, 2),
                                           _ptr(out["sm"]), _ptr(out["sc"]), _ptr(out["proj_mean"]),
                                           _ptr(out["proj_var"]), _ptr(out["state_var"]), _ptr(out["status"]),
                                           self._stream())
        _lib.check(self.lib, rc, "mkf_smoother_project")
        return out

    def project(self, means, covs, Hp, variances=True):
        """H x_t and max(diag(H P_t H^T), 0) for all t (kalmanfilter.py:596-602)."""
        means, covs, Hp = self.tensor(means), self.tensor(covs), self.tensor(Hp)
        if means.dim() == 2:
            means = means[None]
            covs = covs[None] if covs is not None else None
        B, T, n = (int(s) for s in means.shape)
        p = int(Hp.shape[-2])
        mout = self.empty(B, T, p)
        vout = self.empty(B, T, p) if variances else None
        rc = self.lib.mkf_project(B, T, n, p, _ptr(means), _ptr(covs), _ptr(Hp), self._stride(Hp, 2),
                                  _ptr(mout), _ptr(vout), self._stream())
        _lib.check(self.lib, rc, "mkf_project")
        return mout, vout

    # ------------------------------------------------------------------ Metran-shaped DFM
    def dfm_supported(self, m, k):
        return 1 <= m <= self.limits["dfm_m_max"] and 0 <= k <= self.limits["dfm_k_max"]

    def dfm_loglik(self, y, phi, q, gamma, r=None, warmup=1, out=None):
        """-2 log L of Metran-shaped models, thread-per-model kernel (mkf_dfm_loglik).

        y[B,T,m] (or [T,m] shared), phi/q[B,n], gamma[B,m,k] (or shared [m,k]), r[B,m]|None;
        ``out`` = (mle[B], nobs[B] int64, status[B] int32) device tensors to reuse.
        """
        y, phi, q, gamma, r = (self.tensor(v) for v in (y, phi, q, gamma, r))
        B = self._batch((y, 2), (phi, 1), (q, 1), (gamma, 2), (r, 1))
        T, m = int(y.shape[-2]), int(y.shape[-1])
        k = int(phi.shape[-1]) - m
        if gamma.shape[-2:] != (m, k):
            raise ValueError("gamma must be [.., m, k] with k = n - m")
        if out is None:
            out = (self.empty(B), self.empty(B, dtype=torch.int64), self.empty(B, dtype=torch.int32))
        mle, nobs, status = out
        rc = self.lib.mkf_dfm_loglik(
            B, T, m, k, _ptr(y), self._stride(y, 2), _ptr(phi), self._stride(phi, 1),
            _ptr(q), self._stride(q, 1), _ptr(gamma), self._stride(gamma, 2),
            _ptr(r), self._stride(r, 1), int(warmup),
            _ptr(mle), _ptr(nobs), _ptr(status), self._stream())
        _lib.check(self.lib, rc, "mkf_dfm_loglik")
        return mle, nobs, status

    def dfm_filter_history(self, y, phi, q, gamma, r=None, warmup=1, covariances=True):
        """filter_history for Metran-shaped models on the thread-per-model kernel
        (mkf_dfm_filter_history): same result dict as ``filter_history``."""
        y, phi, q, gamma, r = (self.tensor(v) for v in (y, phi, q, gamma, r))
        B = self._batch((y, 2), (phi, 1), (q, 1), (gamma, 2), (r, 1))
        T, m = int(y.shape[-2]), int(y.shape[-1])
        n = int(phi.shape[-1])
        k = n - m
        if gamma.shape[-2:] != (m, k):
            raise ValueError("gamma must be [.., m, k] with k = n - m")
        out = dict(
            sigma=self.empty(B, T), detf=self.empty(B, T), cnt=self.empty(B, T, dtype=torch.int32),
            fm=self.empty(B, T, n), pm=self.empty(B, T, n),
            fc=self.empty(B, T, n, n) if covariances else None,
            pc=self.empty(B, T, n, n) if covariances else None,
            mle=self.empty(B), nobs=self.empty(B, dtype=torch.int64),
            status=self.empty(B, dtype=torch.int32))
        rc = self.lib.mkf_dfm_filter_history(
            B, T, m, k, _ptr(y), self._stride(y, 2), _ptr(phi), self._stride(phi, 1),
            _ptr(q), self._stride(q, 1), _ptr(gamma), self._stride(gamma, 2),
            _ptr(r), self._stride(r, 1), int(warmup),
            _ptr(out["sigma"]), _ptr(out["detf"]), _ptr(out["cnt"]),
            _ptr(out["fm"]), _ptr(out["fc"]), _ptr(out["pm"]), _ptr(out["pc"]),
            _ptr(out["mle"]), _ptr(out["nobs"]), _ptr(out["status"]), self._stream())
        _lib.check(self.lib, rc, "mkf_dfm_filter_history")
        out["flags"] = _lib.FLAG_A_DIAGONAL | _lib.FLAG_Q_DIAGONAL
        return out

    def dfm_loglik_alpha(self, y, alpha, gamma, dt=1.0, warmup=1, out=None):
        """As dfm_loglik with Metran's alpha -> (phi, q) map fused (metran.py:246-416)."""
        y, alpha, gamma = self.tensor(y), self.tensor(alpha), self.tensor(gamma)
        B = self._batch((y, 2), (alpha, 1), (gamma, 2))
        T, m = int(y.shape[-2]), int(y.shape[-1])
        k = int(alpha.shape[-1]) - m
        if gamma.shape[-2:] != (m, k):
            raise ValueError("gamma must be [.., m, k] with k = n - m")
        if out is None:
            out = (self.empty(B), self.empty(B, dtype=torch.int64), self.empty(B, dtype=torch.int32))
        mle, nobs, status = out
        rc = self.lib.mkf_dfm_loglik_alpha(
            B, T, m, k, _ptr(y), self._stride(y, 2), _ptr(alpha), self._stride(alpha, 1), float(dt),
            _ptr(gamma), self._stride(gamma, 2), int(warmup),
            _ptr(mle), _ptr(nobs), _ptr(status), self._stream())
        _lib.check(self.lib, rc, "mkf_dfm_loglik_alpha")
        return mle, nobs, status

    # ---- compacted time grid --------------------------------------------
    def compact_observations(self, y):
        """Drop the time steps without any observation: y[B,T,m] | [T,m] (NaN-coded) ->
        (yc[B,J,m] | [J,m], tidx[B,J] | [J] int32, J, observed_fraction).

        Metran's daily grid (metran.py:571) leaves most steps of irregularly sampled series
        empty (94.5 % of the reference's example data); the filter only predicts on them
        (kalmanfilter.py:335 not taken), which mkf_dfm_loglik_compact does in closed form.
        J is the largest observed-step count of the batch; shorter models are padded with
        NaN rows that repeat their last time index (a no-op for the kernel)."""
        y = self.tensor(y)
        shared = y.dim() == 2
        yy = y[None] if shared else y
        B, T, m = (int(v) for v in yy.shape)
        obs = torch.isfinite(yy) & (yy != -1e10)           # rule of set_observations (:657, :666-667)
        anyobs = obs.any(dim=-1)
        cnt = anyobs.sum(dim=-1)
        J = max(int(cnt.max()) if B else 0, 1)
        order = torch.argsort((~anyobs).to(torch.int8), dim=-1, stable=True)[:, :J]
        valid = torch.arange(J, device=yy.device)[None, :] < cnt[:, None]
        last = order.gather(1, (cnt - 1).clamp(min=0)[:, None])
        last = torch.where(cnt[:, None] > 0, last, torch.full_like(last, -1))
        tidx = torch.where(valid, order, last.expand(-1, J)).to(torch.int32).contiguous()
        yc = yy.gather(1, order[:, :, None].expand(-1, -1, m)).contiguous()
        yc.masked_fill_(~valid[:, :, None], float("nan"))
        frac = float(cnt.sum()) / max(B * T, 1)
        if shared:
            return yc[0], tidx[0], J, frac
        return yc, tidx, J, frac

    def dfm_loglik_compact(self, yc, tidx, gamma, alpha=None, phi=None, q=None, r=None, dt=1.0, warmup=1, out=None):
        """-2 log L of Metran-shaped models on a compacted time grid (mkf_dfm_loglik_compact):
        yc/tidx from ``compact_observations``; either alpha[B,n] (with dt) or phi, q[B,n]."""
        yc, gamma = self.tensor(yc), self.tensor(gamma)
        tidx = tidx if isinstance(tidx, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(tidx))
        tidx = tidx.to(device=self.device, dtype=torch.int32).contiguous()
        alpha, phi, q, r = (self.tensor(v) for v in (alpha, phi, q, r))
        if (alpha is None) == (phi is None or q is None):
            raise ValueError("give either alpha or phi and q")
        par = alpha if alpha is not None else phi
        B = self._batch((yc, 2), (par, 1), (gamma, 2), (r, 1))
        J, m = int(yc.shape[-2]), int(yc.shape[-1])
        k = int(par.shape[-1]) - m
        if gamma.shape[-2:] != (m, k):
            raise ValueError("gamma must be [.., m, k] with k = n - m")
        if tidx.shape[-1] != J or (tidx.dim() == 2) != (yc.dim() == 3):
            raise ValueError("tidx must match yc: [B, J] with yc[B, J, m], [J] with yc[J, m]")
        if out is None:
            out = (self.empty(B), self.empty(B, dtype=torch.int64), self.empty(B, dtype=torch.int32))
        mle, nobs, status = out
        rc = self.lib.mkf_dfm_loglik_compact(
            B, J, m, k, _ptr(yc), self._stride(yc, 2), _ptr(tidx), 0 if tidx.dim() == 1 else J,
            _ptr(alpha), self._stride(alpha, 1), float(dt), _ptr(phi), self._stride(phi, 1),
            _ptr(q), self._stride(q, 1), _ptr(gamma), self._stride(gamma, 2), _ptr(r), self._stride(r, 1),
            int(warmup), _ptr(mle), _ptr(nobs), _ptr(status), self._stream())
        _lib.check(self.lib, rc, "mkf_dfm_loglik_compact")
        return mle, nobs, status

    def dfm_loglik_alpha_host(self, y_host, alpha, gamma, dt=1.0, warmup=1, chunk_models=2048):
        """Host-resident observations -> -2 log L per model, H2D overlapped with compute.

        ``y_host``: CPU tensor [B,T,m] (pin it for full PCIe speed); ``alpha`` [B,n] and
        ``gamma`` [B,m,k] | [m,k]: host arrays.  The batch is cut into chunks of
        ``chunk_models`` models; chunk c+1 is copied on a side stream into the second of
        two device staging buffers while the filter kernel runs on chunk c.  Returns
        (mle, nobs, status) as pinned CPU tensors after the final device->host read.
        """
        if self.device.type != "cuda":
            raise _lib.MetranB200Error("dfm_loglik_alpha_host needs a CUDA device")
        y_host = torch.as_tensor(y_host)
        B, T, m = (int(s) for s in y_host.shape)
        chunk = int(min(chunk_models, B))
        key = (chunk, T, m)
        st = getattr(self, "_staging", None)
        if st is None or st["key"] != key:
            st = dict(key=key, bufs=[self.empty(chunk, T, m), self.empty(chunk, T, m)],
                      copy_stream=torch.cuda.Stream(self.device),
                      copied=[torch.cuda.Event(), torch.cuda.Event()],
                      freed=[torch.cuda.Event(), torch.cuda.Event()])
            self._staging = st
        main = torch.cuda.current_stream(self.device)
        alpha_d = self.tensor(torch.as_tensor(alpha)) if not isinstance(alpha, torch.Tensor) else alpha.to(self.device, non_blocking=True)
        gamma_d = self.tensor(torch.as_tensor(gamma)) if not isinstance(gamma, torch.Tensor) else gamma.to(self.device, non_blocking=True)
        mle = self.empty(B)
        nobs = self.empty(B, dtype=torch.int64)
        status = self.empty(B, dtype=torch.int32)
        shared_gamma = gamma_d.dim() == 2
        for ci, b0 in enumerate(range(0, B, chunk)):
            b1 = min(B, b0 + chunk)
            s = ci & 1
            buf = st["bufs"][s][: b1 - b0]
            with torch.cuda.stream(st["copy_stream"]):
                if ci >= 2:
                    st["copy_stream"].wait_event(st["freed"][s])
                else:
                    st["copy_stream"].wait_stream(main)
                buf.copy_(y_host[b0:b1], non_blocking=True)
                st["copied"][s].record(st["copy_stream"])
            main.wait_event(st["copied"][s])
            self.dfm_loglik_alpha(buf, alpha_d[b0:b1], gamma_d if shared_gamma else gamma_d[b0:b1],
                                  dt=dt, warmup=warmup, out=(mle[b0:b1], nobs[b0:b1], status[b0:b1]))
            st["freed"][s].record(main)
        res = getattr(self, "_host_out", None)
        if res is None or res[0].shape[0] != B:
            res = (torch.empty(B, dtype=torch.float64).pin_memory(),
                   torch.empty(B, dtype=torch.int64).pin_memory(),
                   torch.empty(B, dtype=torch.int32).pin_memory())
            self._host_out = res
        res[0].copy_(mle, non_blocking=True)
        res[1].copy_(nobs, non_blocking=True)
        res[2].copy_(status, non_blocking=True)
        main.synchronize()
        return res

    def set_handoff(self, enabled):
        """Three-launch path of the DFM log-likelihood (transient -> steady state -> resume) on/off;
        returns the previous setting.  Results are bit-identical either way."""
        return bool(self.lib.mkf_dfm_set_handoff(1 if enabled else 0))

    def dfm_profile(self, fn):
        """Run ``fn()`` (one DFM log-likelihood call) with per-launch CUDA events recorded inside the
        library; returns {"ms": [transient, steady, resume], "warps", "warps_handed_off",
        "warp_steps_after_handoff"} or None when the call took the single-kernel path."""
        prev = self.lib.mkf_dfm_set_profile(1)
        try:
            fn()
            ms = (ctypes.c_float * 3)()
            st = (ctypes.c_int64 * 3)()
            rc = self.lib.mkf_dfm_last_profile(self._stream(), ctypes.cast(ms, ctypes.c_void_p),
                                               ctypes.cast(st, ctypes.c_void_p))
        finally:
            self.lib.mkf_dfm_set_profile(prev)
        if rc != _lib.MKF_OK or ms[0] < 0:
            return None
        return {"ms": [float(v) for v in ms], "warps": int(st[0]), "warps_handed_off": int(st[1]),
                "warp_steps_after_handoff": int(st[2])}

    def probe_fp64_tflops(self, blocks=148 * 8, iters=4096, reps=5):
        """Measured fp64 FMA throughput (TFLOP/s) of this device: roofline denominator."""
        scratch = self.empty(1)
        best = None
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = self.lib.mkf_probe_fp64(int(blocks), int(iters), _ptr(scratch), self._stream())
            e1.record()
            _lib.check(self.lib, rc, "mkf_probe_fp64")
            torch.cuda.synchronize(self.device)
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        flops = blocks * 256.0 * iters * 64 * 2
        return flops / (best * 1e-3) / 1e12

    def reduce_sum(self, v, out=None):
        """Deterministic sum of a device vector into a 1-element tensor (all-reduce input)."""
        v = self.tensor(v)
        if out is None:
            out = self.empty(1)
        rc = self.lib.mkf_reduce_sum(_ptr(v), int(v.numel()), _ptr(out), self._stream())
        _lib.check(self.lib, rc, "mkf_reduce_sum")
        return out


_default = None


def default_engine():
    global _default
    if _default is None:
        _default = KalmanEngine()
    return _default
