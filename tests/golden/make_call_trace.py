"""Record the call trace the UNMODIFIED reference facade makes on its Kalman object.

    python tests/golden/make_call_trace.py        # build container only (needs /root/reference)

``Metran`` (metran/metran.py) only ever talks to ``SPKalmanFilter`` through a handful of methods
and attributes.  Here the reference's own class is subclassed by a
recorder -- every call with its arguments, in order, plus what the REFERENCE (numba engine)
answered: get_mle values, sigmas / detfs, state arrays at selected steps, simulate / decompose
outputs -- while the reference's example workflow runs: solve(), the state / simulation /
decomposition getters, mask_observations(), get_simulation() on the masked data, unmask
(examples/example_script.py:20-53, tests/test_metran.py:32-40).

tests/test_call_trace.py replays the trace against this package's SPKalmanFilter: on the CPU
through the oracle-backed stand-in (proves the harness), on the GPU box against the real CUDA
library -- the composition "unmodified facade + CUDA engine" without needing the reference there.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refload  # noqa: E402


def _sel(T):
    idx = sorted(set(list(range(0, 12)) + list(range(12, T, 211)) + list(range(T - 12, T))))
    return np.array([i for i in idx if 0 <= i < T], dtype=np.int64)


def main():
    metran = _refload.load_reference()
    kfmod = metran.kalmanfilter
    RefKF = kfmod.SPKalmanFilter
    events = []          # (name, {field: array})
    depth = [0]

    def rec(name, **fields):
        if depth[0] == 0:
            events.append((name, {k: np.asarray(v) for k, v in fields.items() if v is not None}))

    class RecordingKF(RefKF):
        def __init__(self, engine="numba"):
            depth[0] += 1
            super().__init__(engine=engine)
            depth[0] -= 1
            rec("__init__", engine=engine)

        def init_states(self):
            super().init_states()
            rec("init_states")

        def set_observations(self, oseries):
            super().set_observations(oseries)
            rec("set_observations", values=oseries.values.astype(float), observation_count=self.observation_count)

        def set_matrices(self, A, Q, H, R):
            super().set_matrices(A, Q, H, R)
            rec("set_matrices", A=A, Q=Q, H=H, R=R)

        def _state_fields(self, smoothed):
            T = self.filtered_state_means.shape[0]
            s = _sel(T)
            out = dict(sel=s, sigmas=self.sigmas, detfs=self.detfs, fm=self.filtered_state_means[s],
                       fc=self.filtered_state_covariances[s], pm=self.predicted_state_means[s],
                       pc=self.predicted_state_covariances[s])
            if smoothed:
                out.update(sm=self.smoothed_state_means[s], sc=self.smoothed_state_covariances[s])
            return out

        def run_filter(self, initial_state_mean=None, initial_state_covariance=None, engine=None):
            depth[0] += 1
            super().run_filter(initial_state_mean, initial_state_covariance, engine)
            depth[0] -= 1
            rec("run_filter", mask=np.bool_(self.mask), **self._state_fields(False))

        def run_smoother(self):
            depth[0] += 1
            super().run_smoother()
            depth[0] -= 1
            rec("run_smoother", mask=np.bool_(self.mask), **self._state_fields(True))

        def get_mle(self, warmup=1):
            r = super().get_mle(warmup)
            rec("get_mle", warmup=warmup, result=np.float64(r))
            return r

        def simulate(self, observation_matrix, method="smoother"):
            means, variances = super().simulate(observation_matrix, method)
            s = _sel(len(means))
            rec("simulate", H=observation_matrix, method=method, sel=s, means=np.asarray(means)[s],
                variances=np.asarray(variances)[s])
            return means, variances

        def decompose(self, observation_matrix, method="smoother"):
            sdf, cdf = super().decompose(observation_matrix, method)
            s = _sel(len(sdf))
            rec("decompose", H=observation_matrix, method=method, sel=s, sdf=np.asarray(sdf)[s],
                cdf=np.stack([np.asarray(c)[s] for c in cdf]))
            return sdf, cdf

    metran.metran.SPKalmanFilter = RecordingKF
    try:
        mt = metran.Metran(_refload.load_example_series(), name="B21B0214")
        mt.solve(report=False)
        mt.get_state_means()
        mt.get_state_variances()
        mt.get_simulated_means()
        mt.get_simulated_variances()
        mt.get_simulation("B21B0214005", alpha=0.05)
        mt.decompose_simulation("B21B0214001")
        mt.get_state_means(method="filter")
        oseries = mt.get_observations()
        mask = (0 * oseries).astype(bool)
        mask.loc["1997-8-28", "B21B0214005"] = True
        mt.mask_observations(mask)
        mt.get_simulation("B21B0214005", alpha=0.05)
        mt.unmask_observations()
        mt.get_simulation("B21B0214005", alpha=0.05)
    finally:
        metran.metran.SPKalmanFilter = RefKF

    out = {"n_events": np.int64(len(events)), "names": np.array([e[0] for e in events])}
    for i, (name, fields) in enumerate(events):
        for k, v in fields.items():
            out["%04d/%s" % (i, k)] = v
    np.savez_compressed(os.path.join(HERE, "call_trace_b21.npz"), **out)
    from collections import Counter

    print(len(events), "events:", dict(Counter(e[0] for e in events)))


if __name__ == "__main__":
    main()
