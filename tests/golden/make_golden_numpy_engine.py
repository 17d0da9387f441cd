"""Golden values from the reference's SECOND engine, the numpy one.

    python tests/golden/make_golden_numpy_engine.py        # build container only

Runs ``SPKalmanFilter(engine="numpy")`` of the UNMODIFIED reference -- filter_predict /
filter_update / seqkalmanfilter_np, metran/kalmanfilter.py:15-233 -- on inputs already stored in
synthetic_cases.npz (so the numba-engine results of the same inputs sit next to them) and writes
tests/golden/numpy_engine.npz.  One diagonal-A case (Metran-shaped, 30 % missing) and one full-A
case: the numpy engine predicts the covariance with the full ``A P A^T + Q`` (:40-45) while the
numba engine uses diag(A) only (:324-331, quirk Q1), so the two engines agree to rounding on the
first and differ on the second.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refload  # noqa: E402

CASES = ("dfm_n8_m5_miss30", "dense_n4_m3_fullA", "dense_n16_m7")


def main():
    metran = _refload.load_reference()
    kfmod = metran.kalmanfilter
    z = np.load(os.path.join(HERE, "synthetic_cases.npz"))
    out = {}
    for tag in CASES:
        y, A, Q, H, R = (z[tag + "/" + k] for k in ("y", "A", "Q", "H", "R"))
        kf = kfmod.SPKalmanFilter(engine="numpy")
        assert kf.filtermethod is kfmod.seqkalmanfilter_np
        kf.set_observations(pd.DataFrame(y))
        kf.set_matrices(A, Q, H, R)
        kf.run_filter()
        sm, sc = kfmod.kalmansmoother(kf.filtered_state_means, kf.filtered_state_covariances,
                                      kf.predicted_state_means, kf.predicted_state_covariances, A)
        d = dict(sigmas=np.asarray(kf.sigmas, dtype=float), detfs=np.asarray(kf.detfs, dtype=float),
                 fm=kf.filtered_state_means, fc=kf.filtered_state_covariances,
                 pm=kf.predicted_state_means, pc=kf.predicted_state_covariances, sm=sm, sc=sc,
                 mle=np.float64(kf.get_mle()), mle_w0=np.float64(kf.get_mle(warmup=0)))
        for k, v in d.items():
            out[tag + "/" + k] = np.asarray(v)
        print(tag, "numpy-engine mle=%r  numba-engine mle=%r" % (d["mle"], float(z[tag + "/mle"])))
    np.savez_compressed(os.path.join(HERE, "numpy_engine.npz"), **out)


if __name__ == "__main__":
    main()
