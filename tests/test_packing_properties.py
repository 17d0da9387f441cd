"""Property tests (hypothesis) of the host-side observation packing against the oracle's loop
restatement of SPKalmanFilter.set_observations (metran/kalmanfilter.py:646-674)."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

from metran_b200.kalmanfilter import pack_observations, unpack_observations
from oracle import kalman_oracle as ko

values = st.one_of(st.floats(-1e6, 1e6), st.just(np.nan), st.just(np.inf), st.just(-np.inf), st.just(-1e10), st.just(0.0))


@settings(max_examples=150, deadline=None)
@given(hnp.arrays(np.float64, hnp.array_shapes(min_dims=2, max_dims=2, min_side=1, max_side=9), elements=values))
def test_pack_matches_oracle_and_round_trips(y):
    obs, idx, cnt, observed = pack_observations(y)
    o2, i2, c2 = ko.pack_observations(y)
    assert np.array_equal(obs, o2) and np.array_equal(idx, i2) and np.array_equal(cnt, c2)
    assert obs.dtype == np.float64 and idx.dtype == np.float64 and cnt.dtype == np.int64
    # missing = not finite or exactly -1e10 (kalmanfilter.py:657, 666-667)
    expect = np.isfinite(y) & (y != -1e10)
    assert np.array_equal(observed, expect)
    back = unpack_observations(obs, idx, cnt)
    assert np.array_equal(np.isnan(back), ~expect)
    assert np.array_equal(back[expect], y[expect])
