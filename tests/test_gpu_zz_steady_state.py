"""Steady-state step of the thread-per-model kernel against the oracle,
for every collapse-capable (m, k) shape: T = 2000 fully observed steps, long enough for whole warps
to switch, with gaps that undo the switch.  The body (tests/steady_state_cases.py) also runs on the
CPU through the oracle-backed stand-in library (tests/test_host.py).  Runs last in the GPU suite.
"""
import pytest

import steady_state_cases as ssc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from metran_b200.batched import KalmanEngine

    return KalmanEngine()


@pytest.mark.parametrize("m,k", ssc.SHAPES)
def test_dfm_steady_state_vs_oracle(eng, m, k):
    ssc.check(eng, m, k, lambda t: t.detach().cpu().numpy())
