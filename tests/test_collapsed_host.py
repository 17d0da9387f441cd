"""CPU check of the collapsed (factor-only) dense step used by the DFM kernel.

metran_b200/csrc/dfm_collapsed.cuh holds the arithmetic as __host__ __device__ functions;
tests/native/host_collapsed.cu wraps it with the kernel's mode logic for ONE model on the
host.  Here that build is compared with the oracle (restatement of
metran/kalmanfilter.py:315-390, 550-567) over all supported shapes and over missing-data
patterns that force every collapse / expand transition.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import kalman_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "host_collapsed.cu")
OUT = os.path.join(HERE, "native", "_build", "libhost_collapsed.so")
HDR = os.path.join(os.path.dirname(HERE), "metran_b200", "csrc", "dfm_collapsed.cuh")


@pytest.fixture(scope="module")
def hostlib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    stale = (not os.path.isfile(OUT)) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if stale:
        subprocess.run([nvcc, "-w", "-O2", "-std=c++17", "--extended-lambda", "--expt-relaxed-constexpr",
                        "-Xcompiler", "-fPIC", "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-cudart", "static", "-o", OUT, SRC], check=True)
    lib = ctypes.CDLL(OUT)
    lib.host_dfm_loglik.restype = ctypes.c_double
    return lib


def _run(lib, y, phi, q, G, warmup, use_collapse):
    T, m = y.shape
    k = G.shape[1]
    y = np.ascontiguousarray(y, dtype=np.float64)
    G = np.ascontiguousarray(G, dtype=np.float64)
    nc = ctypes.c_int64(0)
    P = ctypes.c_void_p
    v = lib.host_dfm_loglik(m, k, ctypes.c_int64(T), P(y.ctypes.data), P(phi.ctypes.data), P(q.ctypes.data),
                            P(G.ctypes.data), warmup, use_collapse, ctypes.byref(nc))
    return v, nc.value


def _model(rng, m, k):
    G = rng.uniform(-0.5, 0.5, (m, k)) / np.sqrt(k)
    phi = np.exp(-1.0 / rng.uniform(2, 60, m + k))
    comm = (G ** 2).sum(1)
    q = np.concatenate([(1 - phi[:m] ** 2) * (1 - comm), 1 - phi[m:] ** 2])
    return G, phi, q


def _pattern(rng, y, pat):
    T, m = y.shape
    if pat == "blocks":
        for s in range(20, T, 57):
            y[s:s + rng.integers(1, 9)] = np.nan
    elif pat == "scatter":
        y[rng.random((T, m)) < 0.1] = np.nan
    elif pat == "rowgaps":
        y[rng.random(T) < 0.3] = np.nan
        y[5, 0] = -1e10            # the reference's sentinel for "dropped" (kalmanfilter.py:666)
    elif pat == "t0missing":
        y[0] = np.nan
    return y


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("pat", ["dense", "blocks", "scatter", "rowgaps", "t0missing"])
def test_collapsed_matches_oracle(hostlib, k, pat):
    rng = np.random.default_rng(100 * k + len(pat))
    for m in range(1, 9):
        T = 300
        G, phi, q = _model(rng, m, k)
        y = _pattern(rng, rng.standard_normal((T, m)).cumsum(0) * 0.1, pat)
        H = np.concatenate([np.eye(m), G], axis=1)
        for warmup in (1, 3):
            ref, _ = ko.filter_loglik(y, np.diag(phi), np.diag(q), H, np.zeros(m), warmup=warmup)
            got, ncol = _run(hostlib, y, phi, q, G, warmup, 1)
            plain, _ = _run(hostlib, y, phi, q, G, warmup, 0)
            assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref)), (m, k, pat, got, ref)
            assert abs(plain - ref) <= 1e-11 * max(1.0, abs(ref))
            if pat == "dense":
                assert ncol == T - 1          # every step after the first runs factor-only
            if pat == "t0missing":
                assert ncol == T - 2


@pytest.mark.parametrize("m,k", [(5, 3), (5, 1), (3, 2), (1, 1), (6, 2), (4, 3)])
def test_steady_state_freeze_matches_oracle(hostlib, m, k):
    """Mode 3 = the kernel's steady-state switch: once the factor covariance has stopped changing
    the gains of one more full step are captured and later steps only touch the data.  Long series
    so that the switch happens; gaps so that it is undone (constants restored, state expanded) and
    happens again."""
    rng = np.random.default_rng(7 * m + k)
    T = 2400
    G, phi, q = _model(rng, m, k)
    H = np.concatenate([np.eye(m), G], axis=1)
    x = np.zeros(m + k)
    y = np.empty((T, m))
    for t in range(T):
        x = phi * x + np.sqrt(q) * rng.standard_normal(m + k)
        y[t] = H @ x
    for pat in ("dense", "gaps"):
        yy = y.copy()
        if pat == "gaps":
            yy[1500:1503] = np.nan
            yy[2000, 0] = np.nan
        ref, _ = ko.filter_loglik(yy, np.diag(phi), np.diag(q), H, np.zeros(m), warmup=1)
        got, nfrozen = _run(hostlib, yy, phi, q, G, 1, 3)
        plain, _ = _run(hostlib, yy, phi, q, G, 1, 1)
        assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref)), (pat, got, ref)
        assert abs(got - plain) <= 1e-12 * max(1.0, abs(ref))
        assert nfrozen > T // 4, (pat, nfrozen)     # the switch did happen


def test_collapsed_golden_dense(hostlib):
    """The committed fixture generated by the unmodified reference (tests/golden/make_golden.py)."""
    from conftest import load_cases
    c = load_cases()["dfm_n8_m5_dense"]
    H = c["H"]
    m = H.shape[0]
    phi = np.ascontiguousarray(np.diag(c["A"]))
    q = np.ascontiguousarray(np.diag(c["Q"]))
    got, ncol = _run(hostlib, c["y"], phi, q, H[:, m:], 1, 1)
    assert ncol == c["y"].shape[0] - 1
    assert abs(got - float(c["mle"])) <= 1e-9


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("pat", ["dense", "blocks", "scatter", "rowgaps", "t0missing"])
def test_information_form_matches_oracle(hostlib, k, pat):
    """Mode 5 = factor-only steps in information form (collapsed_info_step: K pivots in the factor space
    instead of M in the joint space).  Only sigma and det S of a step are reproduced, which is all -2 log L
    needs; every shape the kernel compiles, every gap pattern."""
    rng = np.random.default_rng(300 * k + len(pat))
    worst = 0.0
    for m in range(1, 9):
        T = 300
        G, phi, q = _model(rng, m, k)
        y = _pattern(rng, rng.standard_normal((T, m)).cumsum(0) * 0.1, pat)
        H = np.concatenate([np.eye(m), G], axis=1)
        for warmup in (1, 3):
            ref, _ = ko.filter_loglik(y, np.diag(phi), np.diag(q), H, np.zeros(m), warmup=warmup)
            got, ncol = _run(hostlib, y, phi, q, G, warmup, 5)
            elim, _ = _run(hostlib, y, phi, q, G, warmup, 1)
            worst = max(worst, abs(got - ref) / max(1.0, abs(ref)))
            assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref)), (m, k, pat, got, ref)
            assert abs(got - elim) <= 1e-11 * max(1.0, abs(ref))
            if pat == "dense":
                assert ncol == T - 1
    assert worst < 1e-11


@pytest.mark.parametrize("m,k", [(5, 3), (5, 1), (3, 2), (1, 1), (6, 2), (4, 3), (2, 3), (1, 3)])
def test_information_form_with_steady_state_switch(hostlib, m, k):
    """Mode 7 = information-form transient + the steady-state switch (capture step in the elimination form).
    Shapes with m < k have a rank-deficient N = Dt^T Dt (zero rows in its factor R)."""
    rng = np.random.default_rng(11 * m + k)
    T = 2400
    G, phi, q = _model(rng, m, k)
    H = np.concatenate([np.eye(m), G], axis=1)
    x = np.zeros(m + k)
    y = np.empty((T, m))
    for t in range(T):
        x = phi * x + np.sqrt(q) * rng.standard_normal(m + k)
        y[t] = H @ x
    for pat in ("dense", "gaps"):
        yy = y.copy()
        if pat == "gaps":
            yy[1500:1503] = np.nan
            yy[2000, 0] = np.nan
        ref, _ = ko.filter_loglik(yy, np.diag(phi), np.diag(q), H, np.zeros(m), warmup=1)
        got, nfrozen = _run(hostlib, yy, phi, q, G, 1, 7)
        elim, _ = _run(hostlib, yy, phi, q, G, 1, 3)
        assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref)), (pat, got, ref)
        assert abs(got - elim) <= 1e-11 * max(1.0, abs(ref))
        assert nfrozen > T // 4, (pat, nfrozen)


def test_information_form_equal_decay_rates(hostlib):
    """phi_s == phi_c for every state: D = 0, N = 0, the innovation carries no information about the factor
    error; and the metric workload's parameter range at T = 5000."""
    rng = np.random.default_rng(5)
    m, k, T = 5, 3, 400
    G = rng.uniform(0.3, 0.9, (m, k)) / np.sqrt(k)
    phi = np.full(m + k, 0.9)
    comm = (G ** 2).sum(1)
    q = np.concatenate([(1 - phi[:m] ** 2) * (1 - comm), 1 - phi[m:] ** 2])
    y = rng.standard_normal((T, m))
    H = np.concatenate([np.eye(m), G], axis=1)
    ref, _ = ko.filter_loglik(y, np.diag(phi), np.diag(q), H, np.zeros(m), warmup=1)
    got, _ = _run(hostlib, y, phi, q, G, 1, 5)
    assert abs(got - ref) <= 1e-11 * abs(ref)
    worst = 0.0
    for s in range(6):
        r2 = np.random.default_rng(100 + s)
        alpha = r2.uniform(5, 50, m + k)
        G = r2.uniform(0.3, 0.9, (m, k)) / np.sqrt(k)
        phi = np.exp(-1 / alpha)
        comm = (G ** 2).sum(1)
        q = np.concatenate([(1 - phi[:m] ** 2) * (1 - comm), 1 - phi[m:] ** 2])
        H = np.concatenate([np.eye(m), G], axis=1)
        x = np.zeros(m + k)
        y = np.empty((5000, m))
        for t in range(5000):
            x = phi * x + np.sqrt(q) * r2.standard_normal(m + k)
            y[t] = H @ x
        ref, _ = ko.filter_loglik(y, np.diag(phi), np.diag(q), H, np.zeros(m), warmup=1)
        got, _ = _run(hostlib, y, phi, q, G, 1, 7)
        worst = max(worst, abs(got - ref))
    assert worst <= 5e-10, worst          # absolute: the kernel's contract is 1e-9
