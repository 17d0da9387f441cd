"""Timing of the single-model drop-in paths on the C0 data set (not a test)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_model
from metran_b200.batched import KalmanEngine
from metran_b200.dfm import DynamicFactorModel
from metran_b200.kalmanfilter import SPKalmanFilter
from oracle import kalman_oracle as ko
g = load_model("b21b0214"); eng = KalmanEngine()
mdl = DynamicFactorModel(g["y"], g["factors"], engine=eng)
p = g["alpha_opt"]
mdl.get_mle(p); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(50): mdl.get_mle(p)
print("DynamicFactorModel.get_mle (K1, y resident): %.3f ms/call" % ((time.perf_counter() - t) / 50 * 1e3))
t = time.perf_counter()
for _ in range(20): mdl.get_mle_batch(np.tile(p, (7, 1)))
print("get_mle_batch(7 FD points): %.3f ms/call" % ((time.perf_counter() - t) / 20 * 1e3))
t = time.perf_counter(); mdl.solve(); print("solve() batched FD: %.3f s, nfev %d, launches %s" % (time.perf_counter() - t, mdl.parameters["nfev"], mdl.parameters["batched_launches"]))
t = time.perf_counter(); mdl.solve(batched=False); print("solve() unbatched: %.3f s, nfev %d" % (time.perf_counter() - t, mdl.parameters["nfev"]))
kf = SPKalmanFilter(device_engine=eng); kf.set_observations(g["y"]); kf.set_matrices(*mdl._get_matrices(p))
kf.run_filter(); kf.get_mle()
t = time.perf_counter()
for _ in range(20): kf.run_filter(); kf.get_mle()
print("SPKalmanFilter.run_filter+get_mle (facade path, lazy history): %.3f ms/call" % ((time.perf_counter() - t) / 20 * 1e3))
t = time.perf_counter()
for _ in range(5): kf.run_smoother(); _ = kf.smoothed_state_means
print("run_smoother + smoothed means to host: %.3f ms/call" % ((time.perf_counter() - t) / 5 * 1e3))
t = time.perf_counter(); _ = kf.smoothed_state_covariances; print("covariances to host: %.3f ms" % ((time.perf_counter() - t) * 1e3))
A, Q, H, R = mdl._get_matrices(p)
t = time.perf_counter()
for _ in range(20): ko.filter_loglik(g["y"], A, Q, H, R)
print("CPU oracle (C port) pack+filter+get_mle: %.3f ms/call" % ((time.perf_counter() - t) / 20 * 1e3))
# C2: 8192 draws of this model
draws = np.random.default_rng(0).uniform(1, 100, (8192, 6))
d = eng.tensor(draws); mdl_y = mdl.kf._y_dev[0]; G = eng.tensor(g["factors"])
eng.dfm_loglik_alpha(mdl_y, d, G); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); eng.dfm_loglik_alpha(mdl_y, d, G); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("C2: 8192 draws x T=6255 (real C0 data, 94.5%% empty days): %.3f ms -> %.3e ts*models/s" % (ms, 8192 * 6255 / ms * 1e3))
# K2 (warp-group) vs K1 (thread-per-model) at tiny batch sizes on the C0 data
import torch
for Bt in (1, 7, 28, 64):
    P = np.tile(p, (Bt, 1)) * (1 + 0.01 * np.arange(Bt))[:, None]
    mats = [mdl._get_matrices(pp) for pp in P]
    A4, Q4, H4, R4 = (eng.tensor(np.stack([mm[i] for mm in mats])) for i in range(4))
    Pd = eng.tensor(P)
    def t_of(fn):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / 10
    k1 = t_of(lambda: eng.dfm_loglik_alpha(mdl_y, Pd, G))
    k2 = t_of(lambda: eng.filter_loglik(mdl_y, A4, Q4, H4, R4, flags=3))
    print("B=%d on C0 data: K1 %.3f ms, K2 %.3f ms (kernel only)" % (Bt, k1, k2))
