"""Time fit_dfm_batch on the headline batch (not a test)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from metran_b200.batched import KalmanEngine
from metran_b200.fit import fit_dfm_batch
eng = KalmanEngine(); rng = np.random.default_rng(1)
B, T, m, k = 37888, 5000, 5, 3
alpha, gamma = bench.make_params(rng, B, m, k)
y = bench.simulate_gpu(torch, eng.device, 5, eng.tensor(alpha), eng.tensor(gamma), T, 0.0)
torch.cuda.synchronize()
t = time.perf_counter(); res = fit_dfm_batch(eng, y, gamma); dt = time.perf_counter() - t
at_truth, _, _ = eng.dfm_loglik_alpha(y, alpha, gamma)
print("fit_dfm_batch: %d models x T=%d: %.2f s, %d launches (objective evaluations per model), converged %.1f %%, median iterations %d"
      % (B, T, dt, res.nfev, 100 * res.converged.mean(), int(np.median(res.nit))))
print("mean (-2logL at fit) - (-2logL at generating alpha): %.3f" % float(np.mean(res.obj - at_truth.cpu().numpy())))
print("median |log(alpha_fit/alpha_true)|: %.3f" % float(np.median(np.abs(np.log(res.alpha / alpha)))))
