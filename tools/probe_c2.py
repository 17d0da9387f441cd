"""C2 (BASELINE configs[2]): 8192 parameter draws of the reference's example data on the compacted time grid."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from metran_b200.batched import KalmanEngine
eng = KalmanEngine(); dev = eng.device
g21 = np.load(os.path.join(ROOT, "tests", "golden", "b21b0214.npz"))
y21 = torch.as_tensor(g21["y"], device=dev); f21 = torch.as_tensor(g21["factors"], device=dev)
draws = torch.as_tensor(np.random.default_rng(5).uniform(1.0, 100.0, (8192, y21.shape[1] + f21.shape[1])), device=dev)
yc, ti, J, frac = eng.compact_observations(y21)
ms_full = bench.dev_time(torch, lambda: eng.dfm_loglik_alpha(y21, draws, f21))
ms_c = bench.dev_time(torch, lambda: eng.dfm_loglik_compact(yc, ti, f21, alpha=draws), reps=9)
d = float((eng.dfm_loglik_alpha(y21, draws, f21)[0] - eng.dfm_loglik_compact(yc, ti, f21, alpha=draws)[0]).abs().max())
one = draws[:1].contiguous()
ms_1 = bench.dev_time(torch, lambda: eng.dfm_loglik_compact(yc, ti, f21, alpha=one), reps=9)
print("C2: compact %.4f ms (J = %d of %d steps), full grid %.3f ms, max|d| compact vs full %.2e; one draw: %.4f ms" % (ms_c, J, y21.shape[0], ms_full, d, ms_1))
