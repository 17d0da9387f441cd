import os, sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
from oracle import kalman_oracle as ko
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
try: print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("cpu.max n/a", e)
print("loadavg", open("/proc/loadavg").read().strip())
os.system("lscpu | grep -E 'Model name|Socket|Core|Thread|MHz' | head -8")
rng = np.random.default_rng(0)
B = 2048; T = 2000
alpha, gamma = bench.make_params(rng, B, 5, 3)
y = bench.simulate_cpu(rng, alpha, gamma, T, 0.0)
mats = [ko.dfm_matrices(alpha[b], gamma[b]) for b in range(B)]
A, Q, H, R = (np.stack([m[i] for m in mats]) for i in range(4))
for th in (1, 2, 8, 32, 64, 128):
    n = min(B, 16 * th)
    ko.filter_loglik_batch(y[:th], A[:th], Q[:th], H[:th], R[:th], threads=th)
    t = time.perf_counter(); ko.filter_loglik_batch(y[:n], A[:n], Q[:n], H[:n], R[:n], threads=th); dt = time.perf_counter() - t
    print(th, "threads:", "%.2f M ts/s total, %.3f per thread" % (n * T / dt / 1e6, n * T / dt / th / 1e6), flush=True)
