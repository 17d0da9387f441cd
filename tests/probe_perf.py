"""Quick GPU timing probe (not a test, not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from metran_b200.batched import KalmanEngine
eng = KalmanEngine()
dev = eng.device
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
m, k = 5, 3; n = m + k
for B, T in ((1024, 2000), (32768, 5000), (37888, 5000), (18944, 5000), (75776, 2000)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
    alpha = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 45 + 5
    G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
    tmin, tmed = timeit(lambda: eng.dfm_loglik_alpha(y, alpha, G))
    print(f"dfm_alpha B={B} T={T}: min {tmin:.3f} ms med {tmed:.3f} ms -> {B*T/tmin*1e3:.3e} ts*models/s, y {B*T*m*8/tmin/1e6:.1f} GB/s", flush=True)
    if B <= 8192:
        phi = torch.exp(-1/alpha); A = torch.diag_embed(phi); Q = torch.diag_embed(1-phi**2)
        H = torch.cat([torch.eye(m, dtype=torch.float64, device=dev).expand(B, m, m), G], dim=2).contiguous()
        R = torch.zeros(B, m, dtype=torch.float64, device=dev)
        tmin, tmed = timeit(lambda: eng.filter_loglik(y, A, Q, H, R, flags=3))
        print(f"generic  B={B} T={T}: min {tmin:.3f} ms -> {B*T/tmin*1e3:.3e} ts*models/s", flush=True)
    del y

# missing-data path
B, T = 32768, 2000
g = torch.Generator(device=dev); g.manual_seed(1)
y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
y[torch.rand(B, T, m, device=dev, generator=g) < 0.5] = float("nan")
alpha = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 45 + 5
G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
tmin, tmed = timeit(lambda: eng.dfm_loglik_alpha(y, alpha, G))
print(f"dfm_alpha 50% missing B={B} T={T}: min {tmin:.3f} ms -> {B*T/tmin*1e3:.3e} ts*models/s", flush=True)
