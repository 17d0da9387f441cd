"""ncu / timing target: thread-per-model history (n = 8) followed by the RTS smoother on its output."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from metran_b200.batched import KalmanEngine
eng = KalmanEngine(); dev = eng.device
B, T, m, k = int(sys.argv[1]), int(sys.argv[2]), 5, 3; n = 8
g = torch.Generator(device=dev); g.manual_seed(3)
y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
phi = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 0.2 + 0.75
q = 1 - phi**2
G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
out = eng.dfm_filter_history(y, phi, q, G)
A = torch.diag_embed(phi)
ts = []
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); sm, sc, st = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], A, flags=3); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = min(ts)
print("smoother n=8 B=%d T=%d: %.3f ms  %.3e units/s  %.0f GB/s (1728 B/unit)" % (B, T, ms, B * T / ms * 1e3, B * T * 1728 / ms / 1e6))
