"""Launch the DFM kernel a few times on random data (for ncu captures)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from metran_b200.batched import KalmanEngine
B = int(sys.argv[1]); T = int(sys.argv[2]); miss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
m, k = 5, 3
eng = KalmanEngine(); dev = eng.device
g = torch.Generator(device=dev); g.manual_seed(1)
y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
if miss > 0:
    y[torch.rand(B, T, m, device=dev, generator=g) < miss] = float("nan")
alpha = torch.rand(B, m + k, dtype=torch.float64, device=dev, generator=g) * 45 + 5
G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
for _ in range(3):
    eng.dfm_loglik_alpha(y, alpha, G)
torch.cuda.synchronize()
