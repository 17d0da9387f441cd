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
for _ in range(2):
    out = eng.dfm_filter_history(y, phi, q, G)
torch.cuda.synchronize()
