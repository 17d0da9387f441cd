import sys
import numpy as np, torch
sys.path.insert(0, ".")
from metran_b200.batched import KalmanEngine
eng = KalmanEngine(); dev = eng.device
B, T, n, m = int(sys.argv[1]), int(sys.argv[2]), 32, 64
g = torch.Generator(device=dev); g.manual_seed(2)
y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
phi = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 0.2 + 0.75
A = torch.diag_embed(phi); Q = torch.diag_embed(1 - phi**2)
H = torch.randn(B, m, n, dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
R = torch.full((B, m), 0.1, dtype=torch.float64, device=dev)
for _ in range(2):
    eng.filter_loglik(y, A, Q, H, R, flags=3)
torch.cuda.synchronize()
