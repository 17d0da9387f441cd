"""Launch the C3-shaped filter + smoother once or twice (for ncu captures)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from metran_b200 import _lib
from metran_b200.batched import KalmanEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
eng = KalmanEngine()
y, A, Q, H, R = bench.c3_data(torch, eng.device, B, T, 32, 64)
for _ in range(2):
    out = eng.filter_history(y, A, Q, H, R, flags=3)
    eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], A, flags=3)
torch.cuda.synchronize()
