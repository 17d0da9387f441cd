"""Wide smoother (n = 32) timing for the product library and variant libraries on the C3 shape."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from metran_b200 import _lib
from metran_b200.batched import KalmanEngine
B = int(sys.argv[1]); T = int(sys.argv[2]); libs = [a for a in sys.argv[3:] if a.endswith(".so")]
eng = KalmanEngine()
y, A, Q, H, R = bench.c3_data(torch, eng.device, B, T, 32, 64)
out = eng.filter_history(y, A, Q, H, R, flags=3)
ref = None
for path in [None] + libs:
    if path: eng.lib = _lib.load(path)
    ts = []
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sm, sc, st = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], A, flags=3); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    if ref is None: ref = (sm.clone(), sc.clone())
    print("%-20s n=32 B=%d T=%d smoother %.2f ms  max|d sm| %.1e max|d sc| %.1e status %d" % (
        "product" if path is None else path.split("/")[-1], B, T, min(ts[1:]), float((sm - ref[0]).abs().max()),
        float((sc - ref[1]).abs().max()), int(st.abs().sum())), flush=True)
