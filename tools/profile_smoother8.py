"""ncu / timing target: thread-per-model history (n = 8) followed by the RTS smoother on its output."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from metran_b200.batched import KalmanEngine
from metran_b200 import _lib
eng = KalmanEngine(); dev = eng.device
B, T, m, k = int(sys.argv[1]), int(sys.argv[2]), 5, 3; n = 8
libs = [a for a in sys.argv[3:] if a.endswith(".so")]
g = torch.Generator(device=dev); g.manual_seed(3)
y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
phi = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 0.2 + 0.75
q = 1 - phi**2
G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
out = eng.dfm_filter_history(y, phi, q, G)
A = torch.diag_embed(phi)
ref = None
for path in [None] + libs:
    if path is not None:
        eng.lib = _lib.load(path)
    for thread_kernel in (1, 0):
        prev = eng.lib.mkf_smoother_set_thread_min_batch(1 if thread_kernel else 1 << 40)
        ts = []
        for _ in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); sm, sc, st = eng.smoother(out["fm"], out["fc"], out["pm"], out["pc"], A, flags=3); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        eng.lib.mkf_smoother_set_thread_min_batch(prev)
        ms = min(ts[1:])
        chk = float(sm.sum() + sc.sum())
        if ref is None: ref = chk
        print("%-22s smoother n=8 B=%d T=%d %s: %.3f ms  %.3e units/s  %.0f GB/s (1728 B/unit)  checksum diff %.2e" % (
            "product" if path is None else path.split("/")[-1], B, T, "thread-per-model" if thread_kernel else "lane-per-row   ", ms,
            B * T / ms * 1e3, B * T * 1728 / ms / 1e6, chk - ref), flush=True)
        if path is not None and not thread_kernel: break
