"""GPU timing probe of the generic filter / history / smoother kernels (not a test)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from metran_b200.batched import KalmanEngine
eng = KalmanEngine(); dev = eng.device
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for (B, T, n, m) in ((256, 1000, 32, 64), (4096, 100, 32, 64), (2048, 200, 16, 12), (16384, 500, 8, 5)):
    g = torch.Generator(device=dev); g.manual_seed(2)
    y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
    phi = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 0.2 + 0.75
    A = torch.diag_embed(phi); Q = torch.diag_embed(1 - phi**2)
    H = torch.randn(B, m, n, dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
    R = torch.full((B, m), 0.1, dtype=torch.float64, device=dev)
    t1 = timeit(lambda: eng.filter_loglik(y, A, Q, H, R, flags=3))
    out = eng.filter_history(y, A, Q, H, R, flags=3)
    t2 = timeit(lambda: eng.filter_history(y, A, Q, H, R, flags=3))
    t3 = timeit(lambda: eng.smoother(out['fm'], out['fc'], out['pm'], out['pc'], A, flags=3))
    hb = B * T * 8 * (m + 2 * (n + n * n) + 3); sb = B * T * 8 * 3 * (n + n * n)
    print(f"n={n} m={m} B={B} T={T}: loglik {t1:.2f} ms ({B*T/t1*1e3:.2e}/s) | history {t2:.2f} ms ({B*T/t2*1e3:.2e}/s, {hb/t2/1e6:.0f} GB/s) | smoother {t3:.2f} ms ({B*T/t3*1e3:.2e}/s, {sb/t3/1e6:.0f} GB/s)", flush=True)
    del out
# thread-per-model history kernel for Metran-shaped models
for (B, T) in ((16384, 500), (23680, 400), (28416, 400), (37888, 300)):
    m, k = 5, 3; n = 8
    g = torch.Generator(device=dev); g.manual_seed(3)
    y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
    phi = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 0.2 + 0.75
    q = 1 - phi**2
    G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
    t4 = timeit(lambda: eng.dfm_filter_history(y, phi, q, G))
    hb = B * T * 8 * (m + 2 * (n + n * n) + 2.5)
    print(f"DFM thread-per-model history n=8 m=5 B={B} T={T}: {t4:.2f} ms ({B*T/t4*1e3:.2e}/s, {hb/t4/1e6:.0f} GB/s)", flush=True)
# pure-write and copy bandwidth references (torch kernels)
buf = torch.empty(2 * 1024**3 // 8 * 4, dtype=torch.float64, device=dev)   # 8 GB
tz = timeit(lambda: buf.zero_())
print(f"torch zero_ 8 GB: {tz:.2f} ms -> {buf.numel()*8/tz/1e6:.0f} GB/s write", flush=True)
src = torch.empty_like(buf)
tc = timeit(lambda: buf.copy_(src))
print(f"torch copy_ 8 GB: {tc:.2f} ms -> {2*buf.numel()*8/tc/1e6:.0f} GB/s read+write", flush=True)
