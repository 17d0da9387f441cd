"""Time mkf_dfm_filter_history (thread-per-model history kernel) on B200: TMA tensor stores vs the
shared-memory transposition, optionally for variant libraries (metran_b200/lib/variants/lib_*.so).

    python tools/probe_hist.py [B T] [lib ...]
"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metran_b200 import _lib
dev = torch.device("cuda", 0)
nums = [a for a in sys.argv[1:] if not a.endswith(".so")]
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
B, T = (int(nums[0]), int(nums[1])) if len(nums) >= 2 else (37888, 96)
m, k = 5, 3; n = m + k
g = torch.Generator(device=dev); g.manual_seed(3)
y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
phi = torch.rand(B, n, dtype=torch.float64, device=dev, generator=g) * 0.2 + 0.75
q = 1 - phi ** 2
G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
E = lambda *s, dtype=torch.float64: torch.empty(*s, dtype=dtype, device=dev)
out = dict(sigma=E(B, T), detf=E(B, T), cnt=E(B, T, dtype=torch.int32), fm=E(B, T, n), fc=E(B, T, n, n), pm=E(B, T, n),
           pc=E(B, T, n, n), mle=E(B), nobs=E(B, dtype=torch.int64), status=E(B, dtype=torch.int32))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
bytes_per_unit = 8 * (2 * (n + n * n) + 2) + 4 + 8 * m
ref = None
for path in [_lib.LIB_PATH] + libs:
    L = ctypes.CDLL(path)
    res, args = _lib.SIGNATURES["mkf_dfm_filter_history"]
    L.mkf_dfm_filter_history.restype = res; L.mkf_dfm_filter_history.argtypes = args
    def run(o=out):
        rc = L.mkf_dfm_filter_history(B, T, m, k, P(y), T * m, P(phi), n, P(q), n, P(G), m * k, None, 0, 1,
                                      P(o["sigma"]), P(o["detf"]), P(o["cnt"]), P(o["fm"]), P(o["fc"]), P(o["pm"]), P(o["pc"]),
                                      P(o["mle"]), P(o["nobs"]), P(o["status"]), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    for flag in (1, 0):
        L.mkf_dfm_set_hist_tma(flag)
        for key in ("fc", "pc", "fm", "pm"): out[key].zero_()
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        chk = float(out["fc"].sum() + out["pc"].sum() + out["fm"].sum() + out["pm"].sum() + out["sigma"].sum() + out["detf"].sum()) + float(out["cnt"].sum())
        if ref is None: ref = chk
        ms = min(ts)
        print("%-24s tma=%d  %.3f ms  %.3e units/s  %.0f GB/s   checksum diff %.3e" % (os.path.basename(path), flag, ms, B * T / ms * 1e3, B * T * bytes_per_unit / ms / 1e6, chk - ref), flush=True)
    # covariances only / no scalars
    o2 = dict(out); o2["sigma"] = o2["detf"] = o2["cnt"] = None
    L.mkf_dfm_set_hist_tma(1)
    run(o2); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(o2); e1.record(); torch.cuda.synchronize()
    print("%-24s tma=1 without the step scalars  %.3f ms" % (os.path.basename(path), e0.elapsed_time(e1)), flush=True)
