"""C1 (BASELINE configs[1]: 1024 models, T = 2000) and small batches: per-launch times of the three-launch
path, the single-kernel path, for the product library and variant libraries.

    python tools/probe_c1.py [B T] [lib ...]
"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metran_b200 import _lib
import bench
dev = torch.device("cuda", 0)
nums = [a for a in sys.argv[1:] if not a.endswith(".so")]
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
B, T = (int(nums[0]), int(nums[1])) if len(nums) >= 2 else (1024, 2000)
m, k = 5, 3
rng = np.random.default_rng(1000)
alpha_h, gamma_h = bench.make_params(rng, B, m, k)
alpha = torch.as_tensor(alpha_h, device=dev); G = torch.as_tensor(gamma_h, device=dev)
y = bench.simulate_gpu(torch, dev, 2000, alpha, G, T, 0.0)
mle = torch.empty(B, dtype=torch.float64, device=dev); nobs = torch.empty(B, dtype=torch.int64, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
ref = None
for path in [_lib.LIB_PATH] + libs:
    L = ctypes.CDLL(path)
    for name in ("mkf_dfm_loglik_alpha", "mkf_dfm_last_profile"):
        res, args = _lib.SIGNATURES[name]
        getattr(L, name).restype = res; getattr(L, name).argtypes = args
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def run():
        rc = L.mkf_dfm_loglik_alpha(B, T, m, k, P(y), T * m, P(alpha), m + k, 1.0, P(G), m * k, 1, P(mle), P(nobs), P(st), stream)
        assert rc == 0, rc
    for flag in (1, 0):
        L.mkf_dfm_set_handoff(flag)
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(9):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        out = mle.cpu().numpy().copy()
        if ref is None: ref = out
        line = "%-22s B=%d T=%d handoff=%d  min %.3f ms  %.3e units/s  max|d| %.1e" % (os.path.basename(path), B, T, flag, min(ts), B * T / min(ts) * 1e3, float(np.nanmax(np.abs(out - ref))))
        if flag:
            L.mkf_dfm_set_profile(1)
            best = None
            for _ in range(5):
                run()
                ms = (ctypes.c_float * 3)(); stt = (ctypes.c_int64 * 3)()
                rc = L.mkf_dfm_last_profile(stream, ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(stt, ctypes.c_void_p))
                if rc == 0 and ms[0] >= 0 and (best is None or sum(ms) < sum(best[0])): best = (list(ms), list(stt))
            L.mkf_dfm_set_profile(0)
            if best:
                line += "  launches %.3f / %.3f / %.3f ms, warp-steps after hand-off %d (%.0f per warp)" % (best[0][0], best[0][1], best[0][2], best[1][2], best[1][2] / max(best[1][1], 1))
        print(line, flush=True)
    L.mkf_release_workspaces()
