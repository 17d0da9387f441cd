"""Time mkf_dfm_loglik_alpha of every K1 variant library (metran_b200/lib/variants) on B200.

    python tools/probe_variants.py [B T [miss]] [lib ...]
"""
import ctypes, glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metran_b200 import _lib
import bench
dev = torch.device("cuda", 0)
nums = [a for a in sys.argv[1:] if not a.endswith(".so")]
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
B, T = (int(nums[0]), int(nums[1])) if len(nums) >= 2 else (37888, 5000)
miss = float(nums[2]) if len(nums) >= 3 else 0.0
m, k = 5, 3
rng = np.random.default_rng(1000)
alpha_h, gamma_h = bench.make_params(rng, B, m, k)
alpha = torch.as_tensor(alpha_h, device=dev); G = torch.as_tensor(gamma_h, device=dev)
y = bench.simulate_gpu(torch, dev, 2000, alpha, G, T, miss)
mle = torch.empty(B, dtype=torch.float64, device=dev); nobs = torch.empty(B, dtype=torch.int64, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
ref = None
paths = libs or ([_lib.LIB_PATH] + sorted(glob.glob(os.path.join(ROOT, "metran_b200", "lib", "variants", "lib_*.so"))))
for path in paths:
    L = ctypes.CDLL(path)
    res, args = _lib.SIGNATURES["mkf_dfm_loglik_alpha"]
    L.mkf_dfm_loglik_alpha.restype = res; L.mkf_dfm_loglik_alpha.argtypes = args
    def run():
        rc = L.mkf_dfm_loglik_alpha(B, T, m, k, P(y), T * m, P(alpha), m + k, 1.0, P(G), m * k, 1, P(mle), P(nobs), P(st),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    for flag in ((1, 0) if path == paths[0] else (1,)):
        L.mkf_dfm_set_handoff(flag)
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        out = mle.cpu().numpy().copy()
        if ref is None: ref = out
        print("%-28s handoff=%d min %.3f ms med %.3f ms  max|d mle| vs first %.2e  nan %d" % (os.path.basename(path), flag, min(ts), float(np.median(ts)), float(np.nanmax(np.abs(out - ref))), int(np.isnan(out).sum())), flush=True)
    L.mkf_release_workspaces()
