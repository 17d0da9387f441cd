"""Time mkf_dfm_loglik_alpha of every K1 variant library (metran_b200/lib/variants) on B200."""
import ctypes, glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metran_b200 import _lib
dev = torch.device("cuda", 0)
B, T, m, k = 37888, 5000, 5, 3
g = torch.Generator(device=dev); g.manual_seed(1)
y = torch.randn(B, T, m, dtype=torch.float64, device=dev, generator=g)
alpha = torch.rand(B, m + k, dtype=torch.float64, device=dev, generator=g) * 45 + 5
G = (torch.rand(B, m, k, dtype=torch.float64, device=dev, generator=g) * 0.6 + 0.3) / np.sqrt(k)
mle = torch.empty(B, dtype=torch.float64, device=dev); nobs = torch.empty(B, dtype=torch.int64, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
ref = None
for path in (sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "metran_b200", "lib", "variants", "lib_*.so")))):
    L = ctypes.CDLL(path)
    res, args = _lib.SIGNATURES["mkf_dfm_loglik_alpha"]
    L.mkf_dfm_loglik_alpha.restype = res; L.mkf_dfm_loglik_alpha.argtypes = args
    def run():
        rc = L.mkf_dfm_loglik_alpha(B, T, m, k, P(y), T * m, P(alpha), m + k, 1.0, P(G), m * k, 1, P(mle), P(nobs), P(st),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    out = mle.cpu().numpy().copy()
    if ref is None: ref = out
    print("%-22s min %.3f ms med %.3f ms  max|d mle| vs first %.2e" % (os.path.basename(path), min(ts), float(np.median(ts)), float(np.nanmax(np.abs(out - ref)))), flush=True)
