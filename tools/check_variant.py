"""Run ONE K1 variant library on a small dense + gappy batch and compare with the oracle.
Separate process per variant: a faulting kernel poisons the CUDA context.
    python tools/check_variant.py metran_b200/lib/variants/lib_x.so
"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metran_b200 import _lib
from oracle import kalman_oracle as ko
path = sys.argv[1]
L = ctypes.CDLL(path)
res, args = _lib.SIGNATURES["mkf_dfm_loglik_alpha"]
L.mkf_dfm_loglik_alpha.restype = res; L.mkf_dfm_loglik_alpha.argtypes = args
L.mkf_last_cuda_error.restype = ctypes.c_char_p
dev = torch.device("cuda", 0)
rng = np.random.default_rng(3)
B, T, m, k = 70, 57, 5, 3
n = m + k
y = rng.standard_normal((B, T, m)).cumsum(1) * 0.1
y[40:, 20:23] = np.nan
y[33, 7, 2] = np.nan
alpha = rng.uniform(5, 50, (B, n))
G = rng.uniform(0.3, 0.9, (B, m, k)) / np.sqrt(k)
phi = np.exp(-1.0 / alpha)
comm = (G ** 2).sum(2)
q = np.concatenate([(1 - phi[:, :m] ** 2) * (1 - comm), 1 - phi[:, m:] ** 2], axis=1)
A = np.zeros((B, n, n)); Q = np.zeros((B, n, n)); H = np.zeros((B, m, n))
i = np.arange(n)
A[:, i, i] = phi; Q[:, i, i] = q; H[:, np.arange(m), np.arange(m)] = 1; H[:, :, m:] = G
want, _ = ko.filter_loglik_batch(y, A, Q, H, np.zeros((B, m)))
ty, ta, tg = (torch.as_tensor(v, device=dev) for v in (y, alpha, G))
mle = torch.empty(B, dtype=torch.float64, device=dev); nobs = torch.empty(B, dtype=torch.int64, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
rc = L.mkf_dfm_loglik_alpha(B, T, m, k, P(ty), T * m, P(ta), n, 1.0, P(tg), m * k, 1, P(mle), P(nobs), P(st),
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
try:
    torch.cuda.synchronize()
    err = float(np.max(np.abs(mle.cpu().numpy() - want)))
    print(os.path.basename(path), "rc", rc, "max|mle - oracle| = %.3e" % err)
except Exception as e:
    print(os.path.basename(path), "rc", rc, "FAULT:", str(e).splitlines()[0], L.mkf_last_cuda_error())
