"""ctypes binding of libmetran_b200.so (C ABI: include/metran_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at first use.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmetran_b200.so")

MKF_OK, MKF_ERR_BAD_ARG, MKF_ERR_UNSUPPORTED, MKF_ERR_CUDA = 0, 1, 2, 3
STATUS_NONPOSITIVE_F, STATUS_CHOL_FAILED = 1, 2
FLAG_A_DIAGONAL, FLAG_Q_DIAGONAL = 1, 2

_i64, _i32, _u32, _dbl, _ptr = (ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32,
                                ctypes.c_double, ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/metran_b200.h one to one
SIGNATURES = {
    "mkf_version": (ctypes.c_int, []),
    "mkf_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "mkf_last_cuda_error": (ctypes.c_char_p, []),
    "mkf_limits": (None, [_ptr, _ptr, _ptr]),
    "mkf_filter_history": (ctypes.c_int,
                           [_i64, _i64, _i32, _i32] + [_ptr, _i64] * 7 + [_i32, _u32]
                           + [_ptr] * 10 + [_ptr]),
    "mkf_filter_loglik": (ctypes.c_int,
                          [_i64, _i64, _i32, _i32] + [_ptr, _i64] * 7 + [_i32, _u32]
                          + [_ptr] * 3 + [_ptr]),
    "mkf_dfm_loglik": (ctypes.c_int,
                       [_i64, _i64, _i32, _i32] + [_ptr, _i64] * 5 + [_i32] + [_ptr] * 3 + [_ptr]),
    "mkf_dfm_filter_history": (ctypes.c_int,
                               [_i64, _i64, _i32, _i32] + [_ptr, _i64] * 5 + [_i32] + [_ptr] * 10 + [_ptr]),
    "mkf_dfm_loglik_alpha": (ctypes.c_int,
                             [_i64, _i64, _i32, _i32, _ptr, _i64, _ptr, _i64, _dbl, _ptr, _i64, _i32]
                             + [_ptr] * 3 + [_ptr]),
    "mkf_dfm_loglik_compact": (ctypes.c_int,
                               [_i64, _i64, _i32, _i32, _ptr, _i64, _ptr, _i64, _ptr, _i64, _dbl]
                               + [_ptr, _i64] * 4 + [_i32] + [_ptr] * 3 + [_ptr]),
    "mkf_smoother": (ctypes.c_int, [_i64, _i64, _i32] + [_ptr] * 4 + [_ptr, _i64, _u32]
                     + [_ptr] * 3 + [_ptr]),
    "mkf_smoother_project": (ctypes.c_int, [_i64, _i64, _i32] + [_ptr] * 4 + [_ptr, _i64, _u32]
                             + [_i32, _ptr, _i64] + [_ptr] * 6 + [_ptr]),
    "mkf_project": (ctypes.c_int, [_i64, _i64, _i32, _i32, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr]),
    "mkf_dfm_set_handoff": (ctypes.c_int, [ctypes.c_int]),
    "mkf_dfm_set_hist_tma": (ctypes.c_int, [ctypes.c_int]),
    "mkf_smoother_set_thread_min_batch": (_i64, [_i64]),
    "mkf_release_workspaces": (None, []),
    "mkf_dfm_set_profile": (ctypes.c_int, [ctypes.c_int]),
    "mkf_dfm_last_profile": (ctypes.c_int, [_ptr, _ptr, _ptr]),
    "mkf_reduce_sum": (ctypes.c_int, [_ptr, _i64, _ptr, _ptr]),
    "mkf_probe_fp64": (ctypes.c_int, [_i32, _i32, _ptr, _ptr]),
}

_lib = None


class MetranB200Error(RuntimeError):
    pass


def load(path=None):
    """Load the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise MetranB200Error(
            "libmetran_b200.so not found at %s -- build it with `python -m metran_b200.build` "
            "(there is no CPU fallback)" % p)
    L = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = L
    return L


def check(lib, code, what):
    if code == MKF_OK:
        return
    msg = lib.mkf_error_string(code)
    msg = msg.decode() if isinstance(msg, bytes) else str(msg)
    if code == MKF_ERR_CUDA:
        extra = lib.mkf_last_cuda_error()
        msg += ": " + (extra.decode() if isinstance(extra, bytes) else str(extra))
    raise MetranB200Error("%s failed: %s" % (what, msg))


def limits(lib=None):
    lib = lib or load()
    a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    lib.mkf_limits(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return {"n_max": a.value, "dfm_m_max": b.value, "dfm_k_max": c.value}
