"""Build libmetran_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

No torch / pybind dependency: the library is plain CUDA runtime (statically linked)
behind the C ABI declared in include/metran_b200.h.  ``python -m metran_b200.build``.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "lib", "obj")
LIB = os.path.join(PKG, "lib", "libmetran_b200.so")

SOURCES = [
    "api.cu",
    "filter_generic.cu",
    "filter_dfm.cu",
    "filter_dfm_k0.cu",
    "filter_dfm_k1.cu",
    "filter_dfm_k2.cu",
    "filter_dfm_k3.cu",
    "smoother.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(p):
        raise RuntimeError("nvcc not found; libmetran_b200.so cannot be built")
    return p


def _deps_mtime():
    heads = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    heads.append(os.path.join(os.path.dirname(PKG), "include", "metran_b200.h"))
    return max(os.path.getmtime(h) for h in heads)


def _compile(src, extra):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if (os.path.isfile(obj) and os.path.getmtime(obj) >= os.path.getmtime(path)
            and os.path.getmtime(obj) >= _deps_mtime()):
        return obj, ""
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False, extra_flags=()):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, extra_flags), SOURCES))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        with open(os.path.join(PKG, "lib", "ptxas.log"), "w") as fh:
            fh.write(log)
    if verbose and log:
        sys.stderr.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc_path(), "-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
