"""Build K1 variants into metran_b200/lib/variants/*.so (compile-time switches) for A/B probes."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metran_b200 import build
VARIANTS = {
    # transient kernel: factor-only steps in information form
    "info": ["-DMKF_INFO=1"],
}
if len(sys.argv) > 1:
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in sys.argv[1:]}
out = os.path.join(ROOT, "metran_b200", "lib", "variants")
os.makedirs(out, exist_ok=True)
srcs = list(build.SOURCES)
for name, flags in VARIANTS.items():
    objs = []
    for s in srcs:
        o = os.path.join(out, "%s_%s.o" % (name, s[:-3]))
        cmd = [build.nvcc_path()] + build.NVCC_FLAGS + ["-DMKF_DFM_ONLY_M5"] + flags + ["-c", os.path.join(build.CSRC, s), "-o", o]
        subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
        objs.append(o)
    # stubs for K = 0..2 launchers are not needed: link only what the probe uses (K = 3)
    lib = os.path.join(out, "lib_%s.so" % name)
    subprocess.check_call([build.nvcc_path(), "-shared", "-o", lib] + objs + ["-cudart", "static"])
    print(name, os.path.getsize(lib))
