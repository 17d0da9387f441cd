mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_compact.py tests/test_call_trace.py tests/test_numpy_engine.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
from metran_b200.batched import KalmanEngine
eng = KalmanEngine()
e = bench.c3_config(torch, eng, 6489.3, 36.7, with_cpu=False)
print({k: e[k] for k in ("ms", "ms_filter_history", "ms_smoother", "ms_filter_loglik_only")}, e["roofline"]["frac"])
PY
