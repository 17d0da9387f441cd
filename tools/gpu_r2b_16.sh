mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02y_pytest_tail.txt
timeout 600 python tests/probe_c3.py 2>&1 | head -4 | tee gpurun_out/r02y_probe_c3.txt
