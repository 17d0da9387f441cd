mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_handoff.py tests/test_gpu_zz_steady_state.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/probe_variants.py 2>&1 | tail -5
timeout 600 python tools/probe_variants.py 1024 2000 0 2>&1 | tail -4
timeout 600 python tools/probe_variants.py 131072 10000 0 2>&1 | tail -4
