mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_host.py -m gpu -x -q -k "smooth or c3 or wide or n32 or n20 or run_smoother or simul or decompose" 2>&1 | tail -6
timeout 600 python tests/probe_c3.py 2>&1 | head -4 | tee gpurun_out/r02q_probe_c3.txt
