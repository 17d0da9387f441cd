mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hist_stores.py tests/test_gpu_kernels.py tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/probe_hist.py 37888 96 2>&1 | tail -4 | tee gpurun_out/r02l_probe_hist.txt
timeout 300 python tools/probe_hist.py 37888 250 2>&1 | tail -4 | tee -a gpurun_out/r02l_probe_hist.txt
timeout 300 python tools/probe_hist.py 4096 2000 2>&1 | tail -4 | tee -a gpurun_out/r02l_probe_hist.txt
