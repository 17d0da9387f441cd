mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hist_stores.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/probe_hist.py 37888 96 metran_b200/lib/variants/lib_hist_nbuf1.so 2>&1 | tail -8 | tee gpurun_out/r02j_probe_hist.txt
