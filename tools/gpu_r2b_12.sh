mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoother_thread.py tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/profile_smoother8.py 37888 96 metran_b200/lib/variants/lib_smt_full.so 2>&1 | tail -4 | tee gpurun_out/r02t_smoother8.txt
timeout 300 python tools/profile_smoother8.py 8192 500 metran_b200/lib/variants/lib_smt_full.so 2>&1 | tail -4 | tee -a gpurun_out/r02t_smoother8.txt
timeout 300 python tools/profile_smoother8.py 2048 500 2>&1 | tail -2 | tee -a gpurun_out/r02t_smoother8.txt
timeout 300 python tools/profile_smoother8.py 4096 500 2>&1 | tail -2 | tee -a gpurun_out/r02t_smoother8.txt
