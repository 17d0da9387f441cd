mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoother_thread.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/profile_smoother8.py 37888 96 2>&1 | tail -2 | tee gpurun_out/r02m_smoother8.txt
timeout 300 python tools/profile_smoother8.py 8192 500 2>&1 | tail -2 | tee -a gpurun_out/r02m_smoother8.txt
timeout 300 python tools/probe_c1.py 1024 2000 metran_b200/lib/variants/lib_st_s6.so metran_b200/lib/variants/lib_st_s9.so metran_b200/lib/variants/lib_tr_s3.so 2>&1 | tail -10 | tee gpurun_out/r02m_probe_c1.txt
timeout 300 python tools/probe_c1.py 32 2000 2>&1 | tail -3 | tee -a gpurun_out/r02m_probe_c1.txt
