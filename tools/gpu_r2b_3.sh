mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hist_stores.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/probe_hist.py 37888 96 metran_b200/lib/variants/lib_hist_c1m1.so metran_b200/lib/variants/lib_hist_c1m4.so metran_b200/lib/variants/lib_hist_c2m2.so metran_b200/lib/variants/lib_hist_c2m4.so 2>&1 | tail -16 | tee gpurun_out/r02k_probe_hist.txt
timeout 300 python tools/probe_hist.py 37888 250 2>&1 | tail -4 | tee -a gpurun_out/r02k_probe_hist.txt
timeout 300 python tools/profile_smoother8.py 37888 96 2>&1 | tail -2 | tee gpurun_out/r02k_smoother8.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smoother_kernel -s 1 -c 1 -o gpurun_out/r02k_smoother8 python tools/profile_smoother8.py 37888 96 > gpurun_out/r02k_smoother8.log 2>&1; tail -2 gpurun_out/r02k_smoother8.log
