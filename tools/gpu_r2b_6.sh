mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_smoother_thread.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/profile_smoother8.py 37888 96 2>&1 | tail -2 | tee gpurun_out/r02n_smoother8.txt
timeout 300 python tools/profile_smoother8.py 8192 500 2>&1 | tail -2 | tee -a gpurun_out/r02n_smoother8.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smoother_thread -s 1 -c 1 -o gpurun_out/r02n_smt python tools/profile_smoother8.py 37888 96 > gpurun_out/r02n_smt.log 2>&1; tail -2 gpurun_out/r02n_smt.log
