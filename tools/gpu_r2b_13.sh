mkdir -p gpurun_out
timeout 600 python tools/probe_sw.py 256 1000 metran_b200/lib/variants/lib_sw_pw4.so 2>&1 | tail -3 | tee gpurun_out/r02v_probe_sw.txt
timeout 600 python tools/probe_sw.py 1024 300 metran_b200/lib/variants/lib_sw_pw4.so 2>&1 | tail -3 | tee -a gpurun_out/r02v_probe_sw.txt
