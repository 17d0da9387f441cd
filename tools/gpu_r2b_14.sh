mkdir -p gpurun_out
timeout 600 python tools/probe_variants.py metran_b200/lib/variants/lib_info.so 2>&1 | tail -4 | tee gpurun_out/r02w_probe_info.txt
timeout 600 python tools/probe_c1.py 1024 2000 metran_b200/lib/variants/lib_info.so 2>&1 | tail -4 | tee -a gpurun_out/r02w_probe_info.txt
timeout 600 python tools/probe_variants.py 37888 5000 0.02 metran_b200/lib/variants/lib_info.so 2>&1 | tail -3 | tee -a gpurun_out/r02w_probe_info.txt
