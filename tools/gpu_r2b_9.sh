mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"smoother_wide|filter_block" -c 2 -o gpurun_out/r02q_c3 python tools/profile_c3.py 256 300 > gpurun_out/r02q_c3.log 2>&1; tail -3 gpurun_out/r02q_c3.log
