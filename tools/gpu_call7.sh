mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"filter_block|smoother" -s 2 -c 2 -o gpurun_out/r02g_prof_c3 python tools/profile_c3.py 256 400 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
