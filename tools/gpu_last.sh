mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"smoother_wide" -c 1 -o gpurun_out/r02z_c3sw python tools/profile_c3.py 256 300 > gpurun_out/r02z_c3.log 2>&1; tail -2 gpurun_out/r02z_c3.log
