mkdir -p gpurun_out
timeout 300 ./tools/ubench_tma_store > gpurun_out/r02_tma_store.txt 2>&1; cat gpurun_out/r02_tma_store.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dfm_loglik -s 1 -c 1 -o gpurun_out/r02i_k1h python tools/profile_hist.py 37888 96 > gpurun_out/r02i_k1h.log 2>&1; tail -3 gpurun_out/r02i_k1h.log
