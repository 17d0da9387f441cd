mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_handoff.py tests/test_gpu_zz_steady_state.py tests/test_numpy_engine.py -x -q 2>&1 | tail -8
timeout 600 python tools/probe_variants.py 2>&1 | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dfm_ -s 6 -c 3 -o gpurun_out/r02c_prof_k1 python tools/probe_variants.py 37888 5000 0 metran_b200/lib/libmetran_b200.so > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
