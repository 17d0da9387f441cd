mkdir -p gpurun_out
timeout 300 tools/ubench_rowstream 2>&1 | tee gpurun_out/r02_rowstream.txt
timeout 600 python tools/probe_variants.py 2>&1 | tail -8
timeout 900 python bench.py --steps 10 > gpurun_out/r02d_bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/r02d_bench.json; tail -5 gpurun_out/bench.err
