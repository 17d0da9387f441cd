#!/bin/bash
# One gpurun call: tests, smoke, bench (both arms), ncu launch list + full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
# launch list of the same command (short run)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
grep -c "" gpurun_out/launches.csv
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dfm_loglik_kernel -s 3 -c 1 \
    -o gpurun_out/prof_dfm python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
