#!/bin/bash
# One gpurun call: tests, smoke, bench (both arms), ncu launch list + full capture.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_tail.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/${TAG}_bench_reference.json
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/bench.err; head -c 1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/bench.err
# launch list of the same command, timed region only (cudaProfilerStart/Stop in bench.py)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-other > gpurun_out/ncu_bench.log 2>&1
grep -c "" gpurun_out/${TAG}_launches.csv
# full capture of the three launches of one step
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dfm_ -s 9 -c 3 \
    -o gpurun_out/${TAG}_prof_dfm python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-other > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
