mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/r02a_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dfm_loglik_kernel -s 3 -c 1 -o gpurun_out/r02a_prof_dfm python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
