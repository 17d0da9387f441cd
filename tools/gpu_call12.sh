mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --no-cpu > gpurun_out/r02h_bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])
for k,v in d['other_configs'].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ('ms','value','ms_full_grid','observed_steps','max_abs_dmle_compact_vs_full_grid','ms_filter_history','ms_smoother')}, 'frac', v.get('roofline',{}).get('frac'))
PY
