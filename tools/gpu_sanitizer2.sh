# compute-sanitizer memcheck + racecheck over the kernels that changed in the second half of round 2:
# K1H (TMA tensor stores from swizzled boxes, 16-step scalar boxes, sector stores), K3T (TMA boxes in and
# out, two warps per box set), K3W (Gauss-Jordan inverse, padded staging).  Small shapes.
mkdir -p gpurun_out
SEL='tests/test_gpu_hist_stores.py tests/test_gpu_smoother_thread.py tests/test_gpu_kernels.py'
K='(hist or history or smooth or unaligned or oracle or golden) and not 203 and not 1500'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/r02s_sanitizer_$tool.log \
      python -m pytest $SEL -m gpu -q -x -k "$K" > gpurun_out/r02s_sanitizer_${tool}_pytest.txt 2>&1
  tail -3 gpurun_out/r02s_sanitizer_${tool}_pytest.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/r02s_sanitizer_$tool.log | sort | uniq -c | sort -rn | head -8
done
