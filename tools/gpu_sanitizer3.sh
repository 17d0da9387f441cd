# compute-sanitizer over the two smoother kernels as they are in the final tree: K3T with the 73-KB layout
# (single-buffered P_f / x_f / x_p boxes refilled at the end of the step) and K3W with two Gauss-Jordan
# pivots per barrier.  Small shapes.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r02w_sanitizer_$tool.log \
      python -m pytest tests/test_gpu_smoother_thread.py tests/test_gpu_kernels.py -m gpu -q -x \
      -k "(smooth and not 203 and not 6400) or singular" > gpurun_out/r02w_sanitizer_${tool}_pytest.txt 2>&1
  tail -2 gpurun_out/r02w_sanitizer_${tool}_pytest.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02w_sanitizer_$tool.log | tail -2
  grep -E "Race reported|Error:" gpurun_out/r02w_sanitizer_$tool.log | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -6
done
