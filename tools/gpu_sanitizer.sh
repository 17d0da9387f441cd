# compute-sanitizer memcheck + racecheck over the kernels with asynchronous staging / shared-memory
# exchange: K1 (TMA ring, mbarriers, hand-off), K1S (steady kernel), K1H (history staging), K2 / K2 block /
# K3 / K3 wide (shared-memory broadcasts, block barriers).  Small shapes: the tools slow kernels 10-100x.
mkdir -p gpurun_out
SEL='test_gpu_kernels.py tests/test_gpu_handoff.py tests/test_gpu_compact.py tests/test_gpu_abi_edges.py'
K='not actually_taken and not 1500'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log \
      python -m pytest tests/$SEL -m gpu -q -x -k "$K" > gpurun_out/r02_sanitizer_${tool}_pytest.txt 2>&1
  tail -3 gpurun_out/r02_sanitizer_${tool}_pytest.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/r02_sanitizer_$tool.log | tail -5
done
