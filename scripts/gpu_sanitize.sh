#!/bin/bash
# compute-sanitizer over the CUDA-core kernels and one small tensor-core reconstruction
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gap_kernels or anderson_kernels or residual_kernel" > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool (cuda-core kernels) exit $?"
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke exit $?"
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/sanitizer_memcheck_smoke.log | tail -3
# the tensor-core kernels on wide images (CTA-pair hidden layers, first / last layers, PDL chain) and the
# train-mode statistics epilogue
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "wide_images or train_mode_batchnorm or plan_refresh" > gpurun_out/sanitizer_memcheck_tc.log 2>&1; echo "memcheck tensor-core kernels exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_tc.log | tail -3
# round 2: the fused scale + Adam kernel (single rank) and the per-sample residual path of the solve kernel
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/fused_adam_check.py > gpurun_out/sanitizer_memcheck_adam.log 2>&1; echo "memcheck fused adam exit $?"
grep -E "ERROR SUMMARY|FUSED_ADAM" gpurun_out/sanitizer_memcheck_adam.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/fused_adam_check.py > gpurun_out/sanitizer_racecheck_adam.log 2>&1; echo "racecheck fused adam exit $?"
grep -E "ERROR SUMMARY|FUSED_ADAM" gpurun_out/sanitizer_racecheck_adam.log | tail -3
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stops_per_measurement or admm or driver_equals" > gpurun_out/sanitizer_memcheck_r2.log 2>&1; echo "memcheck round-2 paths exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_r2.log | tail -3
