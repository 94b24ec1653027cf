#!/bin/bash
# compute-sanitizer memcheck over the paths that run the row-stationary pair kernel (227 KB of shared memory exactly)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke exit $?"
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/sanitizer_memcheck_smoke.log | tail -3
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "hidden_layer or wide_images or train_mode_batchnorm or plan_refresh or driver_equals or masked_adjoint" > gpurun_out/sanitizer_memcheck_tc.log 2>&1; echo "memcheck tensor-core kernels exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_tc.log | tail -3
