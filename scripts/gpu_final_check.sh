#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_default.log | cut -c 1-300
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "reference exit $?"; tail -n 1 gpurun_out/bench_reference.log | cut -c 1-200
