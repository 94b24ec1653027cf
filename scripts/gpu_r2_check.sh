#!/bin/bash
# round-2 check: GPU test suite, then the default bench (all blocks) and the reference arm
mkdir -p gpurun_out
TAG=${1:-r2a}
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_default_$TAG.log 2>&1; echo "bench default exit $?"; tail -n 1 gpurun_out/bench_default_$TAG.log | cut -c 1-3000; tail -n 5 gpurun_out/bench_default_$TAG.log | grep -v '^{' | tail -n 4
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_$TAG.log 2>&1; echo "bench reference exit $?"; tail -n 1 gpurun_out/bench_reference_$TAG.log | cut -c 1-1500
