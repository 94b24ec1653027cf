#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
bash scripts/gpu_bench_probe.sh ${1:-8}
