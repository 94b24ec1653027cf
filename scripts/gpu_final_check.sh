#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), the bench with the driver's step counts, the reference arm
mkdir -p gpurun_out
TAG=${1:-final}
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
fi
S=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_$TAG.log 2> gpurun_out/bench_reference_$TAG.err; echo "reference exit $? wall $(( $(date +%s) - S )) s"; tail -n 1 gpurun_out/bench_reference_$TAG.log | cut -c 1-200
S=$(date +%s); timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default_$TAG.log 2> gpurun_out/bench_default_$TAG.err; echo "bench exit $? wall $(( $(date +%s) - S )) s"; tail -n 1 gpurun_out/bench_default_$TAG.log | cut -c 1-300
