#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1; echo "bench exit $?"; tail -5 gpurun_out/bench_b8.log
