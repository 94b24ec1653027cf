#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$1" > gpurun_out/quick.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/quick.log
