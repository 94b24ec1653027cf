#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "hidden_layer and tc_split" > gpurun_out/pair_test.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/pair_test.log
