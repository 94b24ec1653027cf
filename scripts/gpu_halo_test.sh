#!/bin/bash
mkdir -p gpurun_out
for bo in 0 1; do
  echo "=== BASEOFF=$bo"
  DEQSCI_TC_BASEOFF=$bo timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hidden_layer and tc_split" > gpurun_out/halo_bo$bo.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/halo_bo$bo.log
done
