#!/bin/bash
# ncu --set full of the pair kernel at the benchmarked batch, one capture per issue mode (DEQSCI_TC_RS)
mkdir -p gpurun_out
B=${1:-32}
for M in ${2:-0 2}; do
  DEQSCI_TC_RS=$M timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_hidden_2cta -s 20 -c 2 -f -o gpurun_out/prof_hidden_rs$M python bench.py --steps 1 --warmup 1 --batch $B --profile-mode --sample-every 0 > gpurun_out/ncu_hidden_rs$M.log 2>&1; echo "mode $M exit $?"
done
ls -la gpurun_out | grep prof_hidden_rs
