#!/bin/bash
# ncu launch list + --set full captures of every kernel of the chain at the BENCHMARKED batch (never a bench value:
# numbers printed under ncu are discarded).  usage: gpu_profile.sh <tag> [batch]
mkdir -p gpurun_out
TAG=${1:-r2}
B=${2:-32}
CMD="python bench.py --steps 1 --warmup 1 --batch $B --profile-mode --sample-every 0"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list exit $?"
cap() {  # cap <name> <kernel regex> <skip> <count>
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_$1_$TAG $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1; echo "$1 exit $?"
}
cap hidden conv_hidden_2cta 20 2
cap last conv_last_tc 2 1
cap first conv_first_tc 2 1
cap gram anderson_gram 4 1
cap solve anderson_solve 4 1
cap mix anderson_mix 4 1
cap prep gap_prep 2 1
ls -la gpurun_out | grep $TAG
