#!/bin/bash
# launch list + ncu --set full of the hidden-layer kernel only (the other kernels' captures are reused): usage <tag> [batch]
mkdir -p gpurun_out
TAG=${1:-r2h}
B=${2:-32}
CMD="python bench.py --steps 1 --warmup 1 --batch $B --profile-mode --sample-every 0"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_hidden_2cta -s 20 -c 2 -f -o gpurun_out/prof_hidden_$TAG $CMD > gpurun_out/ncu_hidden_$TAG.log 2>&1; echo "hidden exit $?"
