#!/bin/bash
# ncu launch list + full captures of the conv kernels (never a bench value: numbers printed under ncu are discarded)
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --sample-every 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mid_tc -s 20 -c 2 -f -o gpurun_out/prof_hidden_$TAG $CMD > gpurun_out/ncu_hidden_$TAG.log 2>&1; echo "hidden exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_last -s 2 -c 1 -f -o gpurun_out/prof_last_$TAG $CMD > gpurun_out/ncu_last_$TAG.log 2>&1; echo "last exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_first -s 2 -c 1 -f -o gpurun_out/prof_first_$TAG $CMD > gpurun_out/ncu_first_$TAG.log 2>&1; echo "first exit $?"
ls -la gpurun_out
