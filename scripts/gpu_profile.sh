#!/bin/bash
# ncu launch list + full captures of the conv kernels (never a bench value: numbers printed under ncu are discarded)
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --sample-every 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_hidden_2cta -s 20 -c 2 -f -o gpurun_out/prof_hidden_$TAG $CMD > gpurun_out/ncu_hidden_$TAG.log 2>&1; echo "hidden exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_last_tc -s 2 -c 1 -f -o gpurun_out/prof_last_$TAG $CMD > gpurun_out/ncu_last_$TAG.log 2>&1; echo "last exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_first_tc -s 2 -c 1 -f -o gpurun_out/prof_first_$TAG $CMD > gpurun_out/ncu_first_$TAG.log 2>&1; echo "first exit $?"
timeout 900 ncu --set full --clock-control none -k regex:anderson_gram -s 4 -c 1 -f -o gpurun_out/prof_gram_$TAG $CMD > gpurun_out/ncu_gram_$TAG.log 2>&1; echo "gram exit $?"
timeout 900 ncu --set full --clock-control none -k regex:anderson_mix -s 4 -c 1 -f -o gpurun_out/prof_mix_$TAG $CMD > gpurun_out/ncu_mix_$TAG.log 2>&1; echo "mix exit $?"
timeout 900 ncu --set full --clock-control none -k regex:gap_prep -s 2 -c 1 -f -o gpurun_out/prof_prep_$TAG $CMD > gpurun_out/ncu_prep_$TAG.log 2>&1; echo "prep exit $?"
timeout 1200 python bench.py > gpurun_out/bench_default_$TAG.log 2>&1; echo "bench default exit $?"; tail -n 1 gpurun_out/bench_default_$TAG.log | cut -c 1-600
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_$TAG.log 2>&1; echo "bench reference exit $?"; tail -n 1 gpurun_out/bench_reference_$TAG.log | cut -c 1-400
ls -la gpurun_out | tail -20
