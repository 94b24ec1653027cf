#!/bin/bash
# ncu launch list + full captures of the conv kernels (never a bench value: numbers printed under ncu are discarded)
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline --sample-every 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_hidden_2cta -s 20 -c 2 -f -o gpurun_out/prof_hidden_$TAG $CMD > gpurun_out/ncu_hidden_$TAG.log 2>&1; echo "hidden exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_mid_tc -s 2 -c 1 -f -o gpurun_out/prof_last_$TAG $CMD > gpurun_out/ncu_last_$TAG.log 2>&1; echo "last exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_first -s 2 -c 1 -f -o gpurun_out/prof_first_$TAG $CMD > gpurun_out/ncu_first_$TAG.log 2>&1; echo "first exit $?"
timeout 900 ncu --set full --clock-control none -k regex:anderson_gram -s 4 -c 1 -f -o gpurun_out/prof_gram_$TAG $CMD > gpurun_out/ncu_gram_$TAG.log 2>&1; echo "gram exit $?"
timeout 900 ncu --set full --clock-control none -k regex:anderson_mix -s 4 -c 1 -f -o gpurun_out/prof_mix_$TAG $CMD > gpurun_out/ncu_mix_$TAG.log 2>&1; echo "mix exit $?"
for B in 16 32; do
  timeout 900 python bench.py --steps 1 --warmup 1 --batch $B --no-cpu-baseline > gpurun_out/bench_b$B.log 2>&1; echo "bench B=$B exit $?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_b$B.log') if l.startswith('{')][-1])
print('B=$B value %.3f e2e %.3f ms/step %.1f hidden_ms %.4f frac %.3f clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['frac'],d['clocks']))
"
done
ls -la gpurun_out | tail -20
