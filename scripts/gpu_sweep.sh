#!/bin/bash
mkdir -p gpurun_out
for cfg in "$@"; do
  IFS=: read B R <<< "$cfg"
  DEQSCI_TC_ROUNDS=$R timeout 600 python bench.py --steps 1 --warmup 1 --batch $B --no-cpu-baseline > gpurun_out/sweep_${B}_${R}.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/sweep_${B}_${R}.log') if l.startswith('{')][-1])
print('B=$B rounds=$R value %.3f ms/step %.1f hidden_ms %.4f frac %.3f sm_mhz %s'%(d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms'],d['roofline']['frac'],d['clocks']['sm_mhz']))
PY
done
