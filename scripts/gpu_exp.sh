#!/bin/bash
mkdir -p gpurun_out
for v in ${@:-0 1}; do
  DEQSCI_TC_DEBUG_SKIP_STORE=$v timeout 600 python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > gpurun_out/exp_$v.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/exp_$v.log') if l.startswith('{')][-1])
print('skip_store=$v value %.3f hidden_ms %.4f sm_mhz %s'%(d['value'],d['roofline']['avg_launch_ms'],d['clocks']['sm_mhz']))
PY
done
