#!/bin/bash
mkdir -p gpurun_out
for d in SimpleCNN RealSN_SimpleCNN; do
  timeout 900 python bench.py --denoiser $d --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > gpurun_out/side_$d.log 2>&1; echo "$d exit $?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/side_$d.log') if l.startswith('{')][-1])
print(d['metric']); print('value %.3f e2e %.3f ms/step %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d['clocks'])
print('roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), d['roofline']['avg_launch_ms'])
for k,v in d['kernels'].items(): print('  %-16s n=%5d avg_ms=%s est_ms_per_step=%s'%(k, v['launches'], v['avg_ms'], v['est_ms_per_step']))
PY
done
