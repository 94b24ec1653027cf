#!/bin/bash
mkdir -p gpurun_out
B=${1:-8}
timeout 900 python bench.py --steps 1 --warmup 1 --batch $B --no-cpu-baseline > gpurun_out/bench_probe.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_probe.log') if x.startswith('{')]
d=json.loads(l[-1])
print("value %.3f recon/s  e2e %.3f  ms/step %.1f"%(d['value'],d['e2e']['value'],d['ms_per_step']))
print("roofline", d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
for k,v in d['kernels'].items(): print(k, v)
print(d['clocks'])
PY
