#!/bin/bash
mkdir -p gpurun_out
B=${1:-8}
shift
timeout 900 python bench.py --steps 1 --warmup 1 --batch $B --no-cpu-baseline "$@" > gpurun_out/bench_probe.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_probe.log') if x.startswith('{')]
if not l:
    print(open('gpurun_out/bench_probe.log').read()[-3000:])
    raise SystemExit
d=json.loads(l[-1])
print("value %.3f recon/s  e2e %.3f  ms/step %.1f  launches %d"%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches']))
print("roofline", d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
tot=0
for k,v in d['kernels'].items():
    print("%-16s n=%5d avg_ms=%s est_ms_per_step=%s"%(k, v['launches'], v['avg_ms'], v['est_ms_per_step']))
    tot+= v['est_ms_per_step'] or 0
print("sum of kernels %.1f ms of %.1f"%(tot, d['ms_per_step']))
print(d['clocks'], d['check'])
PY
