#!/bin/bash
# native backward: tests, training-step timing (native vs autograd/cuDNN backward), ncu launch list of the backward kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu > gpurun_out/pytest_bwd.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_bwd.log | cut -c 1-400
for d in ffdnet SimpleCNN; do
  timeout 600 python scripts/bench_train.py --denoiser $d --steps 6 --warmup 4 2>/dev/null | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$d native   ', round(d['ms_per_step'],2), 'ms')"
  DEQSCI_NATIVE_BACKWARD=0 timeout 600 python scripts/bench_train.py --denoiser $d --steps 6 --warmup 4 2>/dev/null | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$d autograd ', round(d['ms_per_step'],2), 'ms')"
done
for d in SimpleCNN ffdnet; do timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad|act_bwd" -c 80 --csv --log-file gpurun_out/bwd_launches_$d.csv python scripts/bench_train.py --denoiser $d --steps 1 --warmup 0 > /dev/null 2>&1; echo "$d exit $?"; python - <<PY
import csv,collections
rows=[l for l in open("gpurun_out/bwd_launches_$d.csv") if l.startswith(chr(34))]
agg=collections.OrderedDict()
for r in csv.DictReader(rows):
    n=r["Kernel Name"].split("(")[0]; v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    v = v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=0
for k,v in agg.items(): print("%-40s n=%3d total %8.1f us avg %8.1f us"%(k,v[0],v[1],v[1]/v[0])); tot+=v[1]
print("total %.1f us"%tot)
PY
done
