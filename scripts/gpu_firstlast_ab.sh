#!/bin/bash
# first / last layer experiments: accumulator ring depth of the last layer (DEQSCI_TCL_BUFS)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -2
for rep in 1 2; do
for NB in 8 4; do
  DEQSCI_TCL_BUFS=$NB timeout 900 python bench.py --steps 2 --warmup 2 --batch 32 --no-cpu-baseline > gpurun_out/ab.log 2>&1
  python - "$NB" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/ab.log') if x.startswith('{')]
d=json.loads(l[-1]); k=d['kernels']
print("TCL_BUFS=%s: value %.3f ms/step %.1f first %.4f last %.4f hidden %.4f clocks %s" % (sys.argv[1], d['value'], d['ms_per_step'], k['conv_first']['avg_ms'], k['conv_last']['avg_ms'], k['conv_hidden']['avg_ms'], d['clocks']['sm_mhz']))
PY
done
done
