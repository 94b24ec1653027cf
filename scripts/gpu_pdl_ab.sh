#!/bin/bash
# A/B of programmatic dependent launch on the hidden kernel (DEQSCI_TC_PDL), alternating runs
mkdir -p gpurun_out
for rep in 1 2; do
for B in ${BATCHES:-8 32}; do
  for P in 1 0; do
    DEQSCI_TC_PDL=$P timeout 900 python bench.py --steps 2 --warmup 2 --batch $B --no-cpu-baseline > gpurun_out/ab.log 2>&1
    python - "$B" "$P" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/ab.log') if x.startswith('{')]
d=json.loads(l[-1])
print("batch %s PDL=%s: value %.3f e2e %.3f ms/step %.1f hidden avg %.4f ms clocks %s" % (sys.argv[1], sys.argv[2], d['value'], d['e2e']['value'], d['ms_per_step'], d['kernels']['conv_hidden']['avg_ms'], d['clocks']['sm_mhz']))
PY
  done
done
done
