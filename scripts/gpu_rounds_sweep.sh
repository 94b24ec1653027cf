#!/bin/bash
# strip-height heuristic of the pair kernel at small batches: DEQSCI_TC_ROUNDS sweep
mkdir -p gpurun_out
for B in ${BATCHES:-1 2 4}; do
  for R in ${ROUNDS:-1 2 3 6}; do
    DEQSCI_TC_ROUNDS=$R timeout 900 python bench.py --steps 2 --warmup 2 --batch $B --no-cpu-baseline > gpurun_out/ab.log 2>&1
    python - "$B" "$R" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/ab.log') if x.startswith('{')]
d=json.loads(l[-1])
print("batch %s ROUNDS=%s: value %.3f ms/step %.1f" % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']))
PY
  done
done
