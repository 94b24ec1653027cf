#!/bin/bash
# experiment: does zeroing low mantissa bits of the lo' operands lower tensor-core power (higher clock)?
mkdir -p gpurun_out
for rep in 1 2; do
for M in 65535 65520 65472 0; do
  DEQSCI_TC_LO_MASK=$M timeout 900 python bench.py --steps 2 --warmup 2 --batch 32 --no-cpu-baseline > gpurun_out/ab.log 2>&1
  python - "$M" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/ab.log') if x.startswith('{')]
d=json.loads(l[-1])
print("lo_mask %5s (0x%04X): value %.3f ms/step %.1f hidden %.4f ms clocks %s check %s" % (sys.argv[1], int(sys.argv[1]), d['value'], d['ms_per_step'], d['kernels']['conv_hidden']['avg_ms'], d['clocks']['sm_mhz'], d.get('check')))
PY
done
done
