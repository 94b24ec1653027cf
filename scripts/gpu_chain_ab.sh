#!/bin/bash
# experiment: all hidden layers in one launch (DEQSCI_TC_CHAIN): parity tests first, then batch 1 and batch 32 per mode
mkdir -p gpurun_out
:
for B in ${1:-1 32}; do
for rep in 1 2; do
for M in 0 1; do
  DEQSCI_TC_CHAIN_EXPERIMENTAL=$M timeout 600 python bench.py --steps 2 --warmup 2 --batch $B --no-cpu-baseline --no-extras > gpurun_out/chain_ab_$M.log 2>&1
  python - "$M" "$B" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/chain_ab_%s.log' % sys.argv[1]) if x.startswith('{')]
if not l:
    print(open('gpurun_out/chain_ab_%s.log' % sys.argv[1]).read()[-2500:]); raise SystemExit
d=json.loads(l[-1])
print("B %s CHAIN %s: value %.3f ms/step %.2f hidden %.4f ms x %d launches, frac %.4f clocks %s psnr %.4f" % (sys.argv[2], sys.argv[1], d['value'], d['ms_per_step'], d['kernels']['conv_hidden']['avg_ms'], d['kernels']['conv_hidden']['launches'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['check']['psnr_vs_synthetic_gt_db']))
PY
done
done
done
