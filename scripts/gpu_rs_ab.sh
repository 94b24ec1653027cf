#!/bin/bash
# experiment: row-stationary issue order of the pair kernel (DEQSCI_TC_RS: 0 output-stationary, 1 row-stationary,
# 2 row-stationary + A-collector hints): parity of the hidden layer first, then the benchmark at batch $1 per mode
mkdir -p gpurun_out
B=${1:-32}
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "hidden_layer or denoiser_vs_oracle or driver_equals" > gpurun_out/rs_quick.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/rs_quick.log
for rep in 1 2; do
for M in 0 1 2; do
  DEQSCI_TC_RS=$M timeout 600 python bench.py --steps 2 --warmup 2 --batch $B --no-cpu-baseline --no-extras > gpurun_out/rs_ab_$M.log 2>&1
  python - "$M" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/rs_ab_%s.log' % sys.argv[1]) if x.startswith('{')]
if not l:
    print(open('gpurun_out/rs_ab_%s.log' % sys.argv[1]).read()[-2000:]); raise SystemExit
d=json.loads(l[-1])
print("RS %s: value %.3f ms/step %.1f hidden %.4f ms clocks %s check %s" % (sys.argv[1], d['value'], d['ms_per_step'], d['kernels']['conv_hidden']['avg_ms'], d['clocks']['sm_mhz'], d.get('check')))
PY
done
done
