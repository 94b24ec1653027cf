#!/bin/bash
# same-box sweep: chained kernel on/off per batch size, look-ahead on/off at batch 32 (alternating, two repetitions)
mkdir -p gpurun_out
run() {  # run <label> <batch> env...
  local label=$1 B=$2; shift 2
  env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --batch $B --no-cpu-baseline --no-extras > gpurun_out/sweep.log 2>&1
  python - "$label" "$B" <<'PY'
import json, sys
l=[x for x in open('gpurun_out/sweep.log') if x.startswith('{')]
if not l:
    print(sys.argv[1], "FAILED", open('gpurun_out/sweep.log').read()[-1500:]); raise SystemExit
d=json.loads(l[-1])
print("B %2s %-22s value %.3f ms/recon %.2f hidden %.4f ms x %d, clocks %s" % (sys.argv[2], sys.argv[1], d['value'], d['ms_per_step']/int(sys.argv[2]), d['kernels']['conv_hidden']['avg_ms'], d['kernels']['conv_hidden']['launches'], d['clocks']['sm_mhz']))
PY
}
for rep in 1 2; do
  for B in ${1:-1 2 4 8 16}; do
    run "chain0" $B DEQSCI_TC_CHAIN_EXPERIMENTAL=0
    run "chain1" $B DEQSCI_TC_CHAIN_EXPERIMENTAL=1 DEQSCI_TC_CHAIN_MAX_ROUNDS=100
  done
  run "chain0 look1" 32 DEQSCI_TC_CHAIN_EXPERIMENTAL=0
  run "chain0 look0" 32 DEQSCI_TC_CHAIN_EXPERIMENTAL=0 DEQSCI_TC_LOOKAHEAD=0
  run "chain1 look1" 32 DEQSCI_TC_CHAIN_EXPERIMENTAL=1 DEQSCI_TC_CHAIN_MAX_ROUNDS=100
  run "chain1 look0" 32 DEQSCI_TC_CHAIN_EXPERIMENTAL=1 DEQSCI_TC_CHAIN_MAX_ROUNDS=100 DEQSCI_TC_LOOKAHEAD=0
done
