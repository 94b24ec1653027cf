#!/bin/bash
# round-2 multi-GPU check: fused exchange+Adam kernel (P2P and NCCL variants), the 2-GPU tests, bench at N GPUs
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r2}
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/fused_adam_check.py > gpurun_out/fused_adam_n${N}_$TAG.log 2>&1; echo "fused adam exit $?"; grep -v "^W\|warn" gpurun_out/fused_adam_n${N}_$TAG.log | tail -n 8
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu -x -k "two_gpus" > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n${N}_$TAG.log 2>&1; echo "bench exit $?"; grep '^{' gpurun_out/bench_n${N}_$TAG.log | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','n_gpus','ms_per_step','ranks')}); print(d.get('train_step'))"
grep -v '^{' gpurun_out/bench_n${N}_$TAG.log | tail -n 5
