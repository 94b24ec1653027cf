#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --batch 16 > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N exit $?"
tail -n 3 gpurun_out/bench_n$N.log | cut -c 1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --cpu-iters 4 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref N=$N exit $?"
tail -n 2 gpurun_out/bench_ref_n$N.log | cut -c 1-800
