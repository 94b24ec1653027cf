#!/usr/bin/env python
"""Config 5: DE-GAP-FFDnet implicit-differentiation training step on synthetic 256x256x8 batches,
one process per GPU, gradients averaged with ONE flat NCCL all-reduce (deqsci_b200.distributed).

    python scripts/bench_train.py --steps 3 --warmup 1 --batch 2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py ...

Reports step time (max over ranks, CUDA events) and the all-reduce time inside it.  The train-mode forward
solve (batch-statistics BatchNorm, deqsci_iterate_train) and the backward Anderson solve (GAP-projector VJP
kernel + Anderson kernels) run on the library's kernels; the single graph-attached f call and its backward
run on PyTorch/cuDNN pinned to fp32.  Not the driver's bench contract (that is bench.py); a side measurement."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=2, help="measurements per GPU")
    ap.add_argument("--max-iter", type=int, default=100, help="and_maxiters (entry default)")
    ap.add_argument("--denoiser", default="ffdnet", choices=["ffdnet", "SimpleCNN"])
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from deqsci_b200 import _lib
    _lib.lib()
    r = bench.train_step_bench(dev, rank, world, steps=args.steps, warmup=args.warmup, batch=args.batch,
                               max_iter=args.max_iter, denoiser=args.denoiser)
    if rank == 0:
        r["native"] = ("forward solve (train-mode BatchNorm kernels), backward Anderson solve + GAP VJP, gradient exchange + "
                       "Adam (one kernel); the one graph-attached f call on cuDNN fp32")
        print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
