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
    from deqsci_b200.distributed import allreduce_mean_gradients, max_over_ranks, shard_range
    from deqsci_b200.utils.cg_utils import Phi_sum_, initial_point
    _lib.lib()
    solver, deq = bench.build_deq(dev, "tc_split", args.denoiser, args.max_iter)
    solver.train()
    solver.nonlinear_op.train()
    opt = torch.optim.Adam(solver.parameters(), lr=1e-4)
    lo, hi = shard_range(world * args.batch, rank, world)
    y, phi, gt = (t.to(dev) for t in bench.synthetic_batch(lo, hi - lo))
    loss_fn = torch.nn.MSELoss(reduction="mean")
    ar_ms, step_ms, losses = [], [], []
    for it in range(args.warmup + args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        opt.zero_grad()
        phi_sum = Phi_sum_(phi)
        rec = deq.forward(y, phi, phi_sum, initial_point=initial_point(y, phi, phi_sum, gt))
        loss = loss_fn(rec, gt)
        loss.backward()
        e[1].record()
        n = allreduce_mean_gradients(solver.parameters())
        e[2].record()
        opt.step()
        e[3].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            step_ms.append(max_over_ranks(e[0].elapsed_time(e[3]), dev))
            ar_ms.append(max_over_ranks(e[1].elapsed_time(e[2]), dev))
            losses.append(float(loss.detach()))
    if rank == 0:
        print(json.dumps({"metric": "DE-GAP-%s implicit-diff training step" % args.denoiser, "n_gpus": world,
                          "batch_per_gpu": args.batch, "and_maxiters": args.max_iter,
                          "ms_per_step": float(np.mean(step_ms)), "allreduce_ms": float(np.mean(ar_ms)),
                          "allreduce_floats": int(n), "steps_per_s": 1e3 / float(np.mean(step_ms)),
                          "measurements_per_s": world * args.batch * 1e3 / float(np.mean(step_ms)),
                          "forward_res": deq.forward_res, "backward_res": deq.backward_res, "loss": losses,
                          "native": "forward solve (train-mode BatchNorm kernels), backward Anderson solve + GAP VJP; the one graph-attached f call on cuDNN fp32"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
