#!/usr/bin/env python
"""Per-kernel times of one eval-mode and one train-mode iterate call at a small batch (torch.profiler):
what the train-mode BatchNorm epilogue (per-channel statistics) costs in the hidden conv kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
from deqsci_b200 import _lib  # noqa: E402
_lib.lib()
solver, deq = bench.build_deq(dev, "tc_split", "ffdnet", 100)
op = solver.nonlinear_op
y, phi, gt = (t.to(dev) for t in bench.synthetic_batch(0, B))
ps = phi.sum(3)
ps[ps == 0] = 1
z = torch.rand_like(phi)
op.eval()
pe = op.native_plan(dev)
op.train()
pt = op.native_plan(dev, train=True)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
for name, fn in (("eval", lambda: pe.iterate(z, y, phi, ps, 0.2)),
                 ("train", lambda: pt.iterate_train(z, y, phi, ps, 0.2, op.bn_slots()))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
    print("====", name, "B =", B)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
