#!/usr/bin/env python
"""Where the config-5 training step spends its GPU time: torch.profiler kernel table of one step
(after warm-up) of scripts/bench_train.py's loop.  Diagnostic only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

denoiser = sys.argv[1] if len(sys.argv) > 1 else "ffdnet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
from deqsci_b200 import _lib  # noqa: E402
from deqsci_b200.utils.cg_utils import Phi_sum_, initial_point  # noqa: E402
_lib.lib()
solver, deq = bench.build_deq(dev, "tc_split", denoiser, 100)
solver.train()
solver.nonlinear_op.train()
opt = torch.optim.Adam(solver.parameters(), lr=1e-4)
y, phi, gt = (t.to(dev) for t in bench.synthetic_batch(0, batch))
loss_fn = torch.nn.MSELoss(reduction="mean")


def step():
    opt.zero_grad()
    phi_sum = Phi_sum_(phi)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    rec = deq.forward(y, phi, phi_sum, initial_point=initial_point(y, phi, phi_sum, gt))
    ev[1].record()
    loss = loss_fn(rec, gt)
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]


for _ in range(2):
    step()
print("forward(solve + 2 f calls) / backward(hook solve + autograd) / adam  ms:", step())
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
