#!/usr/bin/env python
"""Times the implicit-differentiation backward hook (Anderson solve on the GAP-projector VJP) separately from
the autograd backward of the graph-attached f call, for the config-5 training step.  Diagnostic only."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
from deqsci_b200 import _lib  # noqa: E402
from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq  # noqa: E402
from deqsci_b200.utils.cg_utils import Phi_sum_, initial_point  # noqa: E402
_lib.lib()
solver, deq = bench.build_deq(dev, "tc_split", "ffdnet", 100)
solver.train()
solver.nonlinear_op.train()
y, phi, gt = (t.to(dev) for t in bench.synthetic_batch(0, 2))
phi_sum = Phi_sum_(phi)
orig = eq.andersonexp
times = []


def timed(f, x0, **kw):
    if not torch.is_grad_enabled() and getattr(f, "supports_out", False) and not hasattr(f, "f"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = orig(f, x0, **kw)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
        return out
    return orig(f, x0, **kw)


deq.solver = timed
for it in range(4):
    solver.zero_grad()
    rec = deq.forward(y, phi, phi_sum, initial_point=initial_point(y, phi, phi_sum, gt))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    torch.nn.functional.mse_loss(rec, gt).backward()
    torch.cuda.synchronize()
    print("backward total %.2f ms, hook solve %s ms" % ((time.perf_counter() - t0) * 1e3, ["%.2f" % t for t in times]))
    times.clear()
