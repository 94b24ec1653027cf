"""Where the training step's time goes with the native / autograd backward: CUDA-event and wall-clock time of the
forward (solve + graph-attached call + second call) and of loss.backward() (hook solve + weight gradients)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

d = sys.argv[1] if len(sys.argv) > 1 else "ffdnet"
dev = torch.device("cuda", 0)
from deqsci_b200.utils.cg_utils import Phi_sum_, initial_point  # noqa: E402
solver, deq = bench.build_deq(dev, "tc_split", d, 100)
solver.train()
solver.nonlinear_op.train()
use_sync = os.environ.get("DIAG_SYNC", "0") == "1"
if use_sync:
    from deqsci_b200.distributed import GradientSynchronizer
    opt = GradientSynchronizer(solver.parameters(), lr=1e-4)
else:
    opt = torch.optim.Adam(solver.parameters(), lr=1e-4)
batches = [tuple(t.to(dev) for t in bench.synthetic_batch(2 * s_, 2)) for s_ in range(2)]
for it in range(8):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t0 = time.perf_counter()
    e[0].record()
    opt.zero_grad()
    y, phi, gt = batches[it % 2]
    ps = Phi_sum_(phi)
    rec = deq.forward(y, phi, ps, initial_point=initial_point(y, phi, ps, gt))
    loss = torch.nn.functional.mse_loss(rec, gt)
    e[1].record()
    t1 = time.perf_counter()
    loss.backward()
    e[2].record()
    t2 = time.perf_counter()
    opt.step()
    e[3].record()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it >= 5:
        print("%s sync=%d native=%s  fwd %.1f ms (cpu queued in %.1f)  backward %.1f ms (cpu %.1f)  opt %.1f  wall %.1f" % (
            d, use_sync, os.environ.get("DEQSCI_NATIVE_BACKWARD", "1"), e[0].elapsed_time(e[1]), (t1 - t0) * 1e3,
            e[1].elapsed_time(e[2]), (t2 - t1) * 1e3, e[2].elapsed_time(e[3]), (t3 - t0) * 1e3))
