import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
dev = torch.device("cuda", 0)
from deqsci_b200 import _lib
_lib.lib()
solver, deq = bench.build_deq(dev, "tc_split", "ffdnet", 100)
solver.train(); solver.nonlinear_op.train()
op = solver.nonlinear_op
for i in range(4):
    for p in op.parameters():
        p.data.add_(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = op.native_plan(dev, train=True)
    torch.cuda.synchronize()
    print("train plan rebuild %.2f ms" % ((time.perf_counter() - t0) * 1e3))
import cProfile, pstats
for p in op.parameters():
    p.data.add_(0)
pr = cProfile.Profile(); pr.enable(); op.native_plan(dev, train=True); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
