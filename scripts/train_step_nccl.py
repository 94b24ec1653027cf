"""torchrun entry: one DE-GAP-CNN implicit-diff training step per rank on different halves of the
golden batch, gradients averaged with ONE flat NCCL all-reduce; checks every rank ends with the same
gradient = mean of the per-rank gradients, then takes an Adam step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from test_gpu_parity import build_solver  # noqa: E402
from deqsci_b200.distributed import allreduce_mean_gradients, shard_range  # noqa: E402
from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq  # noqa: E402
from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_  # noqa: E402

v = dict(np.load(os.path.join(ROOT, "tests", "golden", "train_vectors.npz")))
lo, hi = shard_range(v["gt"].shape[0], rank, world)
solver = build_solver("SimpleCNN", dev)
solver.train()
deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=12, tol=1e-5)
opt = torch.optim.Adam(solver.parameters(), lr=1e-4)
gt, Phi, y = (torch.from_numpy(v[k][lo:hi]).to(dev) for k in ("gt", "Phi", "y"))
rec = deq.forward(y, Phi, Phi_sum_(Phi), initial_point=At_torch_(y, Phi))
loss = torch.nn.MSELoss(reduction="mean")(rec, gt)
loss.backward()
local_flat = torch.cat([p.grad.reshape(-1) for p in solver.parameters()]).clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = allreduce_mean_gradients(solver.parameters())
e1.record()
torch.cuda.synchronize()
avg_flat = torch.cat([p.grad.reshape(-1) for p in solver.parameters()])
gathered = [torch.empty_like(local_flat) for _ in range(world)]
dist.all_gather(gathered, local_flat)
want = torch.stack(gathered).mean(0)
err = float((avg_flat - want).norm() / want.norm())
opt.step()
if rank == 0:
    print("floats all-reduced: %d, all-reduce %.3f ms, rel err vs mean of rank grads %.2e" % (n, e0.elapsed_time(e1), err))
    print("ALLREDUCE_OK" if err < 1e-6 and n == 74880 else "ALLREDUCE_BAD")
dist.destroy_process_group()
