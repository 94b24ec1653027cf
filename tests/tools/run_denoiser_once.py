"""Helper of test_pair_kernel_issue_modes_agree: iterate-map calls (repeated, on a fresh and on a reused plan) + a short
DE-GAP solve in a fresh process (the issue mode of the pair kernel, DEQSCI_TC_RS, is fixed per process), saved to an .npz.
usage: python tests/tools/run_denoiser_once.py <denoiser> <out.npz> <B> <H> <W>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    d, out, B, H, W = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    from test_gpu_parity import build_solver
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    from oracle import deqsci_oracle as orc
    dev = torch.device("cuda", 0)
    data = orc.synthetic_measurements(3, B, H=H, W=W, T=8)
    y, Phi = torch.from_numpy(data["y"]).to(dev), torch.from_numpy(data["Phi"]).to(dev)
    Ps, x0 = Phi_sum_(Phi), At_torch_(y, Phi)
    with torch.no_grad():
        # the same iterate-map call 12 times over (fresh sigma schedule each time): a race in the kernels shows up as
        # run-to-run differences
        reps = [build_solver(d, dev)(x0, y, Phi, Ps).clone() for _ in range(12)]
    rep_maxdiff = max(float((r - reps[0]).abs().max()) for r in reps)
    x1 = (x0 * 0.5 + 0.1).contiguous()
    with torch.no_grad():
        sA = build_solver(d, dev)
        a1 = sA(x0, y, Phi, Ps).clone()
        a2 = sA(x1, y, Phi, Ps).clone()          # second launch on the same plan, other data
        b1 = build_solver(d, dev)(x1, y, Phi, Ps).clone()    # first launch of a fresh plan, other data in the workspace
    solver = build_solver(d, dev)
    with torch.no_grad():
        f1 = solver(x0, y, Phi, Ps).clone()                       # one iterate-map call
        deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=10, tol=1e-6)
        z = deq.forward(y, Phi, Ps, initial_point=x0, train_flag=False)
    torch.cuda.synchronize()
    np.savez(out, f1=f1.cpu().numpy(), z=z.cpu().numpy(), res=np.float64(deq.forward_res), rep_maxdiff=np.float64(rep_maxdiff),
             a1=a1.cpu().numpy(), a2=a2.cpu().numpy(), b1=b1.cpu().numpy())


if __name__ == "__main__":
    main()
