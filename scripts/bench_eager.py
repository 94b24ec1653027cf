#!/usr/bin/env python
"""On-box GPU bar: the reference's algorithm as a user gets it today -- PyTorch eager
ops (cuDNN convs, separate BatchNorm/ReLU kernels, bmm + batched solve, two .item() syncs per
iteration), written after the reference's call sequence -- timed on the same B200 for DE-GAP-FFDnet,
256x256x8, 180 iterations.  Side measurement: the reference itself cannot travel to the box (it needs
/root/reference and import shims), so this restates its PyTorch call sites with
the torch layers of deqsci_b200's module mirrors as weight containers.

    python scripts/bench_eager.py --batch 1 --tf32 1      # PyTorch default flags
    python scripts/bench_eager.py --batch 16 --tf32 0     # fp32 convs (the parity oracle setting)
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run(dev, y, Phi, tf32, max_iter=180, reps=2, weights_solver=None):
    """Reconstructs the batch (y [B,H,W], Phi [B,H,W,T]) with the restated eager call sequence; returns recon/s
    (mean over `reps` after one warm-up).  TF32 flags are restored afterwards."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    try:
        return _run(dev, y, Phi, max_iter, reps, weights_solver)[0]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _run(dev, y, Phi, max_iter, reps, weights_solver=None):
    solver = weights_solver or bench.build_deq(dev, "fp32")[0]
    seq = solver.nonlinear_op.intermediate_dncnn.itermediate_dncnn       # torch layers = weight containers
    Phi_sum = torch.sum(Phi, dim=3)
    Phi_sum[Phi_sum == 0] = 1
    state = {"sigma": None, "ymean": None}

    def f(z):                               # EquilibriumProxGradSCI.forward, tag 'ffdnet', as eager torch ops
        B, H, W, T = z.shape
        fb = torch.sum(z * Phi, dim=3)
        z = z + ((y - fb) / Phi_sum)[:, :, :, None] * Phi
        if state["ymean"] is None or bool(state["ymean"] != y.mean()):    # the reference's per-call sync
            state["sigma"] = torch.full((B * T,), 60 / 255, device=dev)
            state["ymean"] = y.mean()
        else:
            state["sigma"] = state["sigma"] * 0.971
        x = z.permute(0, 3, 1, 2).contiguous().view(B * T, 1, H, W)
        inp = torch.cat((state["sigma"].view(-1, 1, 1, 1).expand(B * T, 1, H // 2, W // 2),
                         F.pixel_unshuffle(x, 2)), 1)
        noise = F.pixel_shuffle(seq(inp), 2)
        return z - noise.view(B, T, H, W).permute(0, 2, 3, 1)

    def andersonexp(x0, m=5, lam=1e-2, max_iter=180, tol=1e-5, beta=1.0):   # reference :153-189
        bsz = x0.shape[0]
        N = x0[0].numel()
        X = torch.zeros(bsz, m, N, device=dev)
        Fh = torch.zeros(bsz, m, N, device=dev)
        X[:, 0], Fh[:, 0] = x0.reshape(bsz, -1), f(x0).reshape(bsz, -1)
        X[:, 1], Fh[:, 1] = Fh[:, 0], f(Fh[:, 0].reshape(x0.shape)).reshape(bsz, -1)
        Hm = torch.zeros(bsz, m + 1, m + 1, device=dev)
        Hm[:, 0, 1:] = Hm[:, 1:, 0] = 1
        rhs = torch.zeros(bsz, m + 1, 1, device=dev)
        rhs[:, 0] = 1
        k = 1
        res = None
        for k in range(2, max_iter):
            n = min(k, m)
            G = Fh[:, :n] - X[:, :n]
            Hm[:, 1:n + 1, 1:n + 1] = torch.bmm(G, G.transpose(1, 2)) + lam * torch.eye(n, device=dev)[None]
            alpha = torch.linalg.solve(Hm[:, :n + 1, :n + 1], rhs[:, :n + 1])[:, 1:n + 1, 0]
            X[:, k % m] = beta * (alpha[:, None] @ Fh[:, :n])[:, 0] + (1 - beta) * (alpha[:, None] @ X[:, :n])[:, 0]
            Fh[:, k % m] = f(X[:, k % m].reshape(x0.shape)).reshape(bsz, -1)
            res = (Fh[:, k % m] - X[:, k % m]).norm().item() / (1e-5 + Fh[:, k % m].norm().item())
            if res < tol:
                break
        return X[:, k % m].view_as(x0), res

    times = []
    res = None
    with torch.no_grad():
        for rep in range(reps + 1):
            state["ymean"] = None
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x0 = y[:, :, :, None] * Phi
            zs, res = andersonexp(x0, max_iter=max_iter)
            z = f(zs)
            f(z)                                # the reference's second post-solver call
            e1.record()
            torch.cuda.synchronize()
            if rep > 0:
                times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    return y.shape[0] * 1e3 / ms, ms, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--tf32", type=int, default=1)
    ap.add_argument("--max-iter", type=int, default=180)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = bool(a.tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(a.tf32)
    dev = torch.device("cuda", 0)
    y, Phi, _ = (t.to(dev) for t in bench.synthetic_batch(0, a.batch))
    rps, ms, res = _run(dev, y, Phi, a.max_iter, a.reps)
    print(json.dumps({"impl": "pytorch eager on GPU (restated reference call sequence)", "tf32": bool(a.tf32),
                      "batch": a.batch, "ms_per_batch": ms, "recon_per_s": rps, "res": res}))


if __name__ == "__main__":
    main()
