#!/usr/bin/env python
"""bench.py — DE-GAP-FFDnet reconstructions/s (256x256x8, 180 Anderson iterations) on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3                    # this repo's CUDA path
    python bench.py --impl reference --steps 1 --warmup 0            # CPU arm (numpy port of the reference)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: `--batch` synthetic measurements per GPU
(random binary masks, independent per measurement) reconstructed with
DEQFixedPoint(EquilibriumProxGradSCI(FFDNet), andersonexp, m=5, beta=1, lam=1e-2, max_iter=180,
tol=1e-5) = 182 iterate-map evaluations + 178 Anderson updates each.
Independent measurements are sharded over ranks with no data-path collective (weak scaling).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: the same call with host
buffers (pinned H2D of y and Phi, D2H of the reconstruction inside the timed region); `roofline`:
the hidden-layer tcgen05 conv kernel (tensor bound), its duration sampled with CUDA events on the
launch stream inside the timed steps; `cpu_baseline`: the numpy port of the reference timed on the
host cores on a bounded sample.  Weights: the reference's FFDNet gray weights (net_gray.pth,
stand-in for the missing ffdnet.ckpt) from tests/golden/.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 256
T = 8
MAX_ITER = 180
M_HIST = 5
F_CALLS = MAX_ITER + 1            # 180 solver calls + the reconstruction call (the reference's 182nd is wasted)
F_CALLS_REFERENCE = MAX_ITER + 2
AND_UPDATES = MAX_ITER - 2
SEED = 20260117
# algorithmic work: hidden 64->64 3x3 layers at 128x128 on 8 frames per measurement
HIDDEN_FLOP_PER_LAUNCH_PER_MEAS = 2 * 9 * 64 * 64 * (H // 2) * (W // 2) * T
STACK_FLOP_PER_F_PER_MEAS = 2 * 9 * (5 * 64 + 13 * 64 * 64 + 64 * 4) * (H // 2) * (W // 2) * T


def _smooth(g, size, cells):
    """Low-pass random image in [0,1]: U[0,1) noise on a cells x cells grid, bicubic-upsampled."""
    import torch.nn.functional as F
    c = torch.rand(1, 1, cells, cells, generator=g)
    return F.interpolate(c, size=(size, size), mode="bicubic", align_corners=False)[0, 0]


def synthetic_cube(g, kind):
    """One ground-truth cube [H,W,T] in [0,1] from generator g.
    uniform: x ~ U[0,1) per voxel (white noise; SURVEY 8(d) default).
    lowpass: every frame an independent low-pass image.
    video:   one low-pass scene (three octaves) translated by a whole-pixel velocity per frame plus 1 % sensor
             noise -- temporally coherent like the benchmark videos ('optionally low-pass
             filtered to be video-like; state which')."""
    if kind == "uniform":
        return torch.rand(H, W, T, generator=g)
    if kind == "lowpass":
        fr = [0.7 * _smooth(g, H, 10) + 0.3 * _smooth(g, H, 40) for _ in range(T)]
        x = torch.stack(fr, 2)
    elif kind == "video":
        P = 2 * 2 * (T - 1)                               # margin for |v| <= 2 pixels per frame
        S = H + P
        base = 0.6 * _smooth(g, S, 10) + 0.3 * _smooth(g, S, 36) + 0.1 * _smooth(g, S, 120)
        v = torch.randint(-2, 3, (2,), generator=g)
        o = P // 2
        fr = [base[o + int(v[0]) * t:o + int(v[0]) * t + H, o + int(v[1]) * t:o + int(v[1]) * t + W] for t in range(T)]
        x = torch.stack(fr, 2) + 0.01 * torch.randn(H, W, T, generator=g)
    else:
        raise ValueError(kind)
    x = x - x.min()
    return (x / x.max().clamp_min(1e-6)).contiguous()


DATA_KIND = os.environ.get("DEQSCI_BENCH_DATA", "video")


def synthetic_batch(start, count, kind=None):
    """Measurement i is drawn from torch.Generator().manual_seed(SEED + i): identical under any
    sharding.  x per `kind` (synthetic_cube), Phi ~ Bernoulli(0.5) independent per measurement,
    y = sum_t Phi*x."""
    kind = kind or DATA_KIND
    ys, ps, xs = [], [], []
    for i in range(start, start + count):
        g = torch.Generator().manual_seed(SEED + i)
        x = synthetic_cube(g, kind)
        phi = (torch.rand(H, W, T, generator=g) < 0.5).float()
        xs.append(x)
        ps.append(phi)
        ys.append((x * phi).sum(2))
    return torch.stack(ys), torch.stack(ps), torch.stack(xs)


def load_ffdnet_weights():
    d = np.load(os.path.join(ROOT, "tests", "golden", "weights_ffdnet_gray.npz"))
    return {k: d[k] for k in d.files if not k.startswith("shape::")}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        # clocks under load: the upper half of the samples (idle samples at the edges excluded)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# CPU arm: numpy port of the reference (oracle/), bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_port_recon_per_s(n_iters, y, phi):
    """Times `n_iters` iterations (iterate map + Anderson update) of one 256x256x8 measurement with
    the numpy port on the host cores and extrapolates to a full reconstruction."""
    from oracle import deqsci_oracle as orc            # bench.py's cpu legs are allowed to use the oracle
    orc.set_conv_backend("torch")                      # the conv kernel the reference itself runs on CPU
    sd = load_ffdnet_weights()
    f = orc.ProxGradSCI("ffdnet", sd)
    yn, pn = y[:1].numpy(), phi[:1].numpy()
    ps = orc.phi_sum(pn)
    t0 = time.perf_counter()
    z, _ = orc.andersonexp(lambda q: f(q, yn, pn, ps), orc.At(yn, pn), m=M_HIST, lam=1e-2, max_iter=n_iters,
                           tol=1e-5, beta=1.0)
    dt = time.perf_counter() - t0
    per_call = dt / n_iters                            # n_iters f calls, n_iters-2 Anderson updates
    full = per_call * F_CALLS_REFERENCE                # the reference evaluates f 182 times per reconstruction
    return 1.0 / full, dt, per_call


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    y, phi, _ = synthetic_batch(0, 1)
    vals = []
    n_iters = args.cpu_iters
    for s in range(args.warmup + args.steps):
        v, dt, per_call = cpu_port_recon_per_s(n_iters, y, phi)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    cores = os.cpu_count()
    sample = ("%d of %d iterate-map evaluations (+ Anderson updates) of one 256x256x8 measurement, port of the reference "
              "(numpy + torch CPU conv2d), extrapolated x%d/%d" % (n_iters, F_CALLS_REFERENCE, F_CALLS_REFERENCE, n_iters))
    line = {"impl": "reference", "metric": "DE-GAP-FFDnet reconstructions/s (256x256x8, 180 Anderson iterations)",
            "value": value, "unit": "recon/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (U[0,1) frames, Bernoulli(0.5) masks); FFDNet weights net_gray.pth stand-in",
            "config": {"workload": "DE-GAP-FFDnet 256x256x8, 180 iterations, batch 1, CPU", "timing": "wall clock"},
            "cpu_baseline": {"value": value, "unit": "recon/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "recon/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def build_deq(dev, precision, denoiser="ffdnet", max_iter=MAX_ITER):
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq_utils
    from deqsci_b200.solvers.equilibrium_solvers_yaping import EquilibriumProxGradSCI
    from deqsci_b200.utils.cg_utils import A_torch_, At_torch_
    from deqsci_b200.video_sci_proxgrad import build_denoiser
    net = build_denoiser(denoiser)
    net.precision = precision
    net.eval()
    solver = EquilibriumProxGradSCI(A=A_torch_, At=At_torch_, nonlinear_operator=net, eta=0.2)
    wfile = {"ffdnet": "weights_ffdnet_gray.npz", "SimpleCNN": "weights_cnn.npz",
             "RealSN_SimpleCNN": "weights_rsn_cnn.npz"}[denoiser]
    d = np.load(os.path.join(ROOT, "tests", "golden", wfile))
    sd = {k: torch.from_numpy(d[k]) for k in d.files if not k.startswith("shape::")}
    solver.load_state_dict(sd, strict=False)
    solver = solver.to(dev)
    deq = eq_utils.DEQFixedPoint(solver, eq_utils.andersonexp, m=M_HIST, beta=1.0, lam=1e-2, max_iter=max_iter,
                                 tol=1e-5)
    return solver, deq


def reconstruct(deq, y, phi):
    """The public-API call a user makes (reference training/sci_equilibrium_training.py:159-178)."""
    from deqsci_b200.utils.cg_utils import Phi_sum_, initial_point
    phi_sum = Phi_sum_(phi)
    x0 = initial_point(y, phi, phi_sum, None)
    return deq.forward(y, phi, phi_sum, initial_point=x0, train_flag=False)


def run_gpu_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    from deqsci_b200 import _lib
    lib = _lib.lib()                                  # raises if libdeqsci.so is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    solver, deq = build_deq(dev, args.precision, args.denoiser, args.max_iter)
    from deqsci_b200.distributed import shard_range
    lo, hi = shard_range(world * B, rank, world)      # contiguous index range per rank, no collective
    y_h, phi_h, gt = synthetic_batch(lo, hi - lo)
    y_h, phi_h = y_h.pin_memory(), phi_h.pin_memory()
    out_h = torch.empty(B, H, W, T).pin_memory()
    y_d, phi_d = y_h.to(dev), phi_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time from CUDA events; max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    res_hold = {}

    def step_resident():
        res_hold["z"] = reconstruct(deq, y_d, phi_d)

    def step_e2e():
        yd = y_h.to(dev, non_blocking=True)
        pd = phi_h.to(dev, non_blocking=True)
        z = reconstruct(deq, yd, pd)
        out_h.copy_(z, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.deqsci_profile_begin(args.sample_every)
    ms_total = timed(step_resident, args.steps)
    k = 7
    ms_sum, n_samp, n_launch = (ctypes.c_double * k)(), (ctypes.c_longlong * k)(), (ctypes.c_longlong * k)()
    lib.deqsci_profile_end(ms_sum, n_samp, n_launch)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # sanity of the timed work: finite output, data-consistent (A z ~ y on seen pixels)
    z = res_hold["z"]
    finite = bool(torch.isfinite(z).all())
    psnr = float(10 * torch.log10(1.0 / ((z.clip(0, 1).cpu() - gt) ** 2).mean()))

    step_e2e()                                        # warm the pinned-copy path
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # latency mode (how the reference itself runs its benchmark: one measurement at a time), reported beside
    # the throughput numbers; single-GPU runs only
    lat_ms = None
    if world == 1 and B > 1:
        y1, p1 = y_d[:1].contiguous(), phi_d[:1].contiguous()
        for _ in range(2):
            reconstruct(deq, y1, p1)
        lat_ms = timed(lambda: reconstruct(deq, y1, p1), 3) / 3

    if rank != 0:
        return
    peaks, peak_src = measured_peaks()
    hid_ms = ms_sum[2] / max(n_samp[2], 1)
    res_div = 2 if args.denoiser == "ffdnet" else 1              # FFDNet convs run at half resolution
    flop_per_launch = 2 * 9 * 64 * 64 * (H // res_div) * (W // res_div) * T * B
    achieved = flop_per_launch / (hid_ms * 1e-3) / 1e12 if hid_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    kinds = ["gap", "conv_first", "conv_hidden", "conv_last", "anderson_gram", "anderson_solve", "anderson_mix"]
    shares = {kinds[i]: {"launches": int(n_launch[i]), "sampled": int(n_samp[i]),
                         "avg_ms": (ms_sum[i] / n_samp[i]) if n_samp[i] else None,
                         "est_ms_per_step": (ms_sum[i] / n_samp[i] * n_launch[i] / args.steps) if n_samp[i] else None}
              for i in range(k)}
    line = {
        "metric": "DE-GAP-FFDnet reconstructions/s (256x256x8, 180 Anderson iterations)" if args.denoiser == "ffdnet"
                  else "DE-GAP-%s reconstructions/s (256x256x8, %d Anderson iterations)" % (args.denoiser, args.max_iter),
        "value": value, "unit": "recon/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic (U[0,1) frames, Bernoulli(0.5) masks, independent per measurement); FFDNet weights = "
                "reference net_gray.pth stand-in for the missing ffdnet.ckpt",
        "config": {"workload": "batch-sharded DE-GAP-FFDnet, 256x256x8, and_maxiters=180, m=5, beta=1, lam=1e-2 "
                               "(BASELINE.json configs[3])",
                   "batch_per_gpu": B, "measurements_per_step": world * B, "precision": args.precision,
                   "arithmetic": "fp32 state and accumulation; conv operands split into fp16 hi + fp16 lo*2^11 "
                                 "(3 tensor-core products, ~22 mantissa bits)" if args.precision == "tc_split"
                                 else args.precision,
                   "f_calls_per_recon": args.max_iter + 1, "anderson_updates_per_recon": args.max_iter - 2,
                   "l2_policy": "working set per step (%.1f GB/GPU) exceeds the 126 MB L2" % (
                       B * (3 * M_HIST * H * W * T * 4 + 2 * 2 * (H // 2) * (W // 2) * T * 64 * 2) / 1e9),
                   "parallelism": "measurements sharded over ranks, no collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "recon/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(y_h.numel() * 4 + phi_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel() * 4)},
        "gpu_launches": int(sum(n_launch)),
        "roofline": {"kernel": "conv_hidden_2cta_kernel (hidden 64->64 3x3 layer, tcgen05 cta_group::2)"
                               if args.precision == "tc_split" else "conv_mid_tc_kernel (hidden layer)",
                     "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     # dram read+write per launch from the ncu --set full capture in profiles/ (B=8: 497.8 MB)
                     "traffic": 497.8e6 / 8 * B if (args.precision == "tc_split" and args.denoiser == "ffdnet") else None,
                     "traffic_unit": "bytes per launch (profiles/r01_kernel_metrics.md, scaled by batch)",
                     "peak_source": peak_src + ", bf16 dense sustained",
                     "avg_launch_ms": hid_ms, "algorithmic_flop_per_launch": flop_per_launch,
                     "issued_mma_flop_factor": 3 if args.precision == "tc_split" else 1,
                     "frac_issued": (3 if args.precision == "tc_split" else 1) * achieved / peak if peak else None,
                     "sampled_launches": int(n_samp[2])},
        "kernels": shares,
        "check": {"finite": finite, "psnr_vs_synthetic_gt_db": psnr},
    }
    if lat_ms is not None:
        line["latency_batch1"] = {"ms_per_recon": lat_ms, "recon_per_s": 1e3 / lat_ms,
                                  "note": "same call at batch 1 (the reference's own test mode), inputs resident"}
    if world == 1 and not args.no_cpu_baseline and args.denoiser == "ffdnet":
        v, dt, per_call = cpu_port_recon_per_s(args.cpu_iters, y_h, phi_h)
        line["cpu_baseline"] = {"value": v, "unit": "recon/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "%d of %d iterate-map evaluations (+ Anderson updates) of one measurement, "
                                          "port of the reference (numpy + the reference's own torch CPU conv2d) on the host cores, %.1f s, extrapolated" % (
                                              args.cpu_iters, F_CALLS_REFERENCE, dt)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("DEQSCI_BENCH_BATCH", 32)),
                    help="measurements per GPU per step")
    ap.add_argument("--precision", default="tc_split", choices=["tc_split", "fp32", "tc_single"])
    ap.add_argument("--sample-every", type=int, default=11, help="event-time every k-th kernel launch")
    ap.add_argument("--cpu-iters", type=int, default=12, help="iterations of the CPU port sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--denoiser", default="ffdnet", choices=["ffdnet", "SimpleCNN", "RealSN_SimpleCNN"],
                    help="side benchmarks (config 3); the headline metric is ffdnet")
    ap.add_argument("--max-iter", type=int, default=None, help="and_maxiters (default 180 ffdnet, 100 otherwise)")
    args = ap.parse_args()

    if args.max_iter is None:
        args.max_iter = MAX_ITER if args.denoiser == "ffdnet" else 100
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_gpu_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
