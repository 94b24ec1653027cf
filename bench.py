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
N_HIDDEN = {"ffdnet": 13, "SimpleCNN": 3, "RealSN_SimpleCNN": 3}      # 64 -> 64 layers of each denoiser
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


DATA_KIND = os.environ.get("DEQSCI_BENCH_DATA", "lowpass")


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
# CPU arm: the reference's own modules (oracle/_ref, staged by oracle/make_ref.py) or, when they are not
# importable, the numpy port (oracle/deqsci_oracle.py); bounded sample
# --------------------------------------------------------------------------------------------------
DATA_TEXT = {"uniform": "U[0,1) white-noise frames", "lowpass": "low-pass random frames (bicubic-upsampled U[0,1) grids, "
             "two octaves; SURVEY 8(d) 'low-pass filtered to be video-like')", "video": "translating low-pass scene + 1 % noise"}


def data_text(kind=None):
    return ("synthetic (%s, Bernoulli(0.5) masks independent per measurement); FFDNet weights = reference net_gray.pth "
            "stand-in for the missing ffdnet.ckpt" % DATA_TEXT[kind or DATA_KIND])


def cpu_threads():
    """Use every host core: under torch.distributed.run the launcher exports OMP_NUM_THREADS=1, which made the
    round-1 CPU arm 2.6x slower at N >= 2."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_sample(n_iters, y, phi):
    """`n_iters` iterate-map evaluations (+ n_iters-2 Anderson updates) of ONE 256x256x8 measurement on the host
    cores through andersonexp, extrapolated to the 182 evaluations of a full reconstruction.  Returns a dict:
    value (recon/s), seconds, kind ('reference' = the reference's own modules, 'port' = numpy restatement),
    z (the iterate andersonexp returns, for the parity check), threads."""
    threads = cpu_threads()
    from oracle import ref_import                      # bench.py's cpu legs are allowed to use oracle/
    y1, p1 = y[:1].contiguous(), phi[:1].contiguous()
    if ref_import.use_staged_reference():
        sd = {k: torch.from_numpy(v) for k, v in load_ffdnet_weights().items()}
        # the reference hard-codes .cuda() (solvers/equilibrium_solvers_yaping.py:394,410): keep its CPU run on the CPU
        with torch.no_grad(), ref_import.cpu_only():
            solver, _ = ref_import.build_reference_deq("ffdnet", max_iter=n_iters, state_dict=sd)
            from solvers import new_equilibrium_utils_yaping as ref_eq      # the reference's module
            from utils.cg_utils import At_torch_ as ref_At
            ps = torch.sum(p1, dim=3)
            ps[ps == 0] = 1
            t0 = time.perf_counter()
            z, _ = ref_eq.andersonexp(lambda q: solver(q, y1, p1, ps), ref_At(y1, p1), m=M_HIST, lam=1e-2,
                                      max_iter=n_iters, tol=1e-5, beta=1.0)
            dt = time.perf_counter() - t0
        kind, z = "reference", z.numpy()
    else:
        from oracle import deqsci_oracle as orc
        orc.set_conv_backend("torch")                  # the conv kernel the reference itself runs on CPU
        f = orc.ProxGradSCI("ffdnet", load_ffdnet_weights())
        yn, pn = y1.numpy(), p1.numpy()
        ps = orc.phi_sum(pn)
        t0 = time.perf_counter()
        z, _ = orc.andersonexp(lambda q: f(q, yn, pn, ps), orc.At(yn, pn), m=M_HIST, lam=1e-2, max_iter=n_iters,
                               tol=1e-5, beta=1.0)
        dt = time.perf_counter() - t0
        kind = "port"
    full = dt / n_iters * F_CALLS_REFERENCE            # the reference evaluates f 182 times per reconstruction
    what = ("the reference's own modules (oracle/_ref: solvers/, networks/ffdnet/, utils/cg_utils.py, unmodified, "
            "PyTorch CPU)" if kind == "reference" else "numpy port of the reference (oracle/deqsci_oracle.py, torch CPU conv2d)")
    return {"value": 1.0 / full, "seconds": dt, "kind": kind, "z": z, "threads": threads,
            "sample": "%d of %d iterate-map evaluations (+ Anderson updates) of one 256x256x8 measurement through "
                      "andersonexp, %s, %d threads, %.1f s, extrapolated x%d/%d" % (
                          n_iters, F_CALLS_REFERENCE, what, threads, dt, F_CALLS_REFERENCE, n_iters)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    y, phi, _ = synthetic_batch(0, 1, args.data)
    vals = []
    for s_ in range(args.warmup + args.steps):
        r = cpu_sample(args.cpu_iters, y, phi)
        if s_ >= args.warmup:
            vals.append(r)
    value = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals])) * 1e3
    r = vals[-1]
    line = {"impl": "reference", "metric": "DE-GAP-FFDnet reconstructions/s (256x256x8, 180 Anderson iterations)",
            "value": value, "unit": "recon/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": data_text(args.data),
            "config": {"workload": "DE-GAP-FFDnet 256x256x8, and_maxiters=180, m=5, beta=1, lam=1e-2, batch 1, host CPU "
                                   "(BASELINE.json configs[3] workload, configs[0] execution)",
                       "step": "one step = one bounded sample: %d of the %d iterate-map evaluations of a reconstruction; "
                               "value = 1 / (sample seconds x %d / %d)" % (args.cpu_iters, F_CALLS_REFERENCE,
                                                                             F_CALLS_REFERENCE, args.cpu_iters),
                       "timing": "wall clock", "threads": r["threads"], "host_cores": os.cpu_count()},
            "cpu_baseline": {"value": value, "unit": "recon/s", "cores": r["threads"], "kind": r["kind"],
                             "sample": r["sample"]},
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
    if denoiser == "RealSN_SimpleCNN":
        # the power-iteration probes `weight_u` (training-only state, 1.2 MB of noise) are not in the fixture:
        # every other key must match
        missing, unexpected = solver.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith("weight_u") for k in missing), (missing, unexpected)
    else:
        solver.load_state_dict(sd, strict=True)
    solver = solver.to(dev)
    deq = eq_utils.DEQFixedPoint(solver, eq_utils.andersonexp, m=M_HIST, beta=1.0, lam=1e-2, max_iter=max_iter,
                                 tol=1e-5)
    return solver, deq


def reconstruct(deq, y, phi):
    """The public-API call a user makes (reference training/sci_equilibrium_training.py:159-178)."""
    from deqsci_b200.utils.cg_utils import Phi_sum_, initial_point
    with torch.no_grad():
        phi_sum = Phi_sum_(phi)
        x0 = initial_point(y, phi, phi_sum, None)
        return deq.forward(y, phi, phi_sum, initial_point=x0, train_flag=False)


def ncu_traffic(kernel, batch):
    """DRAM read+write bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/r02_kernel_metrics.json, written by scripts/summarize_profiles.py), scaled to `batch` when the
    capture was taken at another batch size."""
    p = os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")
    if not os.path.exists(p):
        return None, "no ncu capture committed"
    d = json.load(open(p)).get(kernel)
    if not d or not d.get("dram_bytes_per_launch"):
        return None, "kernel not in profiles/r02_kernel_metrics.json"
    b = d.get("batch", batch)
    return d["dram_bytes_per_launch"] * batch / b, "profiles/r02_kernel_metrics.json (ncu --set full at batch %d%s)" % (
        b, "" if b == batch else ", scaled to batch %d" % batch)


def train_step_bench(dev, rank, world, steps=5, warmup=3, batch=2, max_iter=100, denoiser="ffdnet", data=None):
    """Config 5: implicit-differentiation training steps (reference training/sci_equilibrium_training.py:54-75) on
    `batch` synthetic measurements per GPU: forward solve + graph-attached call + backward solve + gradient
    all-reduce over the ranks + Adam.  Step and all-reduce(+optimizer) times are CUDA-event times, max over ranks."""
    from deqsci_b200.distributed import GradientSynchronizer, max_over_ranks, min_over_ranks, shard_range
    from deqsci_b200.utils.cg_utils import Phi_sum_, initial_point
    solver, deq = build_deq(dev, "tc_split", denoiser, max_iter)
    solver.train()
    solver.nonlinear_op.train()
    sync = GradientSynchronizer(solver.parameters(), lr=1e-4)          # flat gradients + fused all-reduce-scale + Adam
    lo, hi = shard_range(world * batch, rank, world)
    # two batches used alternately, like a data loader: the sigma schedule restarts every step (new measurement mean)
    batches = [tuple(t.to(dev) for t in synthetic_batch(s_ * world * batch + lo, hi - lo, data)) for s_ in range(2)]
    loss_fn = torch.nn.MSELoss(reduction="mean")
    ar_ms, ar_min_ms, step_ms, losses = [], [], [], []
    import torch.distributed as dist
    for it in range(warmup + steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        sync.zero_grad()
        y, phi, gt = batches[it % 2]
        phi_sum = Phi_sum_(phi)
        rec = deq.forward(y, phi, phi_sum, initial_point=initial_point(y, phi, phi_sum, gt))
        loss = loss_fn(rec, gt)
        loss.backward()
        e[1].record()
        sync.step()                                    # all-reduce (sum) -> x 1/world + Adam in ONE kernel
        e[2].record()
        torch.cuda.synchronize(dev)
        if it >= warmup:
            step_ms.append(max_over_ranks(e[0].elapsed_time(e[2]), dev))
            ar_ms.append(max_over_ranks(e[1].elapsed_time(e[2]), dev))
            ar_min_ms.append(min_over_ranks(e[1].elapsed_time(e[2]), dev))
            losses.append(float(loss.detach()))
    return {"workload": "DE-GAP-%s implicit-diff training step, %d synthetic 256x256x8 measurements per GPU, "
                        "and_maxiters=%d, train-mode BatchNorm, MSE, Adam lr 1e-4 (BASELINE.json configs[4])" % (
                            denoiser, batch, max_iter),
            "ranks": world, "steps": steps, "warmup": warmup, "ms_per_step": float(np.mean(step_ms)),
            "allreduce_adam_ms": float(np.mean(ar_ms)), "allreduce_adam_ms_last_rank": float(np.mean(ar_min_ms)),
            "allreduce_note": "allreduce_adam_ms = max over ranks of the exchange + Adam kernel INCLUDING the wait for the slowest "
                              "rank's backward (arrival skew); _last_rank = the same interval on the rank that arrived last "
                              "(no waiting): the exchange + update itself",
            "allreduce_floats": int(sync.numel),
            "allreduce": sync.describe(), "measurements_per_s": world * batch * 1e3 / float(np.mean(step_ms)),
            "forward_res": deq.forward_res, "backward_res": deq.backward_res, "loss_first_last": [losses[0], losses[-1]],
            "backward": ("native: one autograd node -- tensor-core forward keeping the activations, csrc/backward.cu wgrad / ReLU / "
                         "BatchNorm backward, tensor-core dgrads on the adjoint plan (no cuDNN kernel in the step)"
                         if os.environ.get("DEQSCI_NATIVE_BACKWARD", "1") != "0" else "PyTorch autograd / cuDNN fp32")}


def eager_gpu_bar(dev, solver, y, phi, tf32, max_iter=MAX_ITER):
    """The on-box bar: the reference's call sequence as PyTorch eager / cuDNN ops on the same GPU (what a user of
    the reference gets today), batch 1 like the reference's own test loop.  Returns recon/s."""
    import scripts.bench_eager as be
    return be.run(dev, y[:1].contiguous(), phi[:1].contiguous(), bool(tf32), max_iter, reps=1, weights_solver=solver)


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
    # TWO batches per rank, used alternately: consecutive steps see different measurements, so every step is a
    # fresh reconstruction whose sigma schedule restarts at 60/255 -- the reference resets the schedule when the
    # measurement mean changes (solvers/equilibrium_solvers_yaping.py:409-413); re-running ONE batch would continue
    # the decayed schedule (sigma ~ 1e-3) and time reconstructions nobody wants
    sets = []
    for s_ in range(2):
        y_s, phi_s, gt_s = synthetic_batch(s_ * world * B + lo, hi - lo, args.data)
        y_s, phi_s = y_s.pin_memory(), phi_s.pin_memory()
        sets.append({"y_h": y_s, "phi_h": phi_s, "gt": gt_s, "y_d": y_s.to(dev), "phi_d": phi_s.to(dev)})
    y_h, phi_h = sets[0]["y_h"], sets[0]["phi_h"]
    y_d, phi_d = sets[0]["y_d"], sets[0]["phi_d"]
    out_h = torch.empty(B, H, W, T).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time from CUDA events; max over ranks.
        Returns (max ms, [ms of every rank])."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        per_rank = [float(ms.item())]
        if world > 1:
            allr = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(allr, ms)
            per_rank = [float(a.item()) for a in allr]
        return max(per_rank), per_rank

    res_hold = {"i": 0}

    def step_resident():
        cur = sets[res_hold["i"] % 2]
        res_hold["i"] += 1
        res_hold["z"], res_hold["gt"] = reconstruct(deq, cur["y_d"], cur["phi_d"]), cur["gt"]
        res_hold["sigma_calls"] = getattr(solver, "_n", None)

    def step_e2e():
        cur = sets[res_hold["i"] % 2]
        res_hold["i"] += 1
        yd = cur["y_h"].to(dev, non_blocking=True)
        pd = cur["phi_h"].to(dev, non_blocking=True)
        z = reconstruct(deq, yd, pd)
        out_h.copy_(z, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.deqsci_profile_begin(args.sample_every)
    ms_total, ms_ranks = timed(step_resident, args.steps)
    k = 7
    ms_sum, n_samp, n_launch = (ctypes.c_double * k)(), (ctypes.c_longlong * k)(), (ctypes.c_longlong * k)()
    lib.deqsci_profile_end(ms_sum, n_samp, n_launch)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # sanity of the timed work: finite output, PSNR against the synthetic ground truth
    z, gt = res_hold["z"], res_hold["gt"]
    finite = bool(torch.isfinite(z).all())
    psnr = float(10 * torch.log10(1.0 / ((z.clip(0, 1).cpu() - gt) ** 2).mean()))
    sigma_calls = res_hold.get("sigma_calls")          # 182 = the schedule restarted for the last timed reconstruction

    if args.profile_mode:                             # under ncu: the resident steps are all that is needed
        args.no_extras = args.no_cpu_baseline = True
        ms_e2e = float("nan")
    else:
        step_e2e()                                    # warm the pinned-copy path
        ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # latency mode (how the reference itself runs its benchmark: one measurement at a time), reported beside
    # the throughput numbers; single-GPU runs only
    lat_ms = None
    if world == 1 and B > 1 and not args.profile_mode:
        ones = [(y_d[i:i + 1].contiguous(), phi_d[i:i + 1].contiguous()) for i in range(2)]   # alternate: fresh schedule
        cnt = {"i": 0}

        def step_one():
            yy, pp = ones[cnt["i"] % 2]
            cnt["i"] += 1
            reconstruct(deq, yy, pp)
        for _ in range(2):
            step_one()
        lat_ms = timed(step_one, 4)[0] / 4

    extras = {}
    if not args.no_extras and args.denoiser == "ffdnet":
        # config 5 under the same launch: the NCCL gradient all-reduce is exercised at every N >= 2
        extras["train_step"] = train_step_bench(dev, rank, world, steps=args.train_steps, warmup=3, batch=args.train_batch,
                                                data=args.data)
    if not args.no_extras and world == 1 and args.denoiser == "ffdnet":
        side = {}
        for dn in ("SimpleCNN", "RealSN_SimpleCNN"):             # config 3: DE-GAP-CNN / DE-GAP-RSN-CNN, 100 iterations
            _, dq = build_deq(dev, args.precision, dn, 100)
            for _ in range(2):
                reconstruct(dq, y_d, phi_d)
            ms_side = timed(lambda: reconstruct(dq, y_d, phi_d), 3)[0]
            side["DE-GAP-" + {"SimpleCNN": "CNN", "RealSN_SimpleCNN": "RSN-CNN"}[dn]] = {
                "recon_per_s": B * 3 / (ms_side / 1e3), "and_maxiters": 100, "batch": B}
            del dq
        # config 5 with the DnCNN-style denoiser: its backward solve runs on the masked-adjoint conv stack
        ts = train_step_bench(dev, rank, world, steps=3, warmup=3, batch=args.train_batch, denoiser="SimpleCNN", data=args.data)
        side["DE-GAP-CNN train_step"] = {"ms_per_step": ts["ms_per_step"], "batch": args.train_batch, "and_maxiters": 100,
                                         "backward_res": ts["backward_res"]}
        extras["side"] = side
        try:
            extras["gpu_eager_baseline"] = {
                "what": "the reference's call sequence as PyTorch eager/cuDNN ops on this GPU, batch 1 (scripts/bench_eager.py)",
                "fp32_recon_per_s": eager_gpu_bar(dev, solver, y_d, phi_d, 0),
                "tf32_recon_per_s": eager_gpu_bar(dev, solver, y_d, phi_d, 1)}
        except Exception as ex:                                   # a side number must not take the bench line down
            extras["gpu_eager_baseline"] = {"error": repr(ex)[:200]}

    if rank != 0:
        return
    peaks, peak_src = measured_peaks()
    hid_ms = ms_sum[2] / max(n_samp[2], 1)
    res_div = 2 if args.denoiser == "ffdnet" else 1              # FFDNet convs run at half resolution
    # hidden layers one launch covers: 1 (a kernel per layer) or the whole run (chained kernel, conv_tc2.cu)
    layers_per_launch = max(1, int(round(N_HIDDEN[args.denoiser] * n_launch[1] / n_launch[2]))) if n_launch[2] else 1
    flop_per_launch = 2 * 9 * 64 * 64 * (H // res_div) * (W // res_div) * T * B * layers_per_launch
    achieved = flop_per_launch / (hid_ms * 1e-3) / 1e12 if hid_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    kinds = ["gap", "conv_first", "conv_hidden", "conv_last", "anderson_gram", "anderson_solve", "anderson_mix"]
    avg = {kinds[i]: (ms_sum[i] / n_samp[i]) if n_samp[i] else None for i in range(k)}
    shares = {kinds[i]: {"launches": int(n_launch[i]), "sampled": int(n_samp[i]), "avg_ms": avg[kinds[i]],
                         "est_ms_per_step": (avg[kinds[i]] * n_launch[i] / args.steps) if n_samp[i] else None}
              for i in range(k)}
    hidden_kernel = (("conv_hidden_chain_kernel" if layers_per_launch > 1 else "conv_hidden_2cta_kernel")
                     if args.precision == "tc_split" else "conv_mid_tc_kernel")
    traffic, traffic_src = ncu_traffic("hidden", B) if (args.precision == "tc_split" and args.denoiser == "ffdnet") else (None, "n/a")
    if traffic is not None:
        # the committed capture is of one layer's launch (per-layer kernel) or of a whole run (chained kernel)
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")))["hidden"]
        cap_layers = int(d.get("layers_per_launch", 13 if "chain" in d.get("kernel", "") else 1))
        if cap_layers != layers_per_launch:
            traffic = traffic * layers_per_launch / cap_layers
            traffic_src += ", %d layer(s) per captured launch scaled to %d" % (cap_layers, layers_per_launch)
    # HBM-bound kernels (SURVEY 8(d) algorithmic bytes per measurement): GAP step 6,815,744 B per f call;
    # Anderson update (n = m = 5, beta = 1) 25,165,824 B per iteration over its three launches
    roof_hbm = {}
    if avg["gap"]:
        a = 6815744 * B / (avg["gap"] * 1e-3) / 1e9
        roof_hbm["gap_prep"] = {"bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
                                "algorithmic_bytes_per_launch": 6815744 * B, "avg_launch_ms": avg["gap"],
                                "traffic": ncu_traffic("prep", B)[0],
                                "note": "also writes the first conv layer's fp16 operand plane (4 MiB per measurement) "
                                        "that SURVEY's 104 B/pixel does not count"}
    if avg["anderson_gram"] and avg["anderson_mix"] and avg["anderson_solve"]:
        t_and = avg["anderson_gram"] + avg["anderson_solve"] + avg["anderson_mix"]
        a = 25165824 * B / (t_and * 1e-3) / 1e9
        tg, tm = ncu_traffic("gram", B)[0], ncu_traffic("mix", B)[0]
        roof_hbm["anderson"] = {"bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
                                "algorithmic_bytes_per_iteration": 25165824 * B, "avg_iteration_ms": t_and,
                                "launches": ["anderson_gram_kernel", "anderson_solve_kernel", "anderson_mix_kernel"],
                                "traffic": (tg + tm) if (tg and tm) else None}
    line = {
        "metric": "DE-GAP-FFDnet reconstructions/s (256x256x8, 180 Anderson iterations)" if args.denoiser == "ffdnet"
                  else "DE-GAP-%s reconstructions/s (256x256x8, %d Anderson iterations)" % (args.denoiser, args.max_iter),
        "value": value, "unit": "recon/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": data_text(args.data),
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
        "ranks": {"ms_per_step": [m / args.steps for m in ms_ranks], "min_ms": min(ms_ranks) / args.steps,
                  "max_ms": max(ms_ranks) / args.steps},
        "e2e": {"value": e2e_value, "unit": "recon/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(y_h.numel() * 4 + phi_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel() * 4)},
        "gpu_launches": int(sum(n_launch)),
        "roofline": {"kernel": hidden_kernel + " (hidden 64->64 3x3 layer, tcgen05 cta_group::2)"
                               if args.precision == "tc_split" else hidden_kernel + " (hidden layer)",
                     "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_unit": "dram bytes per launch", "traffic_source": traffic_src,
                     "peak_source": peak_src + ", bf16 dense sustained",
                     "avg_launch_ms": hid_ms, "algorithmic_flop_per_launch": flop_per_launch,
                     "hidden_layers_per_launch": layers_per_launch,
                     "issued_mma_flop_factor": 3 if args.precision == "tc_split" else 1,
                     "frac_issued": (3 if args.precision == "tc_split" else 1) * achieved / peak if peak else None,
                     "sampled_launches": int(n_samp[2])},
        "roofline_hbm": roof_hbm,
        "kernels": shares,
        "check": {"finite": finite, "psnr_vs_synthetic_gt_db": psnr, "sigma_schedule_calls_last_recon": sigma_calls,
                  "note": "PSNR of the last timed batch against its synthetic ground truth (a fresh reconstruction: the "
                          "two batches alternate, so the sigma schedule restarts every step)"},
    }
    line.update(extras)
    if lat_ms is not None:
        line["latency_batch1"] = {"ms_per_recon": lat_ms, "recon_per_s": 1e3 / lat_ms,
                                  "note": "same call at batch 1 (the reference's own test mode), inputs resident"}
    if world == 1 and not args.no_cpu_baseline and args.denoiser == "ffdnet":
        r = cpu_sample(args.cpu_iters, y_h, phi_h)
        line["cpu_baseline"] = {"value": r["value"], "unit": "recon/s", "cores": r["threads"], "kind": r["kind"],
                                "sample": r["sample"]}
        # parity of the benchmarked workload: the same andersonexp call (measurement 0, cpu_iters iterations)
        # through this repo's public solver API on the GPU against the CPU arm's iterate
        from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq_utils
        from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
        s2, _ = build_deq(dev, args.precision, "ffdnet", args.cpu_iters)
        with torch.no_grad():
            y1, p1 = y_d[:1].contiguous(), phi_d[:1].contiguous()
            ps1 = Phi_sum_(p1)
            zg, _ = eq_utils.andersonexp(lambda q, out=None: s2(q, y1, p1, ps1), At_torch_(y1, p1), m=M_HIST, lam=1e-2,
                                         max_iter=args.cpu_iters, tol=1e-5, beta=1.0)
        zc = torch.from_numpy(np.ascontiguousarray(r["z"])).double()
        rel = float((zg.cpu().double() - zc).norm() / zc.norm())
        line["check"]["rel_l2_vs_cpu_%s_iter%d" % (r["kind"], args.cpu_iters)] = rel
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
    ap.add_argument("--data", default=DATA_KIND, choices=sorted(DATA_TEXT), help="synthetic ground-truth kind")
    ap.add_argument("--no-extras", action="store_true", help="skip train_step / side / gpu_eager_baseline blocks")
    ap.add_argument("--profile-mode", action="store_true", help="resident steps only (for runs under ncu)")
    ap.add_argument("--train-steps", type=int, default=5)
    ap.add_argument("--train-batch", type=int, default=2, help="measurements per GPU per training step")
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
