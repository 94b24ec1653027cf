#!/usr/bin/env python
"""Design-time emulation of candidate tensor-core operand formats for the 64->64 hidden convolutions,
on the GPU box with plain PyTorch (cuDNN fp32 convolutions of pre-rounded operands, TF32 off): the whole
DE-GAP solve (all iterate-map evaluations) of every real benchmark measurement, per-iterate relative L2
and PSNR shift against the fp32 run of the same code.  Decides whether a cheaper split than fp16 x 3
holds the parity bar (per-iterate <= 1e-3, |dPSNR| <= 0.05 dB) before any kernel is written
(VERDICT r01 item 3).  Nothing in the product imports this file.

    python tests/tools/emulate_gpu.py --net ffdnet --modes split,fp8,hyb_a8 --out gpurun_out/emul.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_scene  # noqa: E402
import bench  # noqa: E402

S = 2048.0
E4 = torch.float8_e4m3fn
E5 = torch.float8_e5m2


def q16(x):
    return x.half().float()


def q8(x, scale, dt=E4):
    """fp8 value grid (round to nearest even, saturating) at a power-of-two scale."""
    lim = 448.0 if dt == E4 else 57344.0
    return (x * scale).clamp(-lim, lim).to(dt).float() / scale


def qi8(x, step):
    return torch.round(x / step).clamp(-127, 127) * step


def pow2_below(t, target):
    m = t.abs().amax()
    return torch.exp2(torch.floor(torch.log2(target / m.clamp_min(1e-30))))


def col_scale(w, target=256.0):
    """per-output-channel power-of-two scale [O,1,1,1]: max|w[o]| * s in [target/2, target)."""
    m = w.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
    return torch.exp2(torch.floor(torch.log2(target / m)))


def conv(x, w):
    return F.conv2d(x, w, padding=1)


# ---- Winograd F(2x2, 3x3): 16 multiplies per 2x2 output tile and (cin, cout) pair instead of 36 ----------------
_BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
_G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float32)
_AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)


def winograd_conv(x, w, split):
    """3x3 pad-1 convolution through F(2x2,3x3) with fp32 transforms; the 16 batched [tiles x cin] . [cin x cout]
    products run either in fp32 (split=False: transform rounding alone) or as the product's fp16 x 3 scheme on the
    TRANSFORMED operands (what a tensor-core Winograd kernel would issue)."""
    N, C, H, W = x.shape
    O = w.shape[0]
    dev = x.device
    BT, G, AT = _BT.to(dev), _G.to(dev), _AT.to(dev)
    U = torch.einsum("ij,ocjk,lk->ocil", G, w, G)                       # [O,C,4,4]
    xp = F.pad(x, (1, 1, 1, 1))
    tiles = xp.unfold(2, 4, 2).unfold(3, 4, 2)                          # [N,C,H/2,W/2,4,4]
    V = torch.einsum("ij,nchwjk,lk->nchwil", BT, tiles, BT)             # [N,C,th,tw,4,4]
    th, tw = V.shape[2], V.shape[3]
    Vm = V.permute(4, 5, 0, 2, 3, 1).reshape(16, N * th * tw, C)        # [16, tiles, C]
    Um = U.permute(2, 3, 1, 0).reshape(16, C, O)                        # [16, C, O]
    if split:
        vh, uh = q16(Vm), q16(Um)
        vl, ul = q16((Vm - vh) * S), q16((Um - uh) * S)
        M = torch.bmm(vh, uh) + (torch.bmm(vh, ul) + torch.bmm(vl, uh)) / S
    else:
        M = torch.bmm(Vm, Um)
    M = M.reshape(4, 4, N, th, tw, O)
    Y = torch.einsum("ij,jknhwo,lk->nohiwl", AT, M, AT)                 # [N,O,th,2,tw,2]
    return Y.reshape(N, O, th * 2, tw * 2)


class Emu:
    """conv3x3(x, w) for hidden layers under a named operand scheme."""

    def __init__(self, mode, act_scale=4.0):
        self.mode = mode
        self.act_scale = act_scale
        self.wcache = {}

    def weights(self, w):
        k = w.data_ptr()
        if k not in self.wcache:
            wh = q16(w)
            wl = q16((w - wh) * S)                 # lo' = lo * 2^11 in fp16
            sc = col_scale(wh)
            d = {"wh": wh, "wl": wl, "wh8": q8(wh, sc), "wl8": q8(wl, col_scale(wl)),
                 "wh8_e5": q8(wh, 1.0, E5), "wl8_e5": q8(wl, 1.0, E5)}
            # int8 with per-output-channel step
            d["wl_i8"] = qi8(wl, wl.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30) / 127)
            d["wh_i8"] = qi8(wh, wh.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30) / 127)
            # two-term fp8 weights (second-order terms), for the K-extended variants
            d["wh8b"] = q8(wh - d["wh8"], col_scale(wh - d["wh8"]))
            d["wl8b"] = q8(wl - d["wl8"], col_scale(wl - d["wl8"]))
            self.wcache[k] = d
        return self.wcache[k]

    def __call__(self, x, w):
        m = self.mode
        if m == "fp32":
            return conv(x, w)
        if m == "jitter":                             # fp32 noise floor: 6e-8 relative input jitter
            return conv(x * (1 + 6e-8 * torch.randn_like(x)), w)
        if m == "wino_fp32":
            return winograd_conv(x, w, False)
        if m == "wino_split":
            return winograd_conv(x, w, True)
        d = self.weights(w)
        ah = q16(x)
        main = conv(ah, d["wh"])
        if m == "single":
            return main
        al = q16((x - ah) * S)
        sa = self.act_scale
        if m == "split":                              # today's product path: fp16 x 3
            return main + (conv(ah, d["wl"]) + conv(al, d["wh"])) / S
        if m == "fp8":                                # both corrections in e4m3
            return main + (conv(q8(ah, sa), d["wl8"]) + conv(q8(al, sa), d["wh8"])) / S
        if m == "fp8_e5a":                            # activations e5m2, weights e4m3
            return main + (conv(q8(ah, 1.0, E5), d["wl8"]) + conv(q8(al, 1.0, E5), d["wh8"])) / S
        if m == "hyb_a8":                             # Ah.[Wh|Wl'] in fp16 (N=128), Al'.Wh in e4m3
            return main + (conv(ah, d["wl"]) + conv(q8(al, sa), d["wh8"])) / S
        if m == "hyb_w8":                             # Ah.Wh + Al'.Wh in fp16, Ah.Wl' in e4m3
            return main + (conv(q8(ah, sa), d["wl8"]) + conv(al, d["wh"])) / S
        if m == "fp8_w2":                             # e4m3 corrections, weights carried as two e4m3 terms (K x 2)
            return main + (conv(q8(ah, sa), d["wl8"] + d["wl8b"]) + conv(q8(al, sa), d["wh8"] + d["wh8b"])) / S
        if m == "fp8_a2":                             # e4m3 corrections, activations carried as two e4m3 terms
            ah8 = q8(ah, sa)
            al8 = q8(al, sa)
            ah8b = q8(ah - ah8, sa * 16)
            al8b = q8(al - al8, sa * 16)
            return main + (conv(ah8 + ah8b, d["wl8"]) + conv(al8 + al8b, d["wh8"])) / S
        if m == "fp8_wi8":                            # activations e4m3 ... not a real instruction mix; bound only
            return main + (conv(q8(ah, sa), d["wl_i8"]) + conv(q8(al, sa), d["wh_i8"])) / S
        if m == "i8":                                 # both corrections int8: global activation step, per-channel weights
            sh = ah.abs().amax().clamp_min(1e-30) / 127
            sl = al.abs().amax().clamp_min(1e-30) / 127
            return main + (conv(qi8(ah, sh), d["wl_i8"]) + conv(qi8(al, sl), d["wh_i8"])) / S
        if m == "acts_only":                          # drop Ah.Wl' (weights rounded to fp16)
            return main + conv(al, d["wh"]) / S
        if m == "weights_only":                       # drop Al'.Wh (activations rounded to fp16)
            return main + conv(ah, d["wl"]) / S
        raise SystemExit("unknown mode " + m)


def build(net_name, dev):
    solver, _ = bench.build_deq(dev, "fp32", net_name, 180 if net_name == "ffdnet" else 100)
    op = solver.nonlinear_op
    if net_name == "ffdnet":
        seq = op.intermediate_dncnn.itermediate_dncnn
    else:
        seq = op.dncnn
    layers = list(seq)
    return layers


def run_stack(layers, x, emu):
    for L in layers:
        if isinstance(L, nn.Conv2d):
            w = L.weight
            if w.shape[0] == 64 and w.shape[1] == 64:
                x = emu(x, w)
            else:
                x = conv(x, w)
        else:
            x = L(x)
    return x


def solve(layers, net_name, y, Phi, emu, max_iter, trace):
    dev = y.device
    Phi_sum = torch.sum(Phi, dim=3)
    Phi_sum[Phi_sum == 0] = 1
    st = {"k": 0}

    def f(z):
        B, H, W, T = z.shape
        fb = torch.sum(z * Phi, dim=3)
        z = z + ((y - fb) / Phi_sum)[:, :, :, None] * Phi
        x = z.permute(0, 3, 1, 2).contiguous().view(B * T, 1, H, W)
        if net_name == "ffdnet":
            sig = np.float32(60 / 255)
            for _ in range(st["k"]):
                sig = np.float32(sig * np.float32(0.971))
            inp = torch.cat((torch.full((B * T, 1, H // 2, W // 2), float(sig), device=dev), F.pixel_unshuffle(x, 2)), 1)
            noise = F.pixel_shuffle(run_stack(layers, inp, emu), 2)
        else:
            noise = run_stack(layers, x, emu)
        st["k"] += 1
        out = z - noise.view(B, T, H, W).permute(0, 2, 3, 1)
        trace.append(out)
        return out

    m, lam, beta = 5, 1e-2, 1.0
    x0 = y[:, :, :, None] * Phi
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
    for k in range(2, max_iter):
        n = min(k, m)
        G = Fh[:, :n] - X[:, :n]
        Hm[:, 1:n + 1, 1:n + 1] = torch.bmm(G, G.transpose(1, 2)) + lam * torch.eye(n, device=dev)[None]
        alpha = torch.linalg.solve(Hm[:, :n + 1, :n + 1], rhs[:, :n + 1])[:, 1:n + 1, 0]
        X[:, k % m] = beta * (alpha[:, None] @ Fh[:, :n])[:, 0]
        Fh[:, k % m] = f(X[:, k % m].reshape(x0.shape)).reshape(bsz, -1)
    return f(X[:, k % m].reshape(x0.shape))


def psnr(gt, z):
    return float(10 * torch.log10(1.0 / ((z.clip(0, 1) - gt) ** 2).mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="ffdnet")
    ap.add_argument("--modes", default="jitter,split,single,fp8,fp8_e5a,hyb_a8,hyb_w8,fp8_w2,fp8_a2,i8,acts_only,weights_only")
    ap.add_argument("--max-iter", type=int, default=None)
    ap.add_argument("--synthetic", type=int, default=2, help="also run this many synthetic bench measurements")
    ap.add_argument("--synthetic-kinds", default="uniform")
    ap.add_argument("--no-real", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--limit", type=int, default=0, help="only the first n cases")
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    max_iter = a.max_iter or (180 if a.net == "ffdnet" else 100)
    layers = build(a.net, dev)
    cases = []
    for scene in (() if a.no_real else ("drop8", "runner8", "traffic")):
        gt, mask, meas = load_scene(scene)
        n = 1 if scene in ("drop8", "runner8") else meas.shape[2]
        for fi in range(n):
            cases.append((scene + ":%d" % fi, torch.from_numpy(meas[None, :, :, fi]).to(dev), torch.from_numpy(mask[None]).to(dev),
                          torch.from_numpy(gt[None, :, :, fi * 8:(fi + 1) * 8]).to(dev)))
    if a.synthetic:
        for kind in a.synthetic_kinds.split(","):
            ys, ps, xs = bench.synthetic_batch(0, a.synthetic, kind)
            for i in range(a.synthetic):
                cases.append(("%s:%d" % (kind, i), ys[i:i + 1].to(dev), ps[i:i + 1].to(dev), xs[i:i + 1].to(dev)))
    if a.limit:
        cases = cases[:a.limit]
    results = {}
    with torch.no_grad():
        for name, y, Phi, gt in cases:
            t0 = []
            z0 = solve(layers, a.net, y, Phi, Emu("fp32"), max_iter, t0)
            p0 = psnr(gt, z0)
            n0 = [float(t.double().norm()) for t in t0]
            results[name] = {"fp32_psnr": p0}
            print("%-14s fp32 psnr %.4f" % (name, p0), flush=True)
            for mode in a.modes.split(","):
                t = []
                z = solve(layers, a.net, y, Phi, Emu(mode), max_iter, t)
                rel = [float((u.double() - v.double()).norm()) / nv for u, v, nv in zip(t, t0, n0)]
                r = {"dpsnr": psnr(gt, z) - p0, "max_rel": max(rel), "rel@40": rel[min(40, len(rel) - 1)], "rel@last": rel[-1],
                     "argmax": int(np.argmax(rel))}
                results[name][mode] = r
                print("   %-12s dPSNR %+.4f  per-iterate max %.2e (@%d)  @40 %.2e  last %.2e" % (
                    mode, r["dpsnr"], r["max_rel"], r["argmax"], r["rel@40"], r["rel@last"]), flush=True)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(results, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
