#!/usr/bin/env python
"""CPU emulation of candidate tensor-core operand formats for the 64->64 hidden convolutions: runs the
oracle's DE-GAP solve with the hidden convs' operands quantised as a kernel would see them and reports
the per-iterate relative L2 distance and the PSNR shift against the fp32 run.  Design-time tool only
(decides whether a cheaper operand split can hold the parity bar); it lives under tests/ because it runs
the oracle, which only test infrastructure may do.  Nothing in the product imports it.

    python tests/tools/emulate_precision.py ffdnet drop8 128 40
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_scene, load_weights  # noqa: E402
from oracle import deqsci_oracle as orc  # noqa: E402

S = 2048.0


def q(x, dt):
    return x.to(dt).to(torch.float32)


def pow2_scale(t, target):
    m = float(t.abs().max())
    return 2.0 ** np.floor(np.log2(target / m)) if m > 0 else 1.0


def conv(x, w):
    return F.conv2d(x, w, padding=1)


def make_conv(mode):
    def conv3x3(x, w):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        w = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
        if mode == "first_kpack":
            # hidden layers as the product runs them (3-product split); the FIRST layer with the K-packing idea of
            # [Ah | Ah | Al'] against [Wh | Wl | Wh * 2^-11], correction weights at true scale
            # in fp16 (subnormals included), everything in one accumulator
            ah, wh = q(x, torch.float16), q(w, torch.float16)
            al = q((x - ah) * S, torch.float16)
            if x.shape[1] == 64 and w.shape[0] == 64:
                wl = q((w - wh) * S, torch.float16)
                return (conv(ah, wh) + (conv(ah, wl) + conv(al, wh)) / S).numpy()
            if w.shape[0] == 64:                      # first layer (cin = 5 or 1)
                wl_true = q(w - wh, torch.float16)
                wh_small = q(wh / S, torch.float16)
                return (conv(ah, wh) + conv(ah, wl_true) + conv(al, wh_small)).numpy()
            return conv(x, w).numpy()
        if mode == "fp32" or x.shape[1] != 64 or w.shape[0] != 64:
            return conv(x, w).numpy()
        ah, wh = q(x, torch.float16), q(w, torch.float16)
        if mode == "single":
            return conv(ah, wh).numpy()
        al, wl = q((x - ah) * S, torch.float16), q((w - wh) * S, torch.float16)
        if mode == "split":
            return (conv(ah, wh) + (conv(ah, wl) + conv(al, wh)) / S).numpy()
        if mode.startswith("split_lo"):   # lo' halves keep only the top n explicit mantissa bits (split_lo5 = 5 bits)
            keep = int(mode[len("split_lo"):])
            mask = torch.tensor((0xFFFF << (10 - keep)) & 0xFFFF, dtype=torch.int32)

            def trunc(t):
                b = t.to(torch.float16).view(torch.int16).to(torch.int32) & 0xFFFF
                return (b & mask).to(torch.int16).view(torch.float16).to(torch.float32)
            al_t, wl_t = trunc(al), trunc(wl)
            return (conv(ah, wh) + (conv(ah, wl_t) + conv(al_t, wh)) / S).numpy()
        if mode == "weights_only":       # Ah*Wh + Ah*Wl: activations rounded to fp16, weights carried in full
            return (conv(ah, wh) + conv(ah, wl) / S).numpy()
        if mode == "acts_only":          # Ah*Wh + Al*Wh
            return (conv(ah, wh) + conv(al, wh) / S).numpy()
        a_dt = torch.float8_e5m2 if "a52" in mode else torch.float8_e4m3fn
        sw1, sw2 = pow2_scale(wh, 256.0), pow2_scale(wl, 256.0)
        sa1 = sa2 = 1.0
        if a_dt == torch.float8_e4m3fn:
            sa1, sa2 = pow2_scale(al, 256.0), pow2_scale(ah, 256.0)
        al8, ah8 = q(al * sa1, a_dt) / sa1, q(ah * sa2, a_dt) / sa2
        wh8, wl8 = q(wh * sw1, torch.float8_e4m3fn) / sw1, q(wl * sw2, torch.float8_e4m3fn) / sw2
        return (conv(ah, wh) + (conv(ah8, wl8) + conv(al8, wh8)) / S).numpy()
    return conv3x3


def run(mode, net, scene, crop, iters, fi=0):
    gt, mask, meas = load_scene(scene)
    gt, mask, meas = gt[:crop, :crop], mask[:crop, :crop], meas[:crop, :crop]
    orc.conv3x3 = make_conv(mode)
    f = orc.ProxGradSCI(net, load_weights(net))
    y, Phi = meas[None, :, :, fi], mask[None]
    trace = []
    z, res = orc.deq_forward(f, y, Phi, orc.phi_sum(Phi), x0=orc.At(y, Phi), m=5, beta=1.0, lam=1e-2,
                             max_iter=iters, tol=1e-5, trace=trace)
    return z, orc.psnr(gt[None, :, :, fi * 8:(fi + 1) * 8], z.clip(0, 1)), trace


if __name__ == "__main__":
    net, scene = sys.argv[1], sys.argv[2]
    crop, iters = int(sys.argv[3]), int(sys.argv[4])
    modes = sys.argv[5:] or ["single", "split", "fp8_a52", "fp8_a43"]
    z0, p0, t0 = run("fp32", net, scene, crop, iters)
    print("fp32: psnr %.4f, iterates traced %d" % (p0, len(t0)))
    for mode in modes:
        z, p, t = run(mode, net, scene, crop, iters)
        rel = [float(np.linalg.norm((a - b).astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))
               for a, b in zip(t, t0)]
        print("%-10s psnr %.4f (%+.4f dB)  final rel %.2e  per-iterate rel: max %.2e  @1 %.2e @10 %.2e @last %.2e"
              % (mode, p, p - p0, float(np.linalg.norm(z - z0) / np.linalg.norm(z0)), max(rel), rel[0],
                 rel[min(10, len(rel) - 1)], rel[-1]), flush=True)
