"""Measures how reproducible the reference's own fp32 trajectory is between two fp32
implementations: runs the numpy oracle (different summation order than PyTorch's conv) for the full
180 iterations on traffic measurements with the FFDNet stand-in weights and records the PSNR and
per-call input norms next to the reference's (tests/golden/full_recon.npz).  CPU, ~4 min each."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_scene, load_weights  # noqa: E402
from oracle import deqsci_oracle as orc  # noqa: E402

full = dict(np.load(os.path.join(HERE, "full_recon.npz")))
out = {}
gt, mask, meas = load_scene("traffic")
for fi in [int(a) for a in sys.argv[1:]] or [1, 2]:
    f = orc.ProxGradSCI("ffdnet", load_weights("ffdnet"))
    seen = []
    fm = lambda z, *a: (seen.append(float(np.linalg.norm(z.astype(np.float64)))), f(z, *a))[1]
    y, Phi = meas[None, :, :, fi], mask[None]
    t0 = time.time()
    z, res = orc.deq_forward(fm, y, Phi, orc.phi_sum(Phi), x0=orc.At(y, Phi), m=5, beta=1.0, lam=1e-2,
                             max_iter=180, tol=1e-5)
    psnr = orc.psnr(gt[None, :, :, fi * 8:(fi + 1) * 8], z.clip(0, 1))
    ref = float(full["ffdnet_traffic_%d_psnr" % fi])
    nr = full["ffdnet_traffic_%d_innorm" % fi]
    rel = np.abs(np.array(seen) - nr) / nr
    out["traffic_%d_oracle_psnr" % fi] = np.array(psnr)
    out["traffic_%d_reference_psnr" % fi] = np.array(ref)
    out["traffic_%d_norm_rel_dev" % fi] = rel
    print("traffic_%d: oracle %.4f dB, reference %.4f dB, diff %+.4f; first call with norm dev > 1e-4: %d; max %.2e (%.0fs)"
          % (fi, psnr, ref, psnr - ref, int(np.argmax(rel > 1e-4)) if (rel > 1e-4).any() else -1, rel.max(),
             time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "noise_floor.npz"), **out)
