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


def oracle_vs_reference(fis):
    path = os.path.join(HERE, "noise_floor.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    gt, mask, meas = load_scene("traffic")
    for fi in fis:
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
        np.savez_compressed(path, **out)



def reference_vs_reference(cases, threads):
    """the spread of the REFERENCE against ITSELF.  Re-runs the unmodified reference
    (oracle/ref_import.py shims, CPU) on the ill-conditioned measurements with the measurement perturbed by
    6e-8 relative (half an fp32 ulp: the same data to within float rounding) and, optionally, another
    thread count, and records PSNR / per-call input-norm deviations against the unperturbed reference run
    stored in full_recon.npz.  ~150 s per measurement on 8 threads."""
    import torch
    from oracle import ref_import
    ref_import.install_shims()
    from utils.cg_utils import At_torch_
    torch.set_num_threads(threads)
    path = os.path.join(HERE, "noise_floor.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for scene, fi in cases:
        key = "refjit_%s_%d" % (scene, fi)
        if key + "_psnr" in out:
            continue
        gt, mask, meas = load_scene(scene)
        g = torch.Generator().manual_seed(99 + fi)
        y = torch.from_numpy(meas[:, :, fi])[None]
        y = y * (1 + 6e-8 * torch.randn(y.shape, generator=g))
        Phi = torch.from_numpy(mask)[None]
        Phi_sum = torch.sum(Phi, axis=3)
        Phi_sum[Phi_sum == 0] = 1
        solver, deq = ref_import.build_reference_deq("ffdnet", max_iter=180)
        zin = []
        h = solver.register_forward_pre_hook(lambda mod, args: zin.append(float(args[0].detach().norm())))
        t0 = time.time()
        z = deq.forward(y, Phi, Phi_sum, initial_point=At_torch_(y, Phi), train_flag=False).detach()
        h.remove()
        psnr = ref_import.skimage_psnr(gt[None, :, :, fi * 8:(fi + 1) * 8], z.clip(0, 1).numpy())
        ref = float(full["ffdnet_%s_%d_psnr" % (scene, fi)])
        nr = full["ffdnet_%s_%d_innorm" % (scene, fi)]
        rel = np.abs(np.array(zin) - nr) / nr
        out[key + "_psnr"] = np.array(psnr)
        out[key + "_reference_psnr"] = np.array(ref)
        out[key + "_norm_rel_dev"] = rel
        out[key + "_threads"] = np.array(threads)
        print("%s: jittered reference %.4f dB, reference %.4f dB, diff %+.4f; max norm dev %.2e (%.0fs, %d threads)"
              % (key, psnr, ref, psnr - ref, rel.max(), time.time() - t0, threads), flush=True)
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "reference":
        threads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
        sel = sys.argv[3:] or ["traffic:2", "traffic:1", "runner8:0", "traffic:0", "traffic:3", "traffic:4", "traffic:5", "drop8:0"]
        reference_vs_reference([(c.split(":")[0], int(c.split(":")[1])) for c in sel], threads)
    else:
        oracle_vs_reference([int(a) for a in sys.argv[1:]] or [1, 2])
