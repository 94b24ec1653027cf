import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_scene, load_weights
from test_gpu_parity import build_solver, t
from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
from deqsci_b200.utils.metrics import peak_signal_noise_ratio
dev = torch.device("cuda", 0)
full = dict(np.load(os.path.join(ROOT, "tests/golden/full_recon.npz")))
gt, mask, meas = load_scene("traffic")
d = sys.argv[1] if len(sys.argv) > 1 else "ffdnet"
prec = sys.argv[2] if len(sys.argv) > 2 else None
max_iter = 180 if d == "ffdnet" else 100
def run(idx):
    solver = build_solver(d, dev, prec)
    Phi = t(np.stack([mask] * len(idx)), dev)
    y = t(np.stack([meas[:, :, i] for i in idx]), dev)
    Ps = Phi_sum_(Phi)
    seen = []
    h = solver.register_forward_pre_hook(lambda mod, args: seen.append(args[0].detach().clone()))
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=max_iter, tol=1e-5)
    z = deq.forward(y, Phi, Ps, initial_point=At_torch_(y, Phi), train_flag=False)
    h.remove()
    return z, seen, deq.forward_res
zb, seen_b, resb = run(list(range(6)))
for i in range(6):
    z1, seen1, res1 = run([i])
    g = gt[None, :, :, i * 8:(i + 1) * 8]
    p1 = peak_signal_noise_ratio(g, z1.clip(0, 1).cpu().numpy())
    pb = peak_signal_noise_ratio(g, zb[i:i + 1].clip(0, 1).cpu().numpy())
    ref = float(full["%s_traffic_%d_psnr" % (d, i)])
    n1 = np.array([float(s.double().norm()) for s in seen1])
    nb = np.array([float(s[i].double().norm()) for s in seen_b])
    nr = full["%s_traffic_%d_innorm" % (d, i)][:len(n1)]
    rel1 = np.abs(n1 - nr) / nr
    relb = np.abs(nb - nr) / nr
    first_bad = int(np.argmax(rel1 > 1e-4)) if (rel1 > 1e-4).any() else -1
    print("meas %d: psnr single %.4f batched %.4f ref %.4f | res %.3e ref %.3e | max norm dev single %.2e (first>1e-4 at %d) batched %.2e | z single-vs-batched rel %.2e"
          % (i, p1, pb, ref, res1, float(full["%s_traffic_%d_res" % (d, i)]), rel1.max(), first_bad, relb.max(),
             float((z1 - zb[i:i + 1]).norm() / z1.norm())))
