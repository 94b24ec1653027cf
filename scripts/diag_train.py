import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rel_l2
from test_gpu_training import _step
v = dict(np.load(os.path.join(ROOT, "tests/golden/train_vectors.npz")))
dev = torch.device("cuda", 0)
for d in ["SimpleCNN", "ffdnet"]:
    solver, deq, rec, loss = _step(d, v, dev)
    got = dict(solver.named_parameters())
    print(d, "rec rel", rel_l2(rec.detach().cpu().numpy(), v["rec_" + d]), "loss", float(loss), float(v["loss_" + d]),
          "fres", deq.forward_res, float(v["fres_" + d]), "bres", deq.backward_res, float(v["bres_" + d]))
    names = [str(n) for n in v["gradnames_" + d]]
    norms = np.array([float(got[n].grad.norm()) for n in names])
    print("  norm rel dev max", np.max(np.abs(norms - v["gradnorms_" + d]) / v["gradnorms_" + d]))
    for k in v:
        if k.startswith("grad_%s::" % d):
            n = k.split("::", 1)[1]
            print("  ", n, rel_l2(got[n].grad.cpu().numpy(), v[k]))
