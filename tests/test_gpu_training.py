"""GPU (-m gpu): one implicit-differentiation training step (config 5) against the reference's own
loss and parameter gradients (tests/golden/train_vectors.npz: 32x32x8 crop, B=2, max_iter=12,
denoiser in train mode).  Forward solve: native kernels where the denoiser is stateless in train mode
(DE-GAP-CNN), PyTorch ops with batch-statistics BatchNorm for FFDNet; backward solve: Anderson kernels
on the GAP-projector VJP kernel (ffdnet) or the autograd VJP (denoiser tag).  Bar: gradients within
1e-3 relative L2 of the reference."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_weights, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def train_vectors():
    return dict(np.load(os.path.join(GOLDEN, "train_vectors.npz")))


def _step(d, v, dev):
    from test_gpu_parity import build_solver
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    solver = build_solver(d, dev)
    solver.train()
    solver.nonlinear_op.train()
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=12, tol=1e-5)
    gt, Phi, y = (torch.from_numpy(v[k]).to(dev) for k in ("gt", "Phi", "y"))
    Ps = Phi_sum_(Phi)
    rec = deq.forward(y, Phi, Ps, initial_point=At_torch_(y, Phi))
    loss = torch.nn.MSELoss(reduction="mean")(rec, gt)
    loss.backward()
    return solver, deq, rec, loss


@pytest.fixture(scope="module")
def train_wide_vectors():
    return dict(np.load(os.path.join(GOLDEN, "train_wide_vectors.npz")))


@pytest.mark.parametrize("d", ["SimpleCNN", "ffdnet"])
@pytest.mark.parametrize("wide", [False, True])
def test_training_step_gradients_vs_reference(train_vectors, train_wide_vectors, d, wide):
    """wide=True: 32x160x8 crops, where FFDNet's train-mode forward solve runs on the native kernels
    (batch-statistics BatchNorm, running statistics updated per call); wide=False: 32x32x8 crops, where
    it runs on PyTorch ops.  Same bar either way."""
    v = train_wide_vectors if wide else train_vectors
    dev = torch.device("cuda", 0)
    solver, deq, rec, loss = _step(d, v, dev)
    for k in v:                                        # BatchNorm running statistics after the step
        if k.startswith("buf_%s::" % d):
            got_b = dict(solver.named_buffers())[k.split("::", 1)[1]].cpu().numpy()
            if k.endswith("num_batches_tracked"):
                assert int(got_b) == int(v[k])
            else:
                assert rel_l2(got_b, v[k]) <= 1e-3, k
    assert rel_l2(rec.detach().cpu().numpy(), v["rec_" + d]) <= 1e-3
    assert abs(float(loss) - float(v["loss_" + d])) <= 1e-3 * float(v["loss_" + d])
    assert abs(deq.forward_res - float(v["fres_" + d])) <= 1e-2 * float(v["fres_" + d])
    assert abs(deq.backward_res - float(v["bres_" + d])) <= 1e-2 * float(v["bres_" + d])
    names = [str(n) for n in v["gradnames_" + d]]
    got = dict(solver.named_parameters())
    assert sorted(names) == sorted(n for n, p in got.items() if p.grad is not None)
    norms = np.array([float(got[n].grad.norm()) for n in names])
    np.testing.assert_allclose(norms, v["gradnorms_" + d], rtol=2e-3, atol=1e-7)
    checked = 0
    for k in v:
        if k.startswith("grad_%s::" % d):
            n = k.split("::", 1)[1]
            assert rel_l2(got[n].grad.cpu().numpy(), v[k]) <= 1e-3, n
            checked += 1
    assert checked >= 2


def test_gradient_allreduce_two_gpus(train_vectors):
    """NCCL flat-bucket gradient average over 2 ranks (only when the box has >= 2 GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "scripts", "train_step_nccl.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALLREDUCE_OK" in r.stdout


@pytest.mark.parametrize("kind", ["ffdnet", "dncnn_bn"])
def test_train_mode_batchnorm_native_vs_torch(kind):
    """One train-mode iterate-map call under no_grad (what the DEQ forward solve does while training):
    native kernels with batch-statistics BatchNorm (deqsci_iterate_train) vs the PyTorch evaluation of
    the same module, incl. the running statistics both leave behind after two calls."""
    import copy
    from test_gpu_parity import build_solver
    from deqsci_b200.networks.provable.model.models import DnCNN
    from deqsci_b200.solvers.equilibrium_solvers_yaping import EquilibriumProxGradSCI
    from deqsci_b200.utils.cg_utils import A_torch_, At_torch_, Phi_sum_
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if kind == "ffdnet":
        a = build_solver("ffdnet", dev)
    else:
        net = DnCNN(channels=1, num_of_layers=6)
        a = EquilibriumProxGradSCI(A_torch_, At_torch_, net, 0.2).to(dev)
    a.train()
    b = copy.deepcopy(a)
    g = torch.Generator().manual_seed(2)
    shape = (2, 32, 160, 8) if kind == "ffdnet" else (2, 16, 136, 8)
    z = torch.rand(shape, generator=g).to(dev)
    Phi = (torch.rand(shape, generator=g) < 0.5).float().to(dev)
    y = A_torch_(torch.rand(shape, generator=g).to(dev), Phi)
    Ps = Phi_sum_(Phi)
    with torch.no_grad():
        assert a.nonlinear_op.native_train_ok(z)
        n1 = a(z, y, Phi, Ps)
        n2 = a(n1, y, Phi, Ps)
        t1 = b._autograd_forward(z, y, Phi, Ps)
        t2 = b._autograd_forward(t1, y, Phi, Ps)
    assert rel_l2(n1.cpu().numpy(), t1.cpu().numpy()) <= 2e-5
    assert rel_l2(n2.cpu().numpy(), t2.cpu().numpy()) <= 5e-5
    bn_a = [m for m in a.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    bn_b = [m for m in b.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    assert len(bn_a) == len(bn_b) > 0
    for ma, mb in zip(bn_a, bn_b):
        assert int(ma.num_batches_tracked) == int(mb.num_batches_tracked) == (2 if kind == "dncnn_bn" else int(mb.num_batches_tracked))
        assert rel_l2(ma.running_mean.cpu().numpy(), mb.running_mean.cpu().numpy()) <= 1e-4
        assert rel_l2(ma.running_var.cpu().numpy(), mb.running_var.cpu().numpy()) <= 1e-4
    if kind == "ffdnet":
        assert a._n == b._n == 2
