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

from conftest import GOLDEN, ROOT, load_weights, rel_l2

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
    assert abs(float(loss.detach()) - float(v["loss_" + d])) <= 1e-3 * float(v["loss_" + d])
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


@pytest.mark.parametrize("kind", ["ffdnet", "dncnn_bn", "ffdnet_odd", "dncnn_bn_odd"])
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
    odd = kind.endswith("_odd")            # conv images of odd height: the pair kernel's last strip runs past the image
    kind = kind.replace("_odd", "")
    if kind == "ffdnet":
        a = build_solver("ffdnet", dev)
    else:
        net = DnCNN(channels=1, num_of_layers=6)
        a = EquilibriumProxGradSCI(A_torch_, At_torch_, net, 0.2).to(dev)
    a.train()
    b = copy.deepcopy(a)
    g = torch.Generator().manual_seed(2)
    shape = (2, 32, 160, 8) if kind == "ffdnet" else (2, 16, 136, 8)
    if odd:
        shape = (2, 38, 160, 8) if kind == "ffdnet" else (2, 19, 136, 8)
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


@pytest.mark.parametrize("d", ["SimpleCNN", "ffdnet"])
def test_train_driver_equals_python_loop(train_wide_vectors, d, monkeypatch):
    """The training forward solve through deqsci_reconstruct(_train) (one C-ABI call) and through the
    Python andersonexp loop queue the same kernels: reconstruction, gradients, BatchNorm running
    statistics, num_batches_tracked and the sigma call counter are identical."""
    dev = torch.device("cuda", 0)
    out = {}
    for driver in ("1", "0"):
        monkeypatch.setenv("DEQSCI_DRIVER", driver)
        solver, deq, rec, loss = _step(d, train_wide_vectors, dev)
        out[driver] = (rec.detach().clone(), {k: p.grad.clone() for k, p in solver.named_parameters() if p.grad is not None},
                       {k: b.clone() for k, b in solver.named_buffers()}, getattr(solver, "_n", None), deq.forward_res,
                       deq.backward_res)
    a, b = out["1"], out["0"]
    assert torch.equal(a[0], b[0])
    # gradients come out of cuDNN's backward kernels (atomics): equal up to their run-to-run noise.  SimpleCNN: with the
    # driver the backward solve's VJPs run on the native masked-adjoint stack, without it on cuDNN dgrads -- two fp32
    # implementations of the same 12-iteration solve (both within 1e-3 of the reference, test above)
    assert a[1].keys() == b[1].keys()
    for k in a[1]:
        assert rel_l2(a[1][k].cpu().numpy(), b[1][k].cpu().numpy()) <= (2e-4 if d == "SimpleCNN" else 1e-5), k
    assert all(torch.equal(a[2][k], b[2][k]) for k in a[2])
    assert a[3] == b[3]
    assert a[4] == b[4]
    if d == "ffdnet":                    # backward solve: deqsci_adjoint_solve vs the Python loop on the same kernels
        assert a[5] == b[5]
    else:
        assert abs(a[5] - b[5]) <= 1e-4 * abs(b[5])


@pytest.mark.parametrize("d", ["SimpleCNN", "ffdnet"])
def test_plan_refresh_on_device_equals_rebuild(d):
    """After an in-place parameter update (optimizer.step()) the cached plan is refreshed by the repack
    kernels (deqsci_denoiser_update_weights) -- same object, no host copy -- and computes exactly what a
    plan packed on the host from the new weights computes."""
    import copy
    from test_gpu_parity import build_solver
    dev = torch.device("cuda", 0)
    solver = build_solver(d, dev)
    op = solver.nonlinear_op
    train = d == "ffdnet"                   # FFDNet: the train plan (BatchNorm not folded) is the refreshable one
    g = torch.Generator(device="cpu").manual_seed(5)
    z = torch.rand(2, 32, 160, 8, generator=g).to(dev)
    y = torch.rand(2, 32, 160, generator=g).to(dev) * 4
    Phi = (torch.rand(2, 32, 160, 8, generator=g) > 0.5).float().to(dev)
    Ps = Phi.sum(3)
    Ps[Ps == 0] = 1

    def run(mod, plan):
        if train:
            return plan.iterate_train(z, y, Phi, Ps, 0.2, mod.bn_slots())
        return plan.iterate(z, y, Phi, Ps, 0.0)

    plan0 = op.native_plan(dev, train=train)
    before = run(op, plan0)
    with torch.no_grad():
        for i, p in enumerate(op.parameters()):
            p.mul_(1.0 + 0.01 * ((i % 3) - 1)).add_(1e-3)
    fresh = copy.deepcopy(op)               # identical weights and buffers, no cached plan: packed on the host
    assert not fresh.__dict__.get("_native_cache")
    plan1 = op.native_plan(dev, train=train)
    assert plan1 is plan0                   # refreshed in place
    got = run(op, plan1)
    want = run(fresh, fresh.native_plan(dev, train=train))
    assert fresh.native_plan(dev, train=train) is not plan0
    assert torch.equal(got, want)
    assert not torch.equal(got, before)


def test_fused_scale_adam_single_gpu():
    """GradientSynchronizer at world 1: flat in-place gradients + ONE fused scale + Adam kernel equals
    torch.optim.Adam step for step; an adopted torch optimizer keeps a meaningful state_dict (csrc/optim.cu)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fused_adam_check.py")], capture_output=True,
                       text=True, timeout=300)
    assert "FUSED_ADAM_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_fused_allreduce_adam_two_gpus():
    """Two ranks: the ONE-kernel exchange (cross-GPU barrier + one-shot all-reduce over NVLink peer memory + 1/world
    scale + Adam) and its NCCL variant both equal all-gather-mean + torch Adam, and the ranks' parameter copies stay
    bitwise identical."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29653",
                        os.path.join(ROOT, "scripts", "fused_adam_check.py")], capture_output=True, text=True, timeout=600)
    assert "FUSED_ADAM_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_masked_adjoint_vjp_vs_autograd(train_wide_vectors):
    """Tag 'denoiser': the VJP of the iterate map on the native kernels -- the conv stack with transposed, flipped weights
    gated by the saved activations' signs (deqsci_iterate_save / deqsci_denoise_residual_masked), then the GAP
    projector -- against torch.autograd.grad through the PyTorch evaluation of the same map
    (reference solvers/new_equilibrium_utils_yaping.py:271-277)."""
    from test_gpu_parity import build_solver
    from deqsci_b200 import ops
    from deqsci_b200.utils.cg_utils import Phi_sum_
    dev = torch.device("cuda", 0)
    v = train_wide_vectors
    solver = build_solver("SimpleCNN", dev)
    solver.train()
    op = solver.nonlinear_op
    gt, Phi, y = (torch.from_numpy(v[k]).to(dev) for k in ("gt", "Phi", "y"))
    Ps = Phi_sum_(Phi)
    g = torch.Generator().manual_seed(11)
    z0 = (gt + 0.05 * torch.randn(gt.shape, generator=g).to(dev)).requires_grad_()
    vec = torch.randn(gt.shape, generator=g).to(dev)
    assert op.native_adjoint_ok(z0)
    f0 = solver._autograd_forward(z0, y, Phi, Ps)
    want = torch.autograd.grad(f0, z0, vec)[0]
    with torch.no_grad():
        out, saved = op.native_plan(dev).iterate_save(z0.detach(), y, Phi, Ps, 0.0)
        acts = saved.acts
        assert rel_l2(out.cpu().numpy(), f0.detach().cpu().numpy()) <= 2e-5          # same forward values
        adj = op.native_adjoint_plan(dev)
        got = ops.gap_vjp(adj.denoise_residual_masked(vec, list(reversed(acts))), Phi, Ps)
    # against autograd: the ReLU masks come from two fp32 evaluations of the forward pass (cuDNN vs the native stack);
    # pre-activations within ~1e-6 of zero flip, each flip is an O(1) error on its path: ~2e-4 overall
    assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) <= 5e-4
    # against the same chain in PyTorch with the SAME masks (the saved native activations): arithmetic only
    import torch.nn.functional as F

    def torch_vjp(acts_, vv=None):
        vv = vec if vv is None else vv
        B, H, W, T = vv.shape
        convs = [m for m in op.dncnn if isinstance(m, torch.nn.Conv2d)]
        u = vv.permute(0, 3, 1, 2).reshape(B * T, 1, H, W)
        for i in range(len(convs) - 1, -1, -1):
            u = F.conv_transpose2d(u, convs[i].weight, padding=1)
            if i > 0:
                hi = acts_[i - 1].view(torch.float16).view(2, B * T, H, W, 64)[0]
                u = u * (hi > 0).permute(0, 3, 1, 2).float()
        r = vv - u.view(B, T, H, W).permute(0, 2, 3, 1)
        return ops.gap_vjp(r.contiguous(), Phi, Ps)

    with torch.no_grad():
        assert rel_l2(got.cpu().numpy(), torch_vjp(acts).cpu().numpy()) <= 2e-5
        # the whole backward solve on a gradient of realistic size (MSE over ~1e6 elements: ~1e-8 per element, far below
        # fp16's subnormal step): the conv stack runs on a power-of-two multiple and the result is scaled back
        from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
        tiny = (vec * 1e-8).contiguous()
        masks = list(reversed(acts))
        g_ref, _ = eq.andersonexp(lambda q: torch_vjp(acts, q) + tiny, tiny, m=5, lam=1e-2, max_iter=8, tol=1e-9, beta=1.0)
        g_nat, _ = adj.adjoint_solve(tiny, Phi, Ps, masks, m=5, lam=1e-2, max_iter=8, tol=1e-9, beta=1.0)
        g_raw, _ = adj.adjoint_solve(tiny, Phi, Ps, masks, m=5, lam=1e-2, max_iter=8, tol=1e-9, beta=1.0, vjp_scale=1.0)
        e_nat, e_raw = rel_l2(g_nat.cpu().numpy(), g_ref.cpu().numpy()), rel_l2(g_raw.cpu().numpy(), g_ref.cpu().numpy())
        assert e_nat <= 5e-5, (e_nat, e_raw)
        assert e_raw > 20 * e_nat, (e_nat, e_raw)             # without the scaling the operand planes underflow
    # after an in-place weight update the adjoint plan is refreshed on the device and still matches
    with torch.no_grad():
        for p_ in op.parameters():
            p_.mul_(1.02)
        acts1 = op.native_plan(dev).iterate_save(z0.detach(), y, Phi, Ps, 0.0)[1].acts
        adj1 = op.native_adjoint_plan(dev)
        assert adj1 is adj
        got1 = ops.gap_vjp(adj1.denoise_residual_masked(vec, list(reversed(acts1))), Phi, Ps)
        assert rel_l2(got1.cpu().numpy(), torch_vjp(acts1).cpu().numpy()) <= 2e-5
    assert rel_l2(got1.cpu().numpy(), got.cpu().numpy()) > 1e-3


@pytest.mark.parametrize("d", ["SimpleCNN", "ffdnet"])
def test_native_weight_gradients_vs_autograd(train_wide_vectors, d, monkeypatch):
    """The graph-attached iterate-map call as ONE native autograd node (deqsci_b200/backward.py: tensor-core forward that
    keeps the activations, csrc/backward.cu for ReLU / train-mode BatchNorm backward and wgrad, the adjoint plan's
    tensor-core dgrads) against the same call through PyTorch autograd / cuDNN: values, every parameter gradient,
    BatchNorm running statistics -- with an upstream gradient of realistic size (~1e-7 per element)."""
    import copy
    from test_gpu_parity import build_solver
    from deqsci_b200.backward import native_backward_ok
    from deqsci_b200.utils.cg_utils import Phi_sum_
    dev = torch.device("cuda", 0)
    v = train_wide_vectors
    a = build_solver(d, dev)
    a.train()
    a.nonlinear_op.train()
    b = copy.deepcopy(a)
    gt, Phi, y = (torch.from_numpy(v[k]).to(dev) for k in ("gt", "Phi", "y"))
    Ps = Phi_sum_(Phi)
    g = torch.Generator().manual_seed(5)
    z = gt + 0.05 * torch.randn(gt.shape, generator=g).to(dev)
    assert native_backward_ok(a, z, y, Phi, Ps)
    fa = a(z, y, Phi, Ps)                                   # native node
    assert type(fa.grad_fn).__name__.startswith("NativeIterate")
    up = (2.0 * (fa.detach() - gt) / gt.numel()).contiguous()      # the MSE loss's gradient: ~1e-7 per element
    (fa * up).sum().backward()
    monkeypatch.setenv("DEQSCI_NATIVE_BACKWARD", "0")
    fb = b(z, y, Phi, Ps)                                   # autograd / cuDNN
    assert not type(fb.grad_fn).__name__.startswith("NativeIterate")
    (fb * up).sum().backward()
    assert rel_l2(fa.detach().cpu().numpy(), fb.detach().cpu().numpy()) <= 2e-5
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    errs = {}
    for k in pb:
        assert pa[k].grad is not None and pb[k].grad is not None, k
        errs[k] = rel_l2(pa[k].grad.cpu().numpy(), pb[k].grad.cpu().numpy())
    print("parameter-gradient rel-L2 vs autograd:", {k.split("nonlinear_op.")[-1]: float("%.2e" % e) for k, e in errs.items()})
    assert max(errs.values()) <= 1e-3, errs
    for (ka, ba), (kb, bb) in zip(a.named_buffers(), b.named_buffers()):
        if ka.endswith("running_mean") or ka.endswith("running_var"):
            assert rel_l2(ba.cpu().numpy(), bb.cpu().numpy()) <= 1e-4, ka
        if ka.endswith("num_batches_tracked"):
            assert int(ba) == int(bb)


def test_train_solver_sci_two_steps_native(train_wide_vectors, tmp_path):
    """The mirrored training loop (reference training/sci_equilibrium_training.py:28-150) end to end on the native path:
    two implicit-differentiation steps of DE-GAP-FFDnet with the reference's optimizer (Adam) -- taken over by the fused
    exchange + Adam kernel -- and scheduler, then the epoch checkpoint with the reference's keys and a meaningful
    optimizer state."""
    from test_gpu_parity import build_solver
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.training import sci_equilibrium_training as tr
    dev = torch.device("cuda", 0)
    v = train_wide_vectors
    solver = build_solver("ffdnet", dev)
    solver.train()
    solver.nonlinear_op.train()
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=12, tol=1e-5)
    opt = torch.optim.Adam(solver.parameters(), lr=1e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    batch = {"gt": torch.from_numpy(v["gt"]), "meas": torch.from_numpy(v["y"]), "mask": torch.from_numpy(v["Phi"])}
    before = {k: p.detach().clone() for k, p in solver.named_parameters()}
    bn0 = solver.nonlinear_op.intermediate_dncnn.itermediate_dncnn[3].running_mean.clone()
    tr.train_solver_sci(single_iterate_solver=solver, train_dataloader=[batch, batch], test_dataloader=None, optimizer=opt,
                        save_model_path=str(tmp_path) + "/", deep_eq_module=deq, loss_function=torch.nn.MSELoss(reduction="mean"),
                        n_epochs=1, scheduler=sched, print_every_n_steps=100, device=dev)
    moved = [float((p.detach() - before[k]).abs().max()) for k, p in solver.named_parameters()]
    assert min(moved) > 0 and max(moved) < 1e-3                       # two Adam steps at lr 1e-4: every tensor moved a little
    assert all(torch.isfinite(p).all() for p in solver.parameters())
    assert not torch.equal(solver.nonlinear_op.intermediate_dncnn.itermediate_dncnn[3].running_mean, bn0)
    st = opt.state_dict()["state"]
    assert len(st) == len(list(solver.parameters())) and float(st[0]["step"]) == 2.0
    assert float(st[0]["exp_avg"].abs().max()) > 0 and float(st[0]["exp_avg_sq"].abs().max()) > 0
    ck = torch.load(os.path.join(str(tmp_path), "epoch_0.ckpt"), map_location="cpu", weights_only=False)
    assert set(ck) == {"solver_state_dict", "epoch", "optimizer_state_dict", "scheduler_state_dict"}
    assert ck["optimizer_state_dict"]["param_groups"][0]["lr"] == pytest.approx(5e-5)      # the scheduler stepped
    fresh = build_solver("ffdnet", dev)
    fresh.load_state_dict(ck["solver_state_dict"], strict=True)           # the checkpoint loads back, reference keys
