"""GPU (-m gpu): parity of the CUDA path, called through the C-ABI, against the numpy oracle on the
same seeded inputs and against the golden vectors produced by the reference's own code.

Bars (BASELINE.json north_star): operator <= 1e-6 relative; per-iterate relative L2 <= 1e-3;
final PSNR within 0.05 dB, SSIM within 1e-3."""
import os

import numpy as np
import pytest
import torch

from conftest import load_scene, load_weights, rel_l2
from oracle import deqsci_oracle as orc

pytestmark = pytest.mark.gpu

DENOISERS = ["ffdnet", "SimpleCNN", "RealSN_SimpleCNN"]
TAG = {"ffdnet": "ffdnet", "SimpleCNN": "denoiser", "RealSN_SimpleCNN": "denoiser"}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from deqsci_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libdeqsci.so missing: the CUDA path must be built"
    _lib.lib()
    return torch.device("cuda", 0)


GRAD_TESTS = {"test_eval_mode_with_grad_builds_the_graph"}


@pytest.fixture(autouse=True)
def _inference_mode(request):
    """These are inference parity tests: grad mode off (what selects the native kernels for direct module
    calls since eval mode alone no longer does) -- except the tests that check the graph path."""
    if request.node.originalname in GRAD_TESTS:
        yield
    else:
        with torch.no_grad():
            yield


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def build_solver(d, dev, precision=None):
    from deqsci_b200.networks.ffdnet.models import FFDNet
    from deqsci_b200.networks.provable.model.SimpleCNN_models import DnCNN
    from deqsci_b200.solvers.equilibrium_solvers_yaping import EquilibriumProxGradSCI
    from deqsci_b200.utils.cg_utils import A_torch_, At_torch_
    if d == "ffdnet":
        net = FFDNet(num_input_channels=1, tag="ffdnet")
    else:
        net = DnCNN(1, num_of_layers=4, lip=1.0 if d == "RealSN_SimpleCNN" else 0.0, no_bn=True, tag="denoiser")
    if precision:
        net.precision = precision
    net.eval()
    solver = EquilibriumProxGradSCI(A=A_torch_, At=At_torch_, nonlinear_operator=net, eta=0.2)
    sd = {k: torch.from_numpy(v) for k, v in load_weights(d).items()}
    if d == "RealSN_SimpleCNN":                        # the power-iteration probes weight_u are not in the fixture
        missing, unexpected = solver.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith("weight_u") for k in missing), (missing, unexpected)
    else:
        solver.load_state_dict(sd, strict=True)
    return solver.to(dev)


# ---------------------------------------------------------------------------------------------
# (1) operator / GAP kernels
# ---------------------------------------------------------------------------------------------
def test_operator_golden(dev, small_vectors):
    from deqsci_b200.utils.cg_utils import A_torch_, At_torch_, initial_point, Phi_sum_
    v = small_vectors
    x, Phi = t(v["op_x"], dev), t(v["op_Phi"], dev)
    y = A_torch_(x, Phi)
    assert rel_l2(y.cpu().numpy(), v["op_A"]) <= 1e-6
    assert np.abs(y.cpu().numpy() - v["op_A"]).max() <= 1e-6 * np.abs(v["op_A"]).max()
    np.testing.assert_array_equal(At_torch_(t(v["op_A"], dev), Phi).cpu().numpy(), v["op_At"])
    np.testing.assert_array_equal(initial_point(t(v["op_A"], dev), Phi, None, None).cpu().numpy(), v["op_x0"])
    np.testing.assert_allclose(Phi_sum_(Phi).cpu().numpy(), v["op_Phi_sum"], rtol=1e-6)
    np.testing.assert_array_equal(Phi_sum_(Phi[:1]).cpu().numpy(), v["op_Phi_sum"][:1])


@pytest.mark.parametrize("shape", [(1, 256, 256, 8), (3, 17, 23, 8), (2, 16, 16, 5), (1, 1, 1, 8), (2, 9, 7, 1)])
def test_gap_kernels_vs_oracle(dev, shape):
    from deqsci_b200 import ops
    rng = np.random.default_rng(sum(shape))
    z = rng.standard_normal(shape).astype(np.float32)
    Phi = (rng.random(shape) < 0.5).astype(np.float32)
    if shape[0] > 1:
        Phi[-1] = rng.random(shape[1:]).astype(np.float32)       # a grey mask
    Phi[0, 0, 0] = 0                                             # a pixel no frame sees
    y = rng.random(shape[:3]).astype(np.float32) * shape[3]
    Ps = orc.phi_sum(Phi)
    zt, Pt, yt, Pst = t(z, dev), t(Phi, dev), t(y, dev), t(Ps, dev)
    tol = 1e-6
    assert rel_l2(ops.gap_forward(zt, Pt).cpu().numpy(), orc.A(z, Phi)) <= tol
    np.testing.assert_array_equal(ops.gap_adjoint(yt, Pt).cpu().numpy(), orc.At(y, Phi))
    np.testing.assert_allclose(ops.phi_sum(Pt).cpu().numpy(), Ps, rtol=1e-6)
    assert rel_l2(ops.gap_step(zt, yt, Pt, Pst).cpu().numpy(), orc.gap_step(z, y, Phi, Ps)) <= tol
    assert rel_l2(ops.gap_vjp(zt, Pt, Pst).cpu().numpy(), orc.gap_vjp(z, Phi, Ps)) <= tol
    add = rng.standard_normal(shape).astype(np.float32)
    assert rel_l2(ops.gap_vjp(zt, Pt, Pst, add=t(add, dev)).cpu().numpy(), orc.gap_vjp(z, Phi, Ps) + add) <= tol


def test_gap_adjointness_and_projector(dev):
    """<A x, y> = <x, At y>; the GAP step is idempotent on the data (A(step(z)) = y where sum Phi > 0)
    and linear in (z, y); the VJP is a projector: vjp(vjp(v)) = vjp(v) for a binary mask."""
    from deqsci_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    shape = (2, 64, 48, 8)
    x = torch.rand(shape, generator=g).to(dev)
    Phi = (torch.rand(shape, generator=g) < 0.5).float().to(dev)
    y = torch.rand(shape[:3], generator=g).to(dev)
    lhs = (ops.gap_forward(x, Phi).double() * y.double()).sum()
    rhs = (x.double() * ops.gap_adjoint(y, Phi).double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * abs(float(rhs))
    Ps = ops.phi_sum(Phi)
    z1 = ops.gap_step(x, y * 4, Phi, Ps)
    seen = Phi.sum(3) > 0
    assert float((ops.gap_forward(z1, Phi) - y * 4)[seen].abs().max()) <= 2e-5
    v1 = ops.gap_vjp(x, Phi, Ps)
    assert rel_l2(ops.gap_vjp(v1, Phi, Ps).cpu().numpy(), v1.cpu().numpy()) <= 2e-6
    assert ops.gap_forward(torch.empty(0, 4, 4, 8, device=dev), torch.empty(0, 4, 4, 8, device=dev)).shape == (0, 4, 4)


# ---------------------------------------------------------------------------------------------
# (2) conv stack
# ---------------------------------------------------------------------------------------------
def _split_planes(x):
    """fp32 [NF,H,W,64] -> fp16 [2,NF,H,W,64] (hi, lo*2^11)."""
    hi = x.half()
    lo = ((x - hi.float()) * 2048.0).half()
    return torch.stack([hi, lo]).contiguous()


def _join_planes(p):
    return p[0].float() + p[1].float() / 2048.0


@pytest.mark.parametrize("NF,Hc,Wc", [(2, 16, 16), (3, 9, 24), (1, 5, 130), (8, 128, 128), (2, 64, 256),
                                      (1, 6, 128), (3, 10, 200), (5, 32, 65)])
@pytest.mark.parametrize("precision", ["tc_split", "fp32"])
def test_hidden_layer_vs_fp64_conv(dev, NF, Hc, Wc, precision):
    """One 64->64 layer (tcgen05 split-fp16 kernel and the fp32 CUDA-core kernel) against
    torch conv2d in fp64 on the same (hi+lo) inputs; affine + ReLU epilogue included."""
    from deqsci_b200.native import NativeDenoiser
    g = torch.Generator(device="cpu").manual_seed(NF * 1000 + Hc * 10 + Wc)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    scale = torch.rand(64, generator=g) * 2 + 0.1
    bias = torch.randn(64, generator=g) * 0.3
    layers = [{"weight": torch.randn(64, 1, 3, 3, generator=g), "relu": True},
              {"weight": w, "scale": scale, "bias": bias, "relu": True},
              {"weight": torch.randn(1, 64, 3, 3, generator=g), "relu": False}]
    plan = NativeDenoiser("dncnn", layers, precision=precision, device=dev)
    x = (torch.randn(NF, Hc, Wc, 64, generator=g) * 2).to(dev)
    planes = _split_planes(x)
    out = _join_planes(plan.debug_hidden_layer(1, planes, NF, Hc, Wc))
    xin = _join_planes(planes).double().permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(xin, w.double().to(dev), padding=1)
    ref = torch.relu(ref * scale.double().to(dev).view(1, -1, 1, 1) + bias.double().to(dev).view(1, -1, 1, 1))
    ref = ref.permute(0, 2, 3, 1)
    err = float((out.double() - ref).norm() / ref.norm())
    assert err <= 2e-6, err          # ~22-bit operands, fp32 accumulate
    assert float((out.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


@pytest.mark.parametrize("d", DENOISERS)
@pytest.mark.parametrize("precision", ["tc_split", "fp32"])
def test_denoiser_vs_oracle(dev, small_vectors, d, precision):
    """denoise_residual (whole stack) on a 64x64x8 crop, B=2, vs the numpy oracle network."""
    solver = build_solver(d, dev, precision)
    rng = np.random.default_rng(3)
    z = (small_vectors["crop_gt"] + 0.1 * rng.standard_normal(small_vectors["crop_gt"].shape)).astype(np.float32)
    B, H, W, T = z.shape
    sigma = np.float32(0.2)
    sd = {k[len("nonlinear_op."):]: v for k, v in load_weights(d).items()}
    frames = np.ascontiguousarray(z.transpose(0, 3, 1, 2)).reshape(B * T, 1, H, W)
    if d == "ffdnet":
        noise = orc.ffdnet_forward(frames, np.full(B * T, sigma, np.float32), sd)
    else:
        noise = orc.dncnn_forward(frames, sd)
    want = z - noise.reshape(B, T, H, W).transpose(0, 2, 3, 1)
    plan = solver.nonlinear_op.native_plan(dev)
    got = plan.denoise_residual(t(z, dev), float(sigma)).cpu().numpy()
    assert rel_l2(got, want) <= 2e-5
    # module-level API of the reference: net(x[N,1,H,W], sigma[N]) -> predicted noise
    xin = t(frames, dev)
    if d == "ffdnet":
        pred = solver.nonlinear_op(xin, torch.full((B * T,), float(sigma), device=dev))
    else:
        pred = solver.nonlinear_op(xin)
    assert rel_l2(pred.cpu().numpy(), noise) <= 1e-4


@pytest.mark.parametrize("d", DENOISERS)
@pytest.mark.parametrize("shape", [(1, 32, 160, 8), (2, 40, 260, 3)])
def test_wide_images_tc_vs_fp32_and_oracle(dev, d, shape):
    """Images wide enough for the 128-pixel row tiles (tensor-core first layer, CTA-pair hidden layers,
    rolling-row last layer), ragged in W: tc_split vs the fp32 CUDA-core stack vs the numpy oracle."""
    rng = np.random.default_rng(shape[2])
    z = rng.random(shape).astype(np.float32)
    Phi = (rng.random(shape) < 0.5).astype(np.float32)
    y = orc.A(rng.random(shape).astype(np.float32), Phi)
    Ps = orc.phi_sum(Phi)
    outs = {}
    for prec in ("tc_split", "fp32"):
        solver = build_solver(d, dev, prec)
        plan = solver.nonlinear_op.native_plan(dev)
        outs[prec] = plan.iterate(t(z, dev), t(y, dev), t(Phi, dev), t(Ps, dev), 0.1).cpu().numpy()
    assert rel_l2(outs["tc_split"], outs["fp32"]) <= 2e-5
    f = orc.ProxGradSCI(TAG[d], load_weights(d))
    f.noise_sigma = np.float32(0.1) / np.float32(0.971)      # next call multiplies by 0.971
    f.y = y.mean(dtype=np.float32)
    want = f(z, y, Phi, Ps)
    assert rel_l2(outs["tc_split"], want) <= 5e-5


@pytest.mark.parametrize("d", DENOISERS)
def test_f_two_calls_golden(dev, small_vectors, d):
    """EquilibriumProxGradSCI.forward twice (sigma decays once) vs the reference's own outputs."""
    v = small_vectors
    solver = build_solver(d, dev)
    Phi, y = t(v["crop_Phi"], dev), t(v["crop_y"], dev)
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    Ps, x0 = Phi_sum_(Phi), At_torch_(y, Phi)
    f1 = solver(x0, y, Phi, Ps)
    f2 = solver(f1, y, Phi, Ps)
    assert rel_l2(f1.cpu().numpy(), v["f1_" + d]) <= 2e-5
    assert rel_l2(f2.cpu().numpy(), v["f2_" + d]) <= 2e-5


# ---------------------------------------------------------------------------------------------
# (3) Anderson kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,m", [(1, 4096 * 3, 5), (3, 1000, 5), (2, 8 * 64 * 64, 3), (4, 37, 6)])
def test_anderson_kernels_vs_numpy(dev, B, N, m):
    """Ramp-up (n = 1..m) and wrap-around of the slot ring: gram rows, alpha, residual, mix."""
    from deqsci_b200.solvers.new_equilibrium_utils_yaping import _AndersonState
    rng = np.random.default_rng(B * 100 + m)
    x0 = torch.zeros(B, N, device=dev)
    st = _AndersonState(x0, m)
    Xh = np.zeros((m, B, N), np.float32)
    Fh = np.zeros((m, B, N), np.float32)
    lam, eps = 1e-2, 1e-5
    for step in range(2 * m + 1):
        slot, n = step % m, min(step + 1, m)
        Xh[slot] = rng.standard_normal((B, N)).astype(np.float32)
        Fh[slot] = Xh[slot] + 0.3 * rng.standard_normal((B, N)).astype(np.float32)
        st.X[slot].copy_(t(Xh[slot], dev))
        st.F[slot].copy_(t(Fh[slot], dev))
        st.update(slot, n, lam, eps)
        res = st.fetch_res(eps)
        G = (Fh[:n] - Xh[:n]).transpose(1, 0, 2)                       # [B,n,N]
        alpha = orc.anderson_alpha(G, lam)
        np.testing.assert_allclose(st.alpha.cpu().numpy()[:, :n], alpha, rtol=2e-4, atol=2e-6)
        gram = np.matmul(G.astype(np.float64), G.astype(np.float64).transpose(0, 2, 1))
        np.testing.assert_allclose(st.gram.cpu().numpy()[:, :n, :n], gram, rtol=2e-6, atol=1e-6)
        want_res = np.linalg.norm(G[:, slot].astype(np.float64)) / (eps + np.linalg.norm(Fh[slot].astype(np.float64)))
        assert abs(res - want_res) <= 2e-6 * want_res
        for beta in (1.0, 0.7):
            nxt = (step + 1) % m
            keep = st.X[nxt].clone()
            st.mix(nxt, n, beta)
            a64 = st.alpha.cpu().numpy()[:, :n].astype(np.float64)
            want = beta * np.einsum("bj,jbn->bn", a64, Fh[:n].astype(np.float64)) + \
                (1 - beta) * np.einsum("bj,jbn->bn", a64, Xh[:n].astype(np.float64))
            assert rel_l2(st.X[nxt].cpu().numpy(), want) <= 2e-6
            st.X[nxt].copy_(keep)


def test_anderson_nchw_entry_point(dev):
    """`anderson` (reference :114-150): NCHW input, returns the list of residuals; same update as
    andersonexp.  Checked on a contractive affine map against the numpy oracle."""
    from deqsci_b200.solvers.new_equilibrium_utils_yaping import anderson
    rng = np.random.default_rng(11)
    x0 = rng.standard_normal((3, 2, 16, 24)).astype(np.float32)
    bvec = rng.standard_normal(x0.shape).astype(np.float32)
    fm_t = lambda q: 0.6 * torch.roll(q, 1, dims=3) + t(bvec, dev)
    z, res = anderson(fm_t, t(x0, dev), m=4, lam=1e-4, max_iter=14, tol=1e-9, beta=0.9)
    fm_n = lambda q: (np.float32(0.6) * np.roll(q, 1, axis=3) + bvec).astype(np.float32)
    trace = []
    zo, reso = orc.andersonexp(fm_n, x0, m=4, lam=1e-4, max_iter=14, tol=1e-9, beta=0.9, trace=trace)
    assert isinstance(res, list) and len(res) == 12
    assert rel_l2(z.cpu().numpy(), zo) <= 1e-3
    assert abs(res[-1] - reso) <= 0.05 * reso + 1e-7


def test_early_stop_lagged_equals_synchronous(dev):
    """tol reached mid-run: the lagged residual check (one speculative iteration, then rollback) returns
    the same iterate, residual and effective call count as the per-iteration check and as the oracle."""
    from deqsci_b200.solvers.new_equilibrium_utils_yaping import andersonexp
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal((2, 8, 8, 8)).astype(np.float32)
    bvec = rng.standard_normal(x0.shape).astype(np.float32)
    bt = t(bvec, dev)

    class Map:
        supports_rollback = True

        def __init__(self):
            self.calls = 0

        def __call__(self, q):
            self.calls += 1
            return 0.5 * torch.flip(q, dims=[2]) + bt

        def rollback(self):
            self.calls -= 1

    lag = Map()
    z1, r1 = andersonexp(lag, t(x0, dev), m=5, lam=1e-6, max_iter=40, tol=1e-4, beta=1.0)
    calls = [0]

    def plain(q):
        calls[0] += 1
        return 0.5 * torch.flip(q, dims=[2]) + bt
    z2, r2 = andersonexp(plain, t(x0, dev), m=5, lam=1e-6, max_iter=40, tol=1e-4, beta=1.0)
    assert lag.calls == calls[0] < 40 and r1 == r2 and r1 < 1e-4
    assert torch.equal(z1, z2)
    n_o = [0]
    fo = lambda q: (n_o.__setitem__(0, n_o[0] + 1), (np.float32(0.5) * np.flip(q, axis=2) + bvec).astype(np.float32))[1]
    zo, ro = orc.andersonexp(fo, x0, m=5, lam=1e-6, max_iter=40, tol=1e-4, beta=1.0)
    assert n_o[0] == calls[0]
    assert rel_l2(z1.cpu().numpy(), zo) <= 1e-4


def test_residual_kernel(dev):
    from deqsci_b200.solvers.new_equilibrium_utils_yaping import forward_iteration
    a = torch.randn(3, 16, 16, 8, device=dev)
    z, res = forward_iteration(lambda x: 0.5 * x + 1.0, a, max_iter=4, tol=0.0)
    x = a * 0.5 + 1.0
    want = []
    for _ in range(4):
        f0 = 0.5 * x + 1.0
        want.append(float((f0 - x).double().norm()) / (1e-7 + float(f0.double().norm())))
        x = f0
    np.testing.assert_allclose(res, want, rtol=1e-5)
    assert rel_l2(z.cpu().numpy(), x.cpu().numpy()) <= 1e-6


# ---------------------------------------------------------------------------------------------
# (4) solvers end to end on the golden crops (B = 2, 30 iterations, all three denoisers)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", DENOISERS)
def test_deq_andersonexp_golden(dev, small_vectors, d):
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    v = small_vectors
    solver = build_solver(d, dev)
    Phi, y = t(v["crop_Phi"], dev), t(v["crop_y"], dev)
    Ps, x0 = Phi_sum_(Phi), At_torch_(y, Phi)
    seen = []
    h = solver.register_forward_pre_hook(lambda mod, args: seen.append(args[0].detach().clone()))
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=30, tol=1e-5)
    z = deq.forward(y, Phi, Ps, initial_point=x0, train_flag=False)
    h.remove()
    assert len(seen) == 31                      # 30 solver calls + the reconstruction call
    norms = np.array([float(s.double().norm()) for s in seen])
    np.testing.assert_allclose(norms, v["deq30_innorm_" + d][:31], rtol=1e-4)
    assert rel_l2(seen[10].cpu().numpy(), v["deq30_in10_" + d]) <= 1e-3      # per-iterate bar
    assert rel_l2(z.cpu().numpy(), v["deq30_z_" + d]) <= 1e-3
    assert abs(deq.forward_res - float(v["deq30_res_" + d])) <= 1e-2 * float(v["deq30_res_" + d])
    # sigma schedule advanced exactly like the reference's 32 calls
    if d == "ffdnet":
        s = np.float32(60 / 255)
        for _ in range(31):
            s = np.float32(s * np.float32(0.971))
        assert float(solver._sigma) == float(s)


@pytest.mark.parametrize("d", DENOISERS)
def test_forward_iteration_golden(dev, small_vectors, d):
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    v = small_vectors
    solver = build_solver(d, dev)
    Phi, y = t(v["crop_Phi"][:1], dev), t(v["crop_y"][:1], dev)
    Ps, x0 = Phi_sum_(Phi), At_torch_(y, Phi)
    z, res = eq.forward_iteration(lambda q: solver(q, y, Phi, Ps), x0, max_iter=6, tol=1e-5)
    assert rel_l2(z.cpu().numpy(), v["picard6_z_" + d]) <= 1e-4
    np.testing.assert_allclose(res, v["picard6_res_" + d], rtol=1e-3)


@pytest.mark.parametrize("d", ["ffdnet", "SimpleCNN"])
def test_driver_equals_python_loop(dev, small_vectors, d, monkeypatch):
    """deqsci_reconstruct (the whole forward() in one C-ABI call) vs the Python-driven solver loop:
    same kernels in the same order => bit-identical reconstruction, residual and sigma-schedule state,
    also across two consecutive measurements (schedule reset / continuation)."""
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    v = small_vectors
    Phi, y = t(v["crop_Phi"], dev), t(v["crop_y"], dev)
    Ps, x0 = Phi_sum_(Phi), At_torch_(y, Phi)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DEQSCI_DRIVER", mode)
        solver = build_solver(d, dev)
        deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=20, tol=1e-5)
        z1 = deq.forward(y, Phi, Ps, initial_point=x0, train_flag=False)
        r1 = deq.forward_res
        z2 = deq.forward(y, Phi, Ps, initial_point=x0, train_flag=False)      # same tensor: schedule continues
        y3 = (y * 0.5).contiguous()
        z3 = deq.forward(y3, Phi, Ps, initial_point=At_torch_(y3, Phi), train_flag=False)   # new mean: reset
        outs[mode] = (z1, r1, z2, z3, deq.forward_res, solver._n)
    a, b = outs["1"], outs["0"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert a[1] == b[1] and a[4] == b[4] and a[5] == b[5]
    if d == "ffdnet":
        assert not torch.equal(a[0], a[2])       # the continued schedule really changes the result



@pytest.mark.parametrize("d,shape", [("ffdnet", (1, 256, 256)), ("ffdnet", (3, 64, 192)), ("SimpleCNN", (2, 48, 128))])
def test_pair_kernel_issue_modes_agree(dev, d, shape, tmp_path):
    """The pair kernel's issue orders (csrc/conv_tc2.cu; fixed per process, hence subprocesses).  The two row-stationary
    modes share one accumulation order, so the narrow (DEQSCI_TC_RS=1) and wide (2, the default) instruction shapes
    must agree BIT FOR BIT -- on single iterate-map calls (fresh plan, reused plan, other data in the workspace) and on
    a 10-iteration solve; the output-stationary order (DEQSCI_TC_RS=0) adds the two correction products in another order
    and agrees to fp32 rounding.  Also: the same call twelve times over gives the same bits."""
    import subprocess
    import sys
    from conftest import ROOT
    B, H, W = shape
    outs = {}
    for name, env in {"wide": {"DEQSCI_TC_RS": "2"}, "narrow": {"DEQSCI_TC_RS": "1"}, "os": {"DEQSCI_TC_RS": "0"},
                      "wide_nolook": {"DEQSCI_TC_RS": "2", "DEQSCI_TC_LOOKAHEAD": "0"}}.items():
        out = str(tmp_path / (name + ".npz"))
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "run_denoiser_once.py"), d, out, str(B), str(H),
                            str(W)], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = np.load(out)
    keys = ("f1", "a1", "a2", "b1", "z")
    report = {n: dict({k: float(np.abs(outs[n][k] - outs["wide"][k]).max()) for k in keys},
                      rep_maxdiff=float(outs[n]["rep_maxdiff"])) for n in outs}
    print(report)
    for n in outs:
        assert float(outs[n]["rep_maxdiff"]) == 0.0, (n, report)     # the same call again: the same bits
    for k in keys:
        assert np.array_equal(outs["narrow"][k], outs["wide"][k]), (k, report)
        assert np.array_equal(outs["wide_nolook"][k], outs["wide"][k]), (k, report)
    assert float(outs["narrow"]["res"]) == float(outs["wide"]["res"])
    assert max(rel_l2(outs["os"][k], outs["wide"][k]) for k in ("f1", "a1", "a2", "b1")) <= 1e-6
    assert rel_l2(outs["os"]["z"], outs["wide"]["z"]) <= 1e-4


def test_per_iterate_trace_vs_oracle(dev, small_vectors):
    """Every iterate of a 12-iteration Anderson run vs the numpy oracle (SimpleCNN weights)."""
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    v = small_vectors
    d = "SimpleCNN"
    Phi_n, y_n = v["crop_Phi"][:1], v["crop_y"][:1]
    f_o = orc.ProxGradSCI(TAG[d], load_weights(d))
    want = []
    orc.andersonexp(lambda z: f_o(z, y_n, Phi_n, orc.phi_sum(Phi_n)), orc.At(y_n, Phi_n), m=5, lam=1e-2,
                    max_iter=12, tol=1e-5, beta=1.0, trace=want)
    solver = build_solver(d, dev)
    Phi, y = t(Phi_n, dev), t(y_n, dev)
    Ps = Phi_sum_(Phi)
    got = []
    h = solver.register_forward_pre_hook(lambda mod, args: got.append(args[0].detach().clone()))
    eq.andersonexp(lambda q, out=None: solver(q, y, Phi, Ps), At_torch_(y, Phi), m=5, lam=1e-2, max_iter=12,
                   tol=1e-5, beta=1.0)
    h.remove()
    got = got[2:]                                 # first two calls are x0 and f(x0)
    assert len(got) == len(want) == 10
    for k, (a, b) in enumerate(zip(got, want)):
        assert rel_l2(a.cpu().numpy(), b) <= 1e-3, k


# ---------------------------------------------------------------------------------------------
# (5) full-size reconstructions: PSNR / SSIM against the reference's numbers
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,scene", [("SimpleCNN", "drop8"), ("RealSN_SimpleCNN", "runner8"), ("ffdnet", "drop8")])
def test_full_scene_psnr_vs_reference(dev, full_recon, d, scene):
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    key = "%s_%s_0" % (d, scene)
    if key + "_psnr" not in full_recon:
        pytest.skip("no reference reconstruction recorded for " + key)
    gt, mask, meas = load_scene(scene)
    solver = build_solver(d, dev)
    Phi, y = t(mask[None], dev), t(meas[None, :, :, 0], dev)
    Ps = Phi_sum_(Phi)
    max_iter = 180 if d == "ffdnet" else 100
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=max_iter, tol=1e-5)
    seen = []
    hk = solver.register_forward_pre_hook(lambda mod, args: seen.append(args[0].detach()[0, 96:160, 96:160].clone()))
    z = deq.forward(y, Phi, Ps, initial_point=At_torch_(y, Phi), train_flag=False)
    hk.remove()
    # per-iterate parity at full size: inputs of calls 2, 20 and the last solver call vs the reference's
    # (64x64x8 crops of the 256x256x8 iterates), and the norm of every iterate
    ncalls = int(full_recon[key + "_ncalls"])
    assert len(seen) == ncalls - 1                     # the reference's extra call is skipped at inference
    for k in (2, 20, ncalls - 2):
        assert rel_l2(seen[k].cpu().numpy(), full_recon[key + "_in%d_crop" % k]) <= 1e-3, k
    rec = z.clip(0, 1).cpu().numpy()
    g = gt[None, :, :, 0:8]
    psnr = orc.psnr(g, rec)
    ssim = orc.ssim(rec.transpose(0, 3, 1, 2), g.transpose(0, 3, 1, 2))
    assert abs(psnr - float(full_recon[key + "_psnr"])) <= 0.05, (psnr, float(full_recon[key + "_psnr"]))
    assert abs(ssim - float(full_recon[key + "_ssim"])) <= 1e-3
    crop = z[0, 96:160, 96:160].cpu().numpy()
    assert rel_l2(crop, full_recon[key + "_z_crop"]) <= 1e-3


def test_benchmark_workload_vs_reference(dev):
    """VERDICT r01 missing #4 / SURVEY 8(d) 'parity subset': the workload bench.py times -- measurements 0 and 1 of
    bench.synthetic_batch, DE-GAP-FFDnet, 180 iterations -- against the reference's own run of the same two
    measurements (tests/golden/synthetic_recon.npz, make_golden.py --stage synthetic).  Bars: north-star
    1e-3 per iterate / 0.05 dB / 1e-3 SSIM; the iterates after call ~60 are compared at 5e-3 because on this
    data two fp32 runs of the SAME code already differ by 3.7e-4 .. 1.6e-3 there (tests/tools/emulate_gpu.py,
    'jitter' rows of profiles/r02_precision_emulation.md)."""
    import bench
    from conftest import GOLDEN
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    g = dict(np.load(os.path.join(GOLDEN, "synthetic_recon.npz")))
    assert str(g["kind"]) == bench.DATA_KIND and int(g["seed"]) == bench.SEED, "golden made for another generator"
    ys, ps, xs = bench.synthetic_batch(0, 2)
    assert abs(float(ys[0].mean()) - float(g["m0_ymean"])) < 1e-6        # same synthetic data as the reference saw
    solver = build_solver("ffdnet", dev)
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=180, tol=1e-5)
    y, Phi = ys.to(dev), ps.to(dev)
    seen = []
    hk = solver.register_forward_pre_hook(lambda mod, args: seen.append(args[0].detach().clone()))
    z = deq.forward(y, Phi, Phi_sum_(Phi), initial_point=At_torch_(y, Phi), train_flag=False)
    hk.remove()
    assert len(seen) == 181
    report = {}
    for i in range(2):
        key = "m%d" % i
        norms = np.array([float(s_[i].double().norm()) for s_ in seen])
        dev_n = np.abs(norms - g[key + "_innorm"][:181]) / g[key + "_innorm"][:181]
        assert dev_n[:41].max() <= 1e-4 and dev_n.max() <= 2e-3, (i, dev_n[:41].max(), dev_n.max())
        rel = {}
        for k in (2, 20, 40, 100, 180):
            rel[k] = rel_l2(seen[k][i, 96:160, 96:160].cpu().numpy(), g[key + "_in%d_crop" % k])
            assert rel[k] <= (1e-3 if k <= 40 else 5e-3), (i, k, rel[k])
        rel["z"] = rel_l2(z[i, 96:160, 96:160].cpu().numpy(), g[key + "_z_crop"])
        assert rel["z"] <= 5e-3
        rec = z[i:i + 1].clip(0, 1).cpu().numpy()
        gt = xs[i:i + 1].numpy()
        psnr = orc.psnr(gt, rec)
        ssim = orc.ssim(rec.transpose(0, 3, 1, 2), gt.transpose(0, 3, 1, 2))
        assert abs(psnr - float(g[key + "_psnr"])) <= 0.05, (psnr, float(g[key + "_psnr"]))
        assert abs(ssim - float(g[key + "_ssim"])) <= 1e-3
        report[key] = {"rel_l2_crops": {str(k): float(v) for k, v in rel.items()}, "dpsnr": psnr - float(g[key + "_psnr"]),
                       "dssim": ssim - float(g[key + "_ssim"]), "max_norm_dev_first41": float(dev_n[:41].max()),
                       "max_norm_dev": float(dev_n.max())}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        json.dump(report, open(os.path.join(out_dir, "parity_report_synthetic.json"), "w"), indent=1)


def test_sigma_schedule_resets_for_function_scoped_measurements(dev, full_recon):
    """drop8 then runner8, each measurement a function-scoped tensor (the caching allocator
    hands the second one the address the first just freed).  The schedule must reset by VALUE (reference
    solvers/equilibrium_solvers_yaping.py:409-413) -- both reconstructions match the reference's; with the
    old address-keyed cache the second one started at sigma_182 ~ 1e-3 instead of 60/255."""
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    solver = build_solver("ffdnet", dev)
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=180, tol=1e-5)
    ptrs = []

    def one(scene):
        gt, mask, meas = load_scene(scene)
        Phi = t(mask[None], dev)
        y = t(meas[None, :, :, 0], dev)                      # dies at return
        ptrs.append(y.data_ptr())
        z = deq.forward(y, Phi, Phi_sum_(Phi), initial_point=At_torch_(y, Phi), train_flag=False)
        return orc.psnr(gt[None, :, :, 0:8], z.clip(0, 1).cpu().numpy()), z[0, 96:160, 96:160].cpu().numpy()

    for scene in ("drop8", "runner8"):
        key = "ffdnet_%s_0" % scene
        psnr, crop = one(scene)
        assert solver._n == 182                              # 180 solver calls + 2 post-solver calls since the reset
        assert abs(psnr - float(full_recon[key + "_psnr"])) <= 0.05, (scene, psnr, float(full_recon[key + "_psnr"]))
        if scene == "drop8":
            assert rel_l2(crop, full_recon[key + "_z_crop"]) <= 1e-3
    assert ptrs[0] == ptrs[1], "the allocator did not reuse the address: the regression scenario was not exercised"


def test_eval_mode_with_grad_builds_the_graph(dev, small_vectors):
    """eval mode + grad enabled + trainable parameters is a differentiable call (reference
    solvers/new_equilibrium_utils_yaping.py:268-280 attaches the graph and the hook regardless of mode);
    train_flag=False or no_grad is inference and gives the same reconstruction."""
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    v = small_vectors
    Phi, y = t(v["crop_Phi"], dev), t(v["crop_y"], dev)
    Ps, x0 = Phi_sum_(Phi), At_torch_(y, Phi)
    solver = build_solver("SimpleCNN", dev)
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=12, tol=1e-5)
    z_inf = deq.forward(y, Phi, Ps, initial_point=x0, train_flag=False)
    assert not z_inf.requires_grad
    z = deq.forward(y, Phi, Ps, initial_point=x0)            # eval mode, grad on, default train_flag
    assert z.requires_grad
    assert rel_l2(z.detach().cpu().numpy(), z_inf.cpu().numpy()) <= 1e-4
    z.square().mean().backward()
    g = [p.grad for p in solver.parameters()]
    assert all(gi is not None and torch.isfinite(gi).all() for gi in g) and any(float(gi.abs().max()) > 0 for gi in g)
    assert deq.backward_res is not None


def test_admm_sci_golden(dev, small_vectors):
    """EquilibriumADMMSCI + admmexp + DEQFixedPointADMM on the native GAP step (z+u, Phi_sum + 1e-8) and
    the native denoiser in 'replace' mode, against the reference's own run (tests/golden/admm_vectors.npz).  As in
    the reference the one-argument frame denoiser needs `conv3d = False` set by hand (its DnCNN lacks the attribute)."""
    from conftest import GOLDEN
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.solvers.equilibrium_solvers_yaping import EquilibriumADMMSCI
    from deqsci_b200.utils.cg_utils import A_torch_, At_torch_, Phi_sum_
    g = dict(np.load(os.path.join(GOLDEN, "admm_vectors.npz")))
    v = small_vectors
    Phi, y = t(v["crop_Phi"], dev), t(v["crop_y"], dev)
    Ps, x0 = Phi_sum_(Phi), At_torch_(y, Phi)
    net = build_solver("SimpleCNN", dev).nonlinear_op
    f = EquilibriumADMMSCI(A_torch_, At_torch_, net, eta=0.2)
    with pytest.raises(AttributeError):                       # same failure as the reference without the attribute
        f(x0, torch.zeros_like(x0), y, Phi, Ps)
    net.conv3d = False
    z1, u1 = f(x0, torch.zeros_like(x0), y, Phi, Ps)
    z2, u2 = f(z1, u1, y, Phi, Ps)
    for got, key in ((z1, "z1"), (u1, "u1"), (z2, "z2"), (u2, "u2")):
        assert rel_l2(got.cpu().numpy(), g[key]) <= 1e-4, key
    deq = eq.DEQFixedPointADMM(f, eq.admmexp, eq.admmexp, m=5, beta=1.0, lam=1e-2, max_iter=6, tol=1e-5)
    z = deq.forward(y, Phi, Ps, initial_point=[x0, torch.zeros_like(x0)], train_flag=False)
    assert rel_l2(z.cpu().numpy(), g["deq_z"]) <= 1e-3
    assert abs(deq.forward_res - float(g["deq_res"])) <= 1e-3 * float(g["deq_res"])


# ---------------------------------------------------------------------------------------------
# (6) the caller: test_solver_sci over all benchmark scenes present (configs 2 and 3)
# ---------------------------------------------------------------------------------------------
# DE-GAP-FFDnet with the stand-in weights (net_gray.pth; ffdnet.ckpt is a missing blob) is NOT contractive on
# the traffic scene: the Anderson iteration stalls at a residual of 3e-3 and amplifies fp32 rounding
# differences from call ~40 on, so two fp32 implementations of the SAME algorithm no longer agree --
# the numpy oracle vs the reference's PyTorch run differ by up to 1.9e-3 in iterate norm and 0.03 dB,
# the fp32 CUDA-core path vs the tcgen05 path by 0.1-0.3 dB (tests/golden/make_noise_floor.py,
# scripts/diag_traffic.py).  For that one (denoiser, scene) pair the PSNR bar is the measured spread;
# everywhere else it is the north-star 0.05 dB / 1e-3 SSIM.
def _psnr_tol(d, scene):
    return (0.35, 2e-2) if (d == "ffdnet" and scene == "traffic") else (0.05, 1e-3)


@pytest.mark.parametrize("d", DENOISERS)
def test_solver_sci_all_scenes_vs_reference(dev, full_recon, d):
    """All 8 benchmark measurements (drop8, runner8, traffic x6) through the mirrored
    test_solver_sci (ONE batched solve over all scenes, device-side PSNR / SSIM): per-measurement PSNR / SSIM
    against the reference's own run, and the reported average against the reference's."""
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.training import sci_equilibrium_training as tr
    from deqsci_b200.utils.metrics import peak_signal_noise_ratio, ssim
    solver = build_solver(d, dev)
    max_iter = 180 if d == "ffdnet" else 100
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=max_iter, tol=1e-5)
    samples, want_scene, n_meas = [], [], {"drop8": 1, "runner8": 1, "traffic": 6}
    for scene in n_meas:
        gt, mask, meas = load_scene(scene)
        if scene in ("drop8", "runner8"):
            meas = np.concatenate([meas] + [meas[:, :, :1]] * 4, axis=2)     # the .mat files carry 5 measurements
        samples.append({"gt": torch.from_numpy(gt)[None], "mask": torch.from_numpy(mask)[None],
                        "meas": torch.from_numpy(meas)[None], "file": [scene + "_cacti.mat"]})
        want_scene.append(np.mean([float(full_recon["%s_%s_%d_psnr" % (d, scene, i)]) for i in range(n_meas[scene])]))
    avg, images = tr.test_solver_sci(deq, samples, save_img_path=None, verbose=False, save_image=False, device=dev)
    assert abs(avg - float(np.mean(want_scene))) <= (0.1 if d == "ffdnet" else 0.05)
    assert len(images) == 8 * 8                                   # one [H,W,1] array per frame
    assert deq.forward_min_sample_res is not None and deq.forward_min_sample_res <= deq.forward_res * 1.0001
    dev_metrics = tr.test_solver_sci.last_metrics                 # reduced on the device
    k0 = "drop8_cacti.mat_reconstruction_0.png"
    assert images[k0].shape == (256, 256, 1) and images[k0].max() <= 255.0
    report = []
    for scene, n in n_meas.items():
        gt, _, _ = load_scene(scene)
        tol_p, tol_s = _psnr_tol(d, scene)
        for fi in range(n):
            rec = np.stack([images["%s_cacti.mat_reconstruction_%d.png" % (scene, fi * 8 + t)][:, :, 0]
                            for t in range(8)], axis=2) / 255.0
            g = gt[:, :, fi * 8:(fi + 1) * 8]
            dp = peak_signal_noise_ratio(g, rec) - float(full_recon["%s_%s_%d_psnr" % (d, scene, fi)])
            ds = float(ssim(torch.from_numpy(rec.astype(np.float32)).permute(2, 0, 1)[None],
                            torch.from_numpy(g).permute(2, 0, 1)[None])) - float(full_recon["%s_%s_%d_ssim" % (d, scene, fi)])
            report.append((scene, fi, round(dp, 4), round(ds, 5)))
            m_dev = dev_metrics[scene + "_cacti.mat"]
            assert abs(m_dev["psnr"][fi] - peak_signal_noise_ratio(g, rec)) <= 2e-3        # device vs host metric
            assert abs(m_dev["ssim"][fi] - (ds + float(full_recon["%s_%s_%d_ssim" % (d, scene, fi)]))) <= 2e-4
            assert abs(dp) <= tol_p and abs(ds) <= tol_s, report
    print("dPSNR/dSSIM vs reference:", report)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):                                    # evidence for profiles/ (scratch dir on the GPU box)
        import json
        with open(os.path.join(out_dir, "parity_report_%s.json" % d), "w") as fh:
            json.dump([{"scene": a, "meas": b, "dpsnr": c, "dssim": e} for a, b, c, e in report], fh)


def test_solver_sci_stops_per_measurement_like_the_reference(dev, small_vectors):
    """the reference solves and STOPS each measurement on its own (batch 1, whole-batch residual =
    that sample's).  With a tolerance that fires, the batched test_solver_sci must fall back to per-measurement
    solves and return exactly what the reference's loop structure returns; metrics come back per measurement."""
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.training import sci_equilibrium_training as tr
    from deqsci_b200.utils.cg_utils import At_torch_, Phi_sum_
    v = small_vectors
    d = "SimpleCNN"
    tol = 2e-2
    sample = {"gt": torch.from_numpy(np.concatenate([v["crop_gt"][0], v["crop_gt"][1]], axis=2))[None],
              "mask": torch.from_numpy(v["crop_Phi"][:1]), "meas": torch.from_numpy(v["crop_y"].transpose(1, 2, 0))[None],
              "file": ["traffic_crop.mat"]}
    solver = build_solver(d, dev)
    deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=40, tol=tol)
    avg, images = tr.test_solver_sci(deq, [sample], save_img_path=None, verbose=False, save_image=False, device=dev)
    assert set(tr.test_solver_sci.last_metrics["traffic_crop.mat"]) == {"psnr", "ssim"}
    assert len(tr.test_solver_sci.last_metrics["traffic_crop.mat"]["psnr"]) == 2
    # the reference's structure: one solve per measurement
    Phi = t(v["crop_Phi"][:1], dev)
    Ps = Phi_sum_(Phi)
    iters = []
    for fi in range(2):
        y = t(v["crop_y"][fi:fi + 1], dev)
        z = deq.forward(y, Phi, Ps, initial_point=At_torch_(y, Phi), train_flag=False)
        iters.append(deq.forward_res)
        want = z.clip(0, 1).cpu().numpy()[0]
        got = np.stack([images["traffic_crop.mat_reconstruction_%d.png" % (fi * 8 + k)][:, :, 0] for k in range(8)], 2) / 255.0
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    assert min(iters) < tol                                      # the stopping test really fired


# ---------------------------------------------------------------------------------------------
# (7) the BatchNorm DnCNN behind `--denoiser DnCNN` (reference networks/provable/model/models.py)
# ---------------------------------------------------------------------------------------------
def test_dncnn_batchnorm_variant(dev):
    from conftest import GOLDEN
    from deqsci_b200.networks.provable.model.models import DnCNN
    v = dict(np.load(os.path.join(GOLDEN, "dncnn_bn_vectors.npz")))
    net = DnCNN(channels=1, num_of_layers=5).eval()
    net.load_state_dict({k[len("sd::"):]: torch.from_numpy(v[k]) for k in v if k.startswith("sd::")}, strict=True)
    net = net.to(dev)
    got = net(t(v["x"], dev)).cpu().numpy()                     # narrow image: per-tap tensor-core path
    assert rel_l2(got, v["y"]) <= 2e-5
    # 17 layers on a wide image (CTA-pair hidden layers): vs the numpy oracle on random weights
    torch.manual_seed(3)
    big = DnCNN(channels=1, num_of_layers=17).eval()
    for m in big.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.8, 1.3)
    sd = {k: p.detach().numpy() for k, p in big.state_dict().items()}
    x = np.random.default_rng(1).random((2, 1, 24, 136)).astype(np.float32)
    want = orc.dncnn_forward(x, sd, num_of_layers=17)
    got = big.to(dev)(t(x, dev)).cpu().numpy()
    assert rel_l2(got, want) <= 5e-5


def test_realsn_dncnn_eval_native(dev):
    """`--denoiser RealSN_DnCNN`: eval mode uses the stored `weight` buffers (reference
    Spectral_Normalize_chen.py:87-89) with BatchNorm folded, on the native conv stack; against the reference's
    own eval output after its train-mode step (so the buffers are the normalised weights)."""
    from test_host_cpu import _realsn_fixture
    net, v = _realsn_fixture()
    sd1 = {k[len("sd1::"):]: torch.from_numpy(v[k]) for k in v if k.startswith("sd1::") and not k.endswith("weight_u")}
    net.load_state_dict(sd1, strict=False)       # probes are not needed in eval mode
    net = net.eval().to(dev)
    got = net(t(v["x"], dev)).cpu().numpy()
    assert rel_l2(got, v["y_eval"]) <= 2e-5


# ---------------------------------------------------------------------------------------------
# (7b) full-size properties (256x256x8, the north-star shape): what must hold at any size
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", ["ffdnet", "SimpleCNN"])
def test_full_size_properties(dev, d):
    """At BASELINE.json's full size the oracle is too slow to rerun in a test, so the CUDA path is checked
    through size-independent properties, all bit-exact:
      * batch independence (= sharding invariance): measurement i reconstructed inside a batch of
        three equals measurement i reconstructed alone (per-sample Anderson weights, no cross-sample term);
      * frame equivariance of the denoiser: permuting the T frames of a cube permutes its output;
      * data consistency of the GAP step: A(z') == y to fp32 rounding, and z' is a fixed point of the step."""
    from deqsci_b200.solvers import new_equilibrium_utils_yaping as eq
    from deqsci_b200.utils.cg_utils import A_torch_, At_torch_, Phi_sum_
    from deqsci_b200 import ops
    data = orc.synthetic_measurements(64, 3)                        # 256 x 256 x 8
    y, Phi = t(data["y"], dev), t(data["Phi"], dev)
    Ps = Phi_sum_(Phi)
    solver = build_solver(d, dev)

    def recon(sl):
        deq = eq.DEQFixedPoint(solver, eq.andersonexp, m=5, beta=1.0, lam=1e-2, max_iter=10, tol=1e-5)
        return deq.forward(y[sl], Phi[sl], Ps[sl], initial_point=At_torch_(y[sl], Phi[sl]), train_flag=False)

    whole = recon(slice(0, 3))
    assert torch.isfinite(whole).all()
    for i in range(3):
        assert torch.equal(recon(slice(i, i + 1))[0], whole[i]), "measurement %d depends on its batch" % i

    plan = solver.nonlinear_op.native_plan(dev)
    z = At_torch_(y, Phi)
    perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4], device=dev)
    base = plan.denoise_residual(z, 0.2)
    assert torch.equal(plan.denoise_residual(z[..., perm].contiguous(), 0.2), base[..., perm])

    zp = ops.gap_step(z, y, Phi, Ps)
    assert rel_l2(A_torch_(zp, Phi).cpu().numpy(), data["y"]) <= 1e-6
    assert rel_l2(ops.gap_step(zp, y, Phi, Ps).cpu().numpy(), zp.cpu().numpy()) <= 1e-6


# ---------------------------------------------------------------------------------------------
# (8) the boundary really is a C-ABI: a plain-C host program against include/deqsci.h
# ---------------------------------------------------------------------------------------------
def test_c_abi_from_plain_c(dev, tmp_path):
    import re
    import shutil
    import subprocess
    from conftest import ROOT
    from deqsci_b200 import _lib
    from deqsci_b200.native import NativeDenoiser
    from deqsci_b200 import ops
    cc = shutil.which("gcc") or shutil.which("cc")
    cuda = "/usr/local/cuda"
    if cc is None or not os.path.isdir(cuda):
        pytest.skip("no C compiler / CUDA toolkit on this box")
    exe = str(tmp_path / "smoke")
    subprocess.check_call([cc, "-std=c99", "-O1", os.path.join(ROOT, "tests", "c_abi", "smoke.c"), "-o", exe,
                           "-I", os.path.join(ROOT, "include"), "-I", cuda + "/include",
                           "-L", os.path.dirname(_lib.LIB_PATH), "-ldeqsci", "-L", cuda + "/lib64", "-lcudart",
                           "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-Wl,-rpath," + cuda + "/lib64", "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"residual=(\S+) iterations=(\d+) f_calls=(\d+) sum=(\S+) sumsq=(\S+)", out.stdout)
    assert m, out.stdout
    # the same problem through the Python host mirror (same LCG stream as smoke.c)
    state = [12345]

    def lcg():
        state[0] = (state[0] * 1664525 + 1013904223) & 0xFFFFFFFF
        return np.float32(np.float32(state[0] >> 8) / np.float32(16777216.0) - np.float32(0.5))
    B, H, W, T = 2, 16, 144, 8
    cin, cout = [1, 64, 64, 64], [64, 64, 64, 1]
    layers = []
    for l in range(4):
        scale = np.float32(0.5 if l == 0 else 0.08)
        w = np.array([lcg() * scale for _ in range(cout[l] * cin[l] * 9)], np.float32).reshape(cout[l], cin[l], 3, 3)
        layers.append({"weight": w, "relu": l < 3})
    n_cube = B * H * W * T
    x, phi = np.empty(n_cube, np.float32), np.empty(n_cube, np.float32)
    for i in range(n_cube):
        x[i] = lcg() + np.float32(0.5)
        phi[i] = 0.0 if lcg() < 0 else 1.0
    xt, pt = t(x.reshape(B, H, W, T), dev), t(phi.reshape(B, H, W, T), dev)
    plan = NativeDenoiser("dncnn", layers, precision="tc_split", device=dev)
    z, r = plan.reconstruct(ops.gap_forward(xt, pt), pt, ops.phi_sum(pt), x0=None, m=5, lam=1e-2, beta=1.0,
                            max_iter=12, tol=1e-5)
    zz = z.double()
    assert int(m.group(2)) == r.iterations and int(m.group(3)) == r.f_calls
    # smoke.c prints 10 significant digits
    assert abs(float(m.group(1)) - r.residual) <= 1e-8 * abs(r.residual)
    assert abs(float(m.group(4)) - float(zz.sum())) <= 1e-8 * abs(float(zz.sum()))
    assert abs(float(m.group(5)) - float((zz * zz).sum())) <= 1e-8 * float((zz * zz).sum())


def test_error_paths_on_device(dev):
    """Loud failures: FFDNet needs even H and W; a too-small workspace is refused, not overrun."""
    import ctypes
    import deqsci_b200
    from deqsci_b200 import _lib
    solver = build_solver("ffdnet", dev)
    plan = solver.nonlinear_op.native_plan(dev)
    z = torch.rand(1, 15, 16, 8, device=dev)
    with pytest.raises(deqsci_b200.DeqsciError, match="even H and W"):
        plan.denoise_residual(z, 0.1)
    z = torch.rand(1, 16, 16, 8, device=dev)
    out = torch.empty_like(z)
    ws = torch.empty(1024, dtype=torch.uint8, device=dev)
    rc = _lib.lib().deqsci_denoise_residual(plan._h, z.data_ptr(), ctypes.c_float(0.1), out.data_ptr(), ws.data_ptr(),
                                            ws.numel(), 1, 16, 16, 8, None)
    assert rc == -3 and b"workspace too small" in _lib.lib().deqsci_last_error()
    with pytest.raises(deqsci_b200.DeqsciError):
        plan.iterate(z, torch.rand(1, 16, 16, device=dev), torch.rand(1, 16, 8, 8, device=dev),
                     torch.rand(1, 16, 16, device=dev), 0.1)                 # inconsistent shapes


def test_integration_md_binding_stub_runs(dev):
    """The ctypes stub INTEGRATION.md tells a reference maintainer to paste into utils/cg_utils.py is executed
    as written (only the library path is made absolute) and must agree bit for bit with the package's own
    A_torch_ / At_torch_."""
    import re
    from conftest import ROOT
    from deqsci_b200 import _lib
    from deqsci_b200.utils.cg_utils import A_torch_, At_torch_
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "def A_torch_(x, Phi):" in b and "ctypes.CDLL" in b)
    stub = stub.replace('ctypes.CDLL("libdeqsci.so")', "ctypes.CDLL(%r)" % _lib.LIB_PATH)
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 24, 40, 8, generator=g).to(dev)
    Phi = (torch.rand(2, 24, 40, 8, generator=g) < 0.5).float().to(dev)
    y = ns["A_torch_"](x, Phi)
    assert torch.equal(y, A_torch_(x, Phi))
    assert torch.equal(ns["At_torch_"](y, Phi), At_torch_(y, Phi))


# ---------------------------------------------------------------------------------------------
# (9) the entry script with the reference's flags, on .mat files and a .ckpt in the reference's formats
# ---------------------------------------------------------------------------------------------
def test_entry_script_flags_mat_files_and_ckpt(dev, full_recon, tmp_path):
    """python -m deqsci_b200.video_sci_proxgrad --denoiser SimpleCNN --inference True ... (test_cnn.sh):
    scenes written as MATLAB v5 files with the reference's variable names, weights as a checkpoint dict
    with the reference's keys (incl. a 'module.' prefix the entry strips).  Reported average PSNR =
    the reference's (38.14 / 32.35 / 23.55 -> 31.35 dB)."""
    import scipy.io as sio
    from deqsci_b200 import video_sci_proxgrad as entry
    data_dir = tmp_path / "test_gray"
    data_dir.mkdir()
    want = []
    for scene, n in (("drop8", 1), ("runner8", 1), ("traffic", 6)):
        gt, mask, meas = load_scene(scene)
        m = np.round(meas * 255.0)
        if scene != "traffic":
            m = np.concatenate([m] + [m[:, :, :1]] * 4, axis=2)
        sio.savemat(str(data_dir / (scene + "_cacti.mat")),
                    {"orig": np.round(gt * 255).astype(np.uint8), "mask": mask.astype(np.uint8), "meas": m})
        want.append(np.mean([float(full_recon["SimpleCNN_%s_%d_psnr" % (scene, i)]) for i in range(n)]))
    sd = {"module." + k: torch.from_numpy(v) for k, v in load_weights("SimpleCNN").items()}
    ckpt = tmp_path / "cnn.ckpt"
    torch.save({"solver_state_dict": sd, "epoch": 7, "optimizer_state_dict": {}, "scheduler_state_dict": {}}, str(ckpt))
    psnr = entry.main(["--savepath", str(tmp_path / "save") + "/", "--testpath", str(data_dir) + "/",
                       "--loadpath", str(ckpt), "--denoiser", "SimpleCNN", "--inference", "True"])
    assert abs(psnr - float(np.mean(want))) <= 0.05
    assert abs(psnr - 31.35) <= 0.05
    pngs = os.listdir(str(tmp_path / "save" / "img" / "test"))
    assert len(pngs) == 64 and "drop8_cacti.mat_reconstruction_0.png" in pngs
