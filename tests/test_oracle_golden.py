"""CPU: pins oracle/deqsci_oracle.py (numpy restatement) against vectors produced by the
reference's own code (tests/golden/make_golden.py).  Tolerances: the operator is elementwise fp32
(1e-6 relative, north_star); everything that runs a conv stack is compared at the fp32
implementation-noise floor measured in BASELINE.md §2 (<= 1e-5 per iterate), far inside the
1e-3 per-iterate bar."""
import numpy as np
import pytest

from conftest import load_weights, rel_l2
from oracle import deqsci_oracle as orc

DENOISERS = ["ffdnet", "SimpleCNN", "RealSN_SimpleCNN"]
TAG = {"ffdnet": "ffdnet", "SimpleCNN": "denoiser", "RealSN_SimpleCNN": "denoiser"}


def test_operator(small_vectors):
    v = small_vectors
    assert rel_l2(orc.A(v["op_x"], v["op_Phi"]), v["op_A"]) <= 1e-6
    np.testing.assert_array_equal(orc.At(v["op_A"], v["op_Phi"]), v["op_At"])
    np.testing.assert_allclose(orc.phi_sum(v["op_Phi"]), v["op_Phi_sum"], rtol=1e-6)   # grey mask: sum order
    np.testing.assert_array_equal(orc.phi_sum(v["op_Phi"][:1]), v["op_Phi_sum"][:1])  # binary mask: exact
    np.testing.assert_array_equal(orc.initial_point(v["op_A"], v["op_Phi"]), v["op_x0"])


def test_adjointness():
    rng = np.random.default_rng(0)
    x = rng.random((2, 8, 8, 8), dtype=np.float32)
    y = rng.random((2, 8, 8), dtype=np.float32)
    P = (rng.random((2, 8, 8, 8)) < 0.5).astype(np.float32)
    assert abs(np.vdot(orc.A(x, P), y) - np.vdot(x, orc.At(y, P))) <= 1e-4 * abs(np.vdot(x, orc.At(y, P)))


def _crop_inputs(v):
    Phi, y = v["crop_Phi"], v["crop_y"]
    return y, Phi, orc.phi_sum(Phi), orc.At(y, Phi)


@pytest.mark.parametrize("d", DENOISERS)
def test_f_two_calls(small_vectors, d):
    y, Phi, Ps, x0 = _crop_inputs(small_vectors)
    f = orc.ProxGradSCI(TAG[d], load_weights(d))
    f1 = f(x0, y, Phi, Ps)
    f2 = f(f1, y, Phi, Ps)          # exercises the sigma decay for ffdnet
    assert rel_l2(f1, small_vectors["f1_" + d]) <= 2e-5
    assert rel_l2(f2, small_vectors["f2_" + d]) <= 2e-5


@pytest.mark.parametrize("d", DENOISERS)
def test_deq_andersonexp_30(small_vectors, d):
    v = small_vectors
    y, Phi, Ps, x0 = _crop_inputs(v)
    f = orc.ProxGradSCI(TAG[d], load_weights(d))
    seen = []
    fm = lambda z, *a: (seen.append(z.copy()), f(z, *a))[1]
    z, res = orc.deq_forward(fm, y, Phi, Ps, x0=x0, m=5, beta=1.0, lam=1e-2, max_iter=30, tol=1e-5)
    assert len(seen) == len(v["deq30_innorm_" + d]) == 32
    norms = np.array([np.linalg.norm(s.astype(np.float64)) for s in seen])
    np.testing.assert_allclose(norms, v["deq30_innorm_" + d], rtol=2e-5)
    assert rel_l2(seen[10], v["deq30_in10_" + d]) <= 1e-4     # bar: 1e-3 per iterate
    assert rel_l2(z, v["deq30_z_" + d]) <= 1e-4
    assert abs(res - float(v["deq30_res_" + d])) <= 1e-3 * float(v["deq30_res_" + d])


@pytest.mark.parametrize("d", DENOISERS)
def test_forward_iteration(small_vectors, d):
    v = small_vectors
    y, Phi, Ps, x0 = _crop_inputs(v)
    f = orc.ProxGradSCI(TAG[d], load_weights(d))
    z, res = orc.forward_iteration(lambda t: f(t, y[:1], Phi[:1], Ps[:1]), x0[:1], max_iter=6, tol=1e-5)
    assert rel_l2(z, v["picard6_z_" + d]) <= 1e-5
    np.testing.assert_allclose(res, v["picard6_res_" + d], rtol=1e-3)


def test_metrics(small_vectors):
    v = small_vectors
    assert abs(orc.ssim(v["ssim_a"], v["ssim_b"]) - float(v["ssim_val"])) <= 1e-5
    assert abs(orc.psnr(v["ssim_a"], v["ssim_b"]) - float(v["psnr_val"])) <= 1e-9


def test_synthetic_sharding_invariance():
    a = orc.synthetic_measurements(3, 4, H=8, W=8)
    b = orc.synthetic_measurements(5, 1, H=8, W=8)
    np.testing.assert_array_equal(a["y"][2], b["y"][0])
    np.testing.assert_array_equal(a["Phi"][2], b["Phi"][0])


def test_dncnn_batchnorm_variant():
    """models.DnCNN (BatchNorm DnCNN, `--denoiser DnCNN`): oracle vs the reference module's output."""
    import os
    from conftest import GOLDEN
    v = dict(np.load(os.path.join(GOLDEN, "dncnn_bn_vectors.npz")))
    sd = {k[len("sd::"):]: v[k] for k in v if k.startswith("sd::")}
    got = orc.dncnn_forward(v["x"], sd, num_of_layers=5)
    assert rel_l2(got, v["y"]) <= 2e-6


def test_noise_floor_fixture():
    """tests/golden/noise_floor.npz (make_noise_floor.py): on the traffic scene with the FFDNet stand-in
    weights two fp32 implementations of the SAME algorithm -- this numpy oracle and the reference's own
    PyTorch run -- end up to 0.29 dB apart, with iterate norms diverging past 1e-3: the 0.05 dB / 1e-3
    bars cannot be met there by any re-implementation (the one documented exception)."""
    import os
    from conftest import GOLDEN
    v = dict(np.load(os.path.join(GOLDEN, "noise_floor.npz")))
    diffs = [abs(float(v["traffic_%d_oracle_psnr" % i]) - float(v["traffic_%d_reference_psnr" % i])) for i in (1, 2, 5)]
    assert max(diffs) > 0.25 and min(diffs) < 0.05
    assert all(float(v["traffic_%d_norm_rel_dev" % i].max()) > 1e-3 for i in (1, 2, 5))
    assert all(float(v["traffic_%d_norm_rel_dev" % i][:30].max()) < 1e-4 for i in (1, 2, 5))   # early iterates agree


def test_admm_sci(small_vectors):
    """EquilibriumADMMSCI / admmexp restatement vs the reference's own run (tests/golden/admm_vectors.npz,
    make_golden.py --stage admm; reference solvers/equilibrium_solvers_yaping.py:438-465,
    solvers/new_equilibrium_utils_yaping.py:396-451)."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "admm_vectors.npz")))
    y, Phi, Ps, x0 = _crop_inputs(small_vectors)
    f = orc.ADMMSCI(load_weights("SimpleCNN"))
    z1, u1 = f(x0, np.zeros_like(x0), y, Phi, Ps)
    z2, u2 = f(z1, u1, y, Phi, Ps)
    for got, key in ((z1, "z1"), (u1, "u1"), (z2, "z2"), (u2, "u2")):
        assert rel_l2(got, g[key]) <= 2e-5, key
    z, u, res = orc.admmexp(lambda a, b: f(a, b, y, Phi, Ps), [x0, np.zeros_like(x0)], m=5, lam=1e-2, max_iter=6,
                            tol=1e-5, beta=1.0)
    assert rel_l2(z, g["deq_z"]) <= 1e-4
    assert abs(res - float(g["deq_res"])) <= 1e-4 * float(g["deq_res"])


def test_benchmark_workload_first_calls():
    """The oracle on the workload bench.py times (bench.synthetic_batch measurement 0, full 256x256x8): the inputs
    of the first 6 iterate-map calls have the norms the reference's own run recorded (synthetic_recon.npz)."""
    import os
    import bench
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "synthetic_recon.npz")))
    assert str(g["kind"]) == bench.DATA_KIND
    ys, ps, _ = bench.synthetic_batch(0, 1)
    y, Phi = ys.numpy(), ps.numpy()
    f = orc.ProxGradSCI("ffdnet", load_weights("ffdnet"))
    orc.set_conv_backend("torch")
    seen = []
    fm = lambda z, *a: (seen.append(float(np.linalg.norm(z.astype(np.float64)))), f(z, *a))[1]
    Ps = orc.phi_sum(Phi)
    orc.andersonexp(lambda z: fm(z, y, Phi, Ps), orc.At(y, Phi), m=5, lam=1e-2, max_iter=6, tol=1e-5, beta=1.0)
    np.testing.assert_allclose(np.array(seen), g["m0_innorm"][:6], rtol=2e-5)
