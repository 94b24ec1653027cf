"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py shims) on CPU in the build container.

    python tests/golden/make_golden.py --stage assets   # weights + scenes  (seconds)
    python tests/golden/make_golden.py --stage small    # operator / f / solver vectors on crops (~2 min)
    python tests/golden/make_golden.py --stage full     # full 256x256x8 reconstructions (~40 min CPU)

The reference cannot travel to the GPU box, the fixtures do.  The reference ships no golden
vectors of its own, so these files are the parity pin for oracle/deqsci_oracle.py
and for the CUDA path.  ffdnet.ckpt is a missing blob in the mount; the FFDNet weights here are
the reference's own networks/ffdnet/models/net_gray.pth (same architecture), re-keyed — the
stand-in named in SURVEY.md F2.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

DENOISERS = ["ffdnet", "SimpleCNN", "RealSN_SimpleCNN"]
WEIGHT_FILE = {"ffdnet": "weights_ffdnet_gray.npz", "SimpleCNN": "weights_cnn.npz",
               "RealSN_SimpleCNN": "weights_rsn_cnn.npz"}
SCENES = ["drop8", "runner8", "traffic"]
CROP = (slice(96, 160), slice(96, 160))


def stage_assets():
    import scipy.io as sio
    for d in DENOISERS:
        sd = ref_import.reference_state_dict(d)
        out = {}
        for k, v in sd.items():
            if k.endswith("weight_u"):          # power-iteration probe: training only, 1.2 MB
                out["shape::" + k] = np.array(v.shape)
                continue
            out[k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, WEIGHT_FILE[d]), **out)
        print(d, len(out), "tensors")
    scenes = {}
    for s in SCENES:
        f = sio.loadmat(os.path.join(ref_import.REFERENCE_ROOT, "data/test_gray/%s_cacti.mat" % s))
        mask, orig, meas = f["mask"], f["orig"], f["meas"]
        assert mask.dtype == np.uint8 and set(np.unique(mask)) <= {0, 1}
        assert orig.dtype == np.uint8
        nm = meas.shape[2]
        T = mask.shape[2]
        # np.float32(meas) is exactly sum_t mask*orig (so it need not be stored)
        for k in range(min(nm, orig.shape[2] // T)):
            re = (mask.astype(np.float64) * orig[:, :, k * T:(k + 1) * T]).sum(2)
            assert np.array_equal(np.float32(re), np.float32(meas[:, :, k])), (s, k)
        scenes[s + "_mask_bits"] = np.packbits(mask.reshape(-1))
        scenes[s + "_mask_shape"] = np.array(mask.shape)
        scenes[s + "_orig"] = orig
        scenes[s + "_nmeas"] = np.array(nm)
        print(s, mask.shape, orig.shape, meas.shape, meas.dtype)
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **scenes)


def load_scene_ref(name):
    """Loads a scene through the reference's own loader (utils/sci_dataloader.py:241-258)."""
    ref_import.install_shims()
    from utils.sci_dataloader import load_test_data
    d = load_test_data(os.path.join(ref_import.REFERENCE_ROOT, "data/test_gray/%s_cacti.mat" % name))
    return d["gt"], d["mask"], d["meas"]


def stage_small():
    ref_import.install_shims()
    from utils.cg_utils import A_torch_, At_torch_, initial_point
    out = {}
    # (a) operator on seeded random inputs, incl. a grey (non-binary) mask and zero columns
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(2, 16, 12, 8, generator=g)
    Phi = (torch.rand(2, 16, 12, 8, generator=g) < 0.5).float()
    Phi[0, 3, 4] = 0
    Phi[1] = torch.rand(16, 12, 8, generator=g)
    y = A_torch_(x, Phi)
    out["op_x"], out["op_Phi"], out["op_A"] = x.numpy(), Phi.numpy(), y.numpy()
    out["op_At"] = At_torch_(y, Phi).numpy()
    Phi_sum = torch.sum(Phi, axis=3)
    Phi_sum[Phi_sum == 0] = 1
    out["op_Phi_sum"] = Phi_sum.numpy()
    out["op_x0"] = initial_point(y, Phi, Phi_sum, None).numpy()

    # (b),(c) f calls and solver runs on a 64x64 crop of the traffic scene (B=1) and a B=2 batch
    gt, mask, meas = load_scene_ref("traffic")
    gt_c = torch.from_numpy(np.stack([gt[CROP][..., 0:8], gt[CROP][..., 8:16]]))      # [2,64,64,8]
    Phi_c = torch.from_numpy(np.stack([mask[CROP], mask[CROP]]))
    y_c = A_torch_(gt_c, Phi_c)
    Phi_sum_c = torch.sum(Phi_c, axis=3)
    Phi_sum_c[Phi_sum_c == 0] = 1
    out["crop_gt"], out["crop_Phi"], out["crop_y"] = gt_c.numpy(), Phi_c.numpy(), y_c.numpy()
    for d in DENOISERS:
        solver, deq = ref_import.build_reference_deq(d, max_iter=30)
        x0 = At_torch_(y_c, Phi_c)
        with torch.no_grad():
            f1 = solver(x0, y_c, Phi_c, Phi_sum_c)
            f2 = solver(f1, y_c, Phi_c, Phi_sum_c)           # second call: sigma decayed once
        out["f1_" + d], out["f2_" + d] = f1.numpy(), f2.numpy()
        # solver trace with B=2 (alpha per sample, res over the whole batch)
        solver, deq = ref_import.build_reference_deq(d, max_iter=30)
        zin = []
        h = solver.register_forward_pre_hook(lambda mod, args: zin.append(args[0].detach().clone()))
        z = deq.forward(y_c, Phi_c, Phi_sum_c, initial_point=x0, train_flag=False)
        h.remove()
        out["deq30_z_" + d] = z.detach().numpy()
        out["deq30_res_" + d] = np.array(deq.forward_res)
        out["deq30_innorm_" + d] = np.array([float(t.norm()) for t in zin])
        out["deq30_in10_" + d] = zin[10].numpy()
        print(d, "ncalls", len(zin), "res", deq.forward_res)
        # forward_iteration (Picard) for 6 steps, B=1
        from solvers import new_equilibrium_utils_yaping as eq
        solver, _ = ref_import.build_reference_deq(d, max_iter=30)
        with torch.no_grad():
            fz, res = eq.forward_iteration(lambda z: solver(z, y_c[:1], Phi_c[:1], Phi_sum_c[:1]),
                                           x0[:1], max_iter=6, tol=1e-5)
        out["picard6_z_" + d], out["picard6_res_" + d] = fz.numpy(), np.array(res)
    # (d) SSIM / PSNR definitions
    import pytorch_ssim
    a = gt_c.permute(0, 3, 1, 2).contiguous()
    b = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    out["ssim_a"], out["ssim_b"] = a.numpy(), b.numpy()
    out["ssim_val"] = np.array(float(pytorch_ssim.ssim(a, b)))
    out["psnr_val"] = np.array(ref_import.skimage_psnr(a.numpy(), b.numpy()))
    np.savez_compressed(os.path.join(HERE, "small_vectors.npz"), **out)


def stage_full(denoisers, scenes):
    ref_import.install_shims()
    import pytorch_ssim
    from utils.cg_utils import At_torch_
    path = os.path.join(HERE, "full_recon.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for d in denoisers:
        max_iter = 180 if d == "ffdnet" else 100       # test_ffdnet.sh:6 / entry default (:28)
        for s in scenes:
            gt, mask, meas = load_scene_ref(s)
            nm = 1 if s in ("drop8", "runner8") else meas.shape[2]   # sci_equilibrium_training.py:167-168
            Phi = torch.from_numpy(mask)[None]
            Phi_sum = torch.sum(Phi, axis=3)
            Phi_sum[Phi_sum == 0] = 1
            for fi in range(nm):
                key = "%s_%s_%d" % (d, s, fi)
                if key + "_psnr" in out:
                    continue
                y = torch.from_numpy(meas[:, :, fi])[None]
                g = torch.from_numpy(gt[:, :, fi * 8:(fi + 1) * 8])[None]
                solver, deq = ref_import.build_reference_deq(d, max_iter=max_iter)
                zin = []
                h = solver.register_forward_pre_hook(lambda mod, args: zin.append(args[0].detach().clone()))
                t0 = time.time()
                z = deq.forward(y, Phi, Phi_sum, initial_point=At_torch_(y, Phi), train_flag=False)
                dt = time.time() - t0
                h.remove()
                z = z.detach()
                rec = z.clip(0, 1)
                out[key + "_psnr"] = np.array(ref_import.skimage_psnr(g.numpy(), rec.numpy()))
                out[key + "_ssim"] = np.array(float(pytorch_ssim.ssim(rec.permute(0, 3, 1, 2).contiguous(),
                                                                        g.permute(0, 3, 1, 2).contiguous())))
                out[key + "_res"] = np.array(deq.forward_res)
                out[key + "_innorm"] = np.array([float(t.norm()) for t in zin])
                out[key + "_seconds"] = np.array(dt)
                out[key + "_ncalls"] = np.array(len(zin))
                if fi == 0:
                    for k in (2, 20, len(zin) - 2):
                        out[key + "_in%d_crop" % k] = zin[k][0][CROP].numpy()
                    out[key + "_z_crop"] = z[0][CROP].numpy()
                print(key, "psnr %.4f ssim %.5f res %.3e calls %d  %.1fs" %
                      (out[key + "_psnr"], out[key + "_ssim"], deq.forward_res, len(zin), dt), flush=True)
                np.savez_compressed(path, **out)


def stage_dncnn_bn():
    """models.DnCNN (the BatchNorm DnCNN behind `--denoiser DnCNN`) with 5 layers, seeded random
    weights and BatchNorm statistics, eval mode: weights + output on a [2,1,40,72] input."""
    ref_import.install_shims()
    from networks.provable.model.models import DnCNN
    torch.manual_seed(7)
    net = DnCNN(channels=1, num_of_layers=5, tag='denoiser')
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    net.eval()
    x = torch.rand(2, 1, 40, 72)
    with torch.no_grad():
        y = net(x)
    out = {"x": x.numpy(), "y": y.numpy()}
    for k, v in net.state_dict().items():
        out["sd::" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "dncnn_bn_vectors.npz"), **out)
    print("dncnn_bn", y.shape, float(y.abs().max()))


def realsn_probe(shape, layer):
    """Deterministic stand-in for the random power-iteration probe `weight_u` (so the fixture need not store
    400 KB of noise per layer): unit-norm sin pattern; tests/test_gpu_parity.py rebuilds it the same way."""
    n = int(np.prod(shape))
    u = torch.sin(torch.arange(n, dtype=torch.float32) * 0.37 + 1.3 * layer).reshape(shape)
    return u / float(torch.sqrt(torch.sum(u * u)))


def stage_realsn_dncnn():
    """realSN_models.DnCNN (`--denoiser RealSN_DnCNN`) with 3 layers, seeded random weights: (a) one
    TRAIN-mode forward on a [2,1,40,72] input -- the spectral-norm hook runs one power iteration per conv
    and stores `weight` / `weight_u` -- with the state before and after (probes `weight_u`: generated by
    realsn_probe before, every 16th element stored after); (b) the EVAL-mode output on the same input with
    the stored weights."""
    ref_import.install_shims()
    from networks.provable.model.realSN_models import DnCNN
    torch.manual_seed(11)
    net = DnCNN(channels=1, num_of_layers=3, tag='denoiser')
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    for i, m in enumerate(net.dncnn):
        if hasattr(m, "weight_u"):
            m.weight_u.copy_(realsn_probe(tuple(m.weight_u.shape), i))
    out = {}
    for k, v in net.state_dict().items():
        if not k.endswith("weight_u"):
            out["sd0::" + k] = v.clone().numpy()
    x = torch.rand(2, 1, 40, 72)
    net.train()
    y_train = net(x)
    for k, v in net.state_dict().items():
        v = v.detach().clone().numpy()
        out["sd1::" + k] = v.reshape(-1)[::16].copy() if k.endswith("weight_u") else v
    net.eval()
    with torch.no_grad():
        y_eval = net(x)
    out.update(x=x.numpy(), y_train=y_train.detach().numpy(), y_eval=y_eval.numpy())
    np.savez_compressed(os.path.join(HERE, "realsn_dncnn_vectors.npz"), **out)
    print("realsn_dncnn", y_eval.shape, float(y_eval.abs().max()), float(y_train.detach().abs().max()))


def stage_admm():
    """EquilibriumADMMSCI + admmexp + DEQFixedPointADMM (reference solvers/equilibrium_solvers_yaping.py:438-465,
    solvers/new_equilibrium_utils_yaping.py:396-451) on the 64x64 traffic crop, B=2, two single steps and a 4-step solve, with the
    one-argument frame denoiser the class supports: SimpleCNN weights (cnn.ckpt) and `conv3d = False` set by
    hand (the reference's DnCNN lacks the attribute)."""
    ref_import.install_shims()
    from solvers.equilibrium_solvers_yaping import EquilibriumADMMSCI
    from solvers import new_equilibrium_utils_yaping as eq
    from utils.cg_utils import A_torch_, At_torch_
    small = dict(np.load(os.path.join(HERE, "small_vectors.npz")))
    Phi_c, y_c = torch.from_numpy(small["crop_Phi"]), torch.from_numpy(small["crop_y"])
    Phi_sum_c = torch.sum(Phi_c, axis=3)
    Phi_sum_c[Phi_sum_c == 0] = 1
    solver0, _ = ref_import.build_reference_deq("SimpleCNN", max_iter=10)
    net = solver0.nonlinear_op
    net.conv3d = False
    f = EquilibriumADMMSCI(A_torch_, At_torch_, net, eta=0.2)
    x0 = At_torch_(y_c, Phi_c)
    out = {}
    with torch.no_grad():
        z1, u1 = f(x0, torch.zeros_like(x0), y_c, Phi_c, Phi_sum_c)
        z2, u2 = f(z1, u1, y_c, Phi_c, Phi_sum_c)
        out["z1"], out["u1"], out["z2"], out["u2"] = z1.numpy(), u1.numpy(), z2.numpy(), u2.numpy()
        deq = eq.DEQFixedPointADMM(f, eq.admmexp, eq.admmexp, m=5, beta=1.0, lam=1e-2, max_iter=6, tol=1e-5)
        z = deq.forward(y_c, Phi_c, Phi_sum_c, initial_point=[x0, torch.zeros_like(x0)], train_flag=False)
        out["deq_z"], out["deq_res"] = z.numpy(), np.array(deq.forward_res)
    np.savez_compressed(os.path.join(HERE, "admm_vectors.npz"), **out)
    print("admm: res", deq.forward_res, "|z|", float(z.norm()))


def stage_synthetic(count=2):
    """VERDICT r01 missing #4: the reference itself on the workload bench.py times -- measurements 0..count-1 of
    bench.synthetic_batch (kind = bench.DATA_KIND), DE-GAP-FFDnet, 180 iterations, batch 1 each: PSNR, SSIM,
    residual, the norm of the input of every iterate-map call, 64x64x8 crops of the inputs of calls
    2, 20, 40, 100 and 180 and of the reconstruction.  ~150 s per measurement on 8 threads."""
    ref_import.install_shims()
    import pytorch_ssim
    from utils.cg_utils import At_torch_
    import bench
    ys, ps, xs = bench.synthetic_batch(0, count)
    out = {"kind": np.array(bench.DATA_KIND), "seed": np.array(bench.SEED)}
    for i in range(count):
        y, Phi, g = ys[i:i + 1], ps[i:i + 1], xs[i:i + 1]
        Phi_sum = torch.sum(Phi, axis=3)
        Phi_sum[Phi_sum == 0] = 1
        solver, deq = ref_import.build_reference_deq("ffdnet", max_iter=180)
        zin = []
        keep = {}
        def hook(mod, args):
            k = len(zin)
            zin.append(float(args[0].detach().norm()))
            if k in (2, 20, 40, 100, 180):
                keep[k] = args[0].detach()[0][CROP].numpy().copy()
        h = solver.register_forward_pre_hook(hook)
        t0 = time.time()
        z = deq.forward(y, Phi, Phi_sum, initial_point=At_torch_(y, Phi), train_flag=False).detach()
        h.remove()
        rec = z.clip(0, 1)
        key = "m%d" % i
        out[key + "_psnr"] = np.array(ref_import.skimage_psnr(g.numpy(), rec.numpy()))
        out[key + "_ssim"] = np.array(float(pytorch_ssim.ssim(rec.permute(0, 3, 1, 2).contiguous(),
                                                                g.permute(0, 3, 1, 2).contiguous())))
        out[key + "_res"] = np.array(deq.forward_res)
        out[key + "_innorm"] = np.array(zin)
        out[key + "_ymean"] = np.array(float(y.mean()))
        for k, v in keep.items():
            out[key + "_in%d_crop" % k] = v
        out[key + "_z_crop"] = z[0][CROP].numpy()
        print(key, "psnr %.4f ssim %.5f res %.3e calls %d  %.1fs" % (out[key + "_psnr"], out[key + "_ssim"],
                                                                     deq.forward_res, len(zin), time.time() - t0), flush=True)
        np.savez_compressed(os.path.join(HERE, "synthetic_recon.npz"), **out)


def stage_train(wide=False):
    """One implicit-differentiation training step (reference training/sci_equilibrium_training.py:54-75)
    on a 32x32x8 crop (wide: 32x160x8, large enough for the native train-mode kernels), B=2,
    max_iter=12, denoiser in train mode: loss, parameter gradients, BatchNorm running statistics."""
    ref_import.install_shims()
    from utils.cg_utils import A_torch_, At_torch_
    gt, mask, meas = load_scene_ref("traffic")
    C = (slice(100, 132), slice(48, 208)) if wide else (slice(100, 132), slice(110, 142))
    gt_c = torch.from_numpy(np.stack([gt[C][..., 0:8], gt[C][..., 16:24]]))
    Phi_c = torch.from_numpy(np.stack([mask[C], mask[C]]))
    y_c = A_torch_(gt_c, Phi_c)
    Phi_sum_c = torch.sum(Phi_c, axis=3)
    Phi_sum_c[Phi_sum_c == 0] = 1
    out = {"gt": gt_c.numpy(), "Phi": Phi_c.numpy(), "y": y_c.numpy()}
    for d in DENOISERS:
        solver, deq = ref_import.build_reference_deq(d, max_iter=12)
        solver.train()
        solver.nonlinear_op.train()
        x0 = At_torch_(y_c, Phi_c)
        rec = deq.forward(y_c, Phi_c, Phi_sum_c, initial_point=x0)
        loss = torch.nn.MSELoss(reduction='mean')(rec, gt_c)
        loss.backward()
        out["loss_" + d] = np.array(float(loss))
        out["fres_" + d] = np.array(deq.forward_res)
        out["bres_" + d] = np.array(deq.backward_res)
        out["rec_" + d] = rec.detach().numpy()
        names, norms = [], []
        for n_, p_ in solver.named_parameters():
            if p_.grad is not None:
                names.append(n_)
                norms.append(float(p_.grad.norm()))
                if p_.grad.numel() <= 4096:
                    out["grad_%s::%s" % (d, n_)] = p_.grad.numpy().copy()
        out["gradnames_" + d] = np.array(names)
        out["gradnorms_" + d] = np.array(norms)
        for n_, b_ in solver.named_buffers():
            if n_.endswith("3.running_mean") or n_.endswith("3.running_var") or n_.endswith("39.running_var") \
                    or n_.endswith("3.num_batches_tracked"):
                out["buf_%s::%s" % (d, n_)] = b_.numpy().copy()
        print(d, "loss", float(loss), "fres", deq.forward_res, "bres", deq.backward_res, "n grads", len(names))
    np.savez_compressed(os.path.join(HERE, "train_wide_vectors.npz" if wide else "train_vectors.npz"), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True, choices=["assets", "small", "full", "train", "train_wide", "dncnn_bn", "realsn_dncnn", "admm", "synthetic"])
    ap.add_argument("--denoisers", nargs="*", default=DENOISERS)
    ap.add_argument("--scenes", nargs="*", default=SCENES)
    a = ap.parse_args()
    torch.manual_seed(0)
    if a.stage == "assets":
        stage_assets()
    elif a.stage == "small":
        stage_small()
    elif a.stage == "train":
        stage_train()
    elif a.stage == "train_wide":
        stage_train(wide=True)
    elif a.stage == "dncnn_bn":
        stage_dncnn_bn()
    elif a.stage == "realsn_dncnn":
        stage_realsn_dncnn()
    elif a.stage == "admm":
        stage_admm()
    elif a.stage == "synthetic":
        stage_synthetic()
    else:
        stage_full(a.denoisers, a.scenes)
