"""Import-time shims that let the UNMODIFIED reference (/root/reference) run on a modern,
CPU-only stack.  TEST INFRASTRUCTURE ONLY: used by tests/golden/make_golden.py (in the build
container, where /root/reference exists) to generate golden vectors.  Nothing in the product
package (deqsci_b200/) imports this file, and it is never used on the GPU box.

Shim list follows SURVEY.md §8(c) / Appendix A; each line says which reference import it unblocks.
"""
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("DEQSCI_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def use_staged_reference() -> bool:
    """Where /root/reference does not exist (the GPU box), import the reference's hot-path files from the tree
    oracle/make_ref.py staged under oracle/_ref/ (git-ignored, travels with the working tree).  Returns
    whether a reference tree is importable afterwards."""
    global REFERENCE_ROOT
    if not reference_available() and os.path.isdir(os.path.join(STAGED_ROOT, "solvers")):
        REFERENCE_ROOT = STAGED_ROOT
    return reference_available()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "solvers"))


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def skimage_psnr(image_true, image_test, data_range=None):
    """skimage.metrics.peak_signal_noise_ratio float rule: range 1 if true.min() >= 0 else 2."""
    t = np.asarray(image_true, np.float64)
    u = np.asarray(image_test, np.float64)
    R = data_range or (1.0 if t.min() >= 0 else 2.0)
    return 10 * np.log10(R * R / np.mean((t - u) ** 2))


_installed = False


class cpu_only:
    """Context manager: the reference hard-codes .cuda() (solvers/equilibrium_solvers_yaping.py:394,410); on a box
    WITH a GPU its CPU run needs those calls to be no-ops.  Patches torch.Tensor.cuda / nn.Module.cuda to identity
    for the duration of the block and restores them (bench.py's CPU legs run next to the GPU arm)."""

    def __enter__(self):
        self._saved = (torch.Tensor.cuda, torch.nn.Module.cuda)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda, torch.nn.Module.cuda = self._saved
        return False


def install_shims():
    """Must run BEFORE anything from the reference is imported."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    # solvers/*_yaping.py:3-5 import matplotlib / imageio at module level
    plt = _mod("matplotlib.pyplot")
    _mod("matplotlib", pyplot=plt, use=lambda *a, **k: None)
    _mod("imageio")
    _mod("h5py")  # utils/sci_dataloader.py:8
    _mod("skimage")
    _mod("skimage.metrics", peak_signal_noise_ratio=skimage_psnr)  # training/sci_equilibrium_training.py:11
    _mod("skimage.restoration", denoise_tv_chambolle=None)  # utils/cg_utils.py:6 (dead code)
    # utils/sci_dataloader.py:10-11 use private scipy names that moved
    import scipy.io.matlab.mio as mio
    import scipy.io.matlab.miobase as miobase
    from scipy.io.matlab._mio import _open_file
    from scipy.io.matlab._miobase import _get_matfile_version
    mio._open_file = _open_file
    miobase.get_matfile_version = _get_matfile_version
    # solvers/new_equilibrium_utils_yaping.py:180 uses the removed torch.solve(B, A)
    torch.solve = lambda B, A: (torch.linalg.solve(A, B), None)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def build_reference_deq(denoiser: str, max_iter: int, m: int = 5, beta: float = 1.0, state_dict=None):
    """Builds the reference objects exactly as video_sci_proxgrad.py:145-245 does (inference).

    denoiser in {'ffdnet', 'SimpleCNN', 'RealSN_SimpleCNN'}; weights: models/cnn.ckpt,
    models/rsn_cnn.ckpt, and (ffdnet.ckpt is a missing blob) networks/ffdnet/models/net_gray.pth
    re-keyed to nonlinear_op.* as the stand-in.
    """
    install_shims()
    import contextlib
    import io
    from solvers.equilibrium_solvers_yaping import EquilibriumProxGradSCI
    from solvers import new_equilibrium_utils_yaping as eq_utils
    from utils.cg_utils import A_torch_, At_torch_

    with contextlib.redirect_stdout(io.StringIO()):
        if denoiser == "ffdnet":
            from networks.ffdnet.models import FFDNet
            net = FFDNet(num_input_channels=1, tag="ffdnet")
        elif denoiser == "SimpleCNN":
            from networks.provable.model.SimpleCNN_models import DnCNN
            net = DnCNN(1, num_of_layers=4, lip=0.0, no_bn=True, tag="denoiser")
        elif denoiser == "RealSN_SimpleCNN":
            from networks.provable.model.SimpleCNN_models import DnCNN
            net = DnCNN(1, num_of_layers=4, lip=1.0, no_bn=True, tag="denoiser")
        else:
            raise ValueError(denoiser)
    net.eval()
    solver = EquilibriumProxGradSCI(A=A_torch_, At=At_torch_, nonlinear_operator=net, eta=0.2,
                                    minval=-1, maxval=1)
    sd = state_dict if state_dict is not None else reference_state_dict(denoiser)
    solver.load_state_dict(sd)
    deq = eq_utils.DEQFixedPoint(solver, eq_utils.andersonexp, m=m, beta=beta, lam=1e-2,
                                 max_iter=max_iter, tol=1e-5)
    return solver, deq


def reference_state_dict(denoiser: str):
    """solver_state_dict (keys 'nonlinear_op.*') for the shipped weights."""
    if denoiser == "ffdnet":
        raw = torch.load(os.path.join(REFERENCE_ROOT, "networks/ffdnet/models/net_gray.pth"),
                         map_location="cpu")
        return {"nonlinear_op." + (k[7:] if k.startswith("module.") else k): v for k, v in raw.items()}
    path = {"SimpleCNN": "models/cnn.ckpt", "RealSN_SimpleCNN": "models/rsn_cnn.ckpt"}[denoiser]
    saved = torch.load(os.path.join(REFERENCE_ROOT, path), map_location="cpu")
    return {(k[7:] if k.startswith("module.") else k): v for k, v in saved["solver_state_dict"].items()}
