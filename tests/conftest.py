import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_vectors():
    return dict(np.load(os.path.join(GOLDEN, "small_vectors.npz")))


@pytest.fixture(scope="session")
def full_recon():
    p = os.path.join(GOLDEN, "full_recon.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/full_recon.npz not generated")
    return dict(np.load(p))


def load_weights(denoiser):
    """solver_state_dict (numpy) of the shipped weights: keys 'nonlinear_op.*'."""
    f = {"ffdnet": "weights_ffdnet_gray.npz", "SimpleCNN": "weights_cnn.npz",
         "RealSN_SimpleCNN": "weights_rsn_cnn.npz"}[denoiser]
    d = np.load(os.path.join(GOLDEN, f))
    return {k: d[k] for k in d.files if not k.startswith("shape::")}


def load_scene(name):
    """(gt [H,W,F] fp32 in [0,1], mask [H,W,T] fp32, meas [H,W,M] fp32 /255) exactly as the
    reference loader returns them (utils/sci_dataloader.py:241-258); meas == sum_t mask*orig."""
    d = np.load(os.path.join(GOLDEN, "scenes.npz"))
    shape = tuple(d[name + "_mask_shape"])
    mask = np.unpackbits(d[name + "_mask_bits"])[:int(np.prod(shape))].reshape(shape).astype(np.float32)
    orig = d[name + "_orig"].astype(np.float32)
    T = shape[2]
    nm = orig.shape[2] // T
    meas = np.stack([(mask * orig[:, :, k * T:(k + 1) * T]).sum(2, dtype=np.float64).astype(np.float32)
                     for k in range(nm)], axis=2)
    return orig / np.float32(255), mask, meas / np.float32(255)


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
