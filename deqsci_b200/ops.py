"""Thin tensor-level wrappers over the C-ABI: shape / dtype / device checks, raw pointers, the
current CUDA stream.  PyTorch is plumbing here (device memory + streams); all arithmetic happens
in libdeqsci.so."""
import ctypes

import torch

from . import _lib
from ._lib import DeqsciError, check, lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _req(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise DeqsciError("%s is on %s: deqsci_b200 has no CPU path (CUDA tensors only)" % (name, t.device))
    if t.dtype != torch.float32:
        raise DeqsciError("%s must be float32 (got %s)" % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise DeqsciError("%s must have %d dims (got shape %s)" % (name, ndim, tuple(t.shape)))
    return t if t.is_contiguous() else t.contiguous()


def _cube_dims(phi):
    B, H, W, T = phi.shape
    return int(B), int(H), int(W), int(T)


def _bcast_phi(phi, like):
    """The reference relies on broadcasting a [1,H,W,T] mask against a batch; materialise it."""
    if phi.shape[0] != like.shape[0]:
        if phi.shape[0] != 1:
            raise DeqsciError("Phi batch %d does not match %d" % (phi.shape[0], like.shape[0]))
        phi = phi.expand(like.shape[0], *phi.shape[1:]).contiguous()
    return phi


def gap_forward(x, phi):
    x, phi = _req(x, "x", 4), _req(phi, "Phi", 4)
    phi = _bcast_phi(phi, x)
    if x.shape != phi.shape:
        raise DeqsciError("x %s and Phi %s differ" % (tuple(x.shape), tuple(phi.shape)))
    B, H, W, T = _cube_dims(phi)
    out = torch.empty((B, H, W), dtype=torch.float32, device=x.device)
    if out.numel():
        with torch.cuda.device(x.device):
            check(lib().deqsci_gap_forward(x.data_ptr(), phi.data_ptr(), out.data_ptr(), B, H, W, T, _stream(x)),
                  "deqsci_gap_forward")
    return out


def gap_adjoint(y, phi):
    y, phi = _req(y, "y", 3), _req(phi, "Phi", 4)
    phi = _bcast_phi(phi, y)
    if tuple(y.shape) != tuple(phi.shape[:3]):
        raise DeqsciError("y %s and Phi %s differ" % (tuple(y.shape), tuple(phi.shape)))
    B, H, W, T = _cube_dims(phi)
    out = torch.empty((B, H, W, T), dtype=torch.float32, device=y.device)
    if out.numel():
        with torch.cuda.device(y.device):
            check(lib().deqsci_gap_adjoint(y.data_ptr(), phi.data_ptr(), out.data_ptr(), B, H, W, T, _stream(y)),
                  "deqsci_gap_adjoint")
    return out


def phi_sum(phi):
    phi = _req(phi, "Phi", 4)
    B, H, W, T = _cube_dims(phi)
    out = torch.empty((B, H, W), dtype=torch.float32, device=phi.device)
    if out.numel():
        with torch.cuda.device(phi.device):
            check(lib().deqsci_phi_sum(phi.data_ptr(), out.data_ptr(), B, H, W, T, _stream(phi)), "deqsci_phi_sum")
    return out


def gap_step(z, y, phi, phi_sum_, out=None):
    z, y, phi, phi_sum_ = _req(z, "z", 4), _req(y, "y", 3), _req(phi, "Phi", 4), _req(phi_sum_, "Phi_sum", 3)
    phi = _bcast_phi(phi, z)
    phi_sum_ = _bcast_phi(phi_sum_, z)
    B, H, W, T = _cube_dims(z)
    if tuple(phi.shape) != (B, H, W, T) or tuple(y.shape) != (B, H, W) or tuple(phi_sum_.shape) != (B, H, W):
        raise DeqsciError("gap_step: inconsistent shapes z %s y %s Phi %s Phi_sum %s" % (
            tuple(z.shape), tuple(y.shape), tuple(phi.shape), tuple(phi_sum_.shape)))
    if out is None:
        out = torch.empty_like(z)
    if out.numel():
        with torch.cuda.device(z.device):
            check(lib().deqsci_gap_step(z.data_ptr(), y.data_ptr(), phi.data_ptr(), phi_sum_.data_ptr(),
                                        out.data_ptr(), B, H, W, T, _stream(z)), "deqsci_gap_step")
    return out


def gap_vjp(v, phi, phi_sum_, add=None, out=None):
    v, phi, phi_sum_ = _req(v, "v", 4), _req(phi, "Phi", 4), _req(phi_sum_, "Phi_sum", 3)
    phi = _bcast_phi(phi, v)
    phi_sum_ = _bcast_phi(phi_sum_, v)
    B, H, W, T = _cube_dims(v)
    if add is not None:
        add = _req(add, "add", 4)
    if out is None:
        out = torch.empty_like(v)
    if out.numel():
        with torch.cuda.device(v.device):
            check(lib().deqsci_gap_vjp(v.data_ptr(), phi.data_ptr(), phi_sum_.data_ptr(),
                                       add.data_ptr() if add is not None else None, out.data_ptr(),
                                       B, H, W, T, _stream(v)), "deqsci_gap_vjp")
    return out


_adjoint_ws = {}


def adjoint_solve(grad, phi, phi_sum_, m=5, lam=1e-4, beta=1.0, max_iter=50, tol=1e-5):
    """andersonexp on g -> gap_vjp(g) + grad, started at grad, in ONE C-ABI call (deqsci_adjoint_solve): the
    backward solve of DEQFixedPoint's hook for tag 'ffdnet' (reference
    solvers/new_equilibrium_utils_yaping.py:274-277).  Returns (g [B,H,W,T], backward_res)."""
    from . import _lib
    grad, phi, phi_sum_ = _req(grad, "grad", 4), _req(phi, "Phi", 4), _req(phi_sum_, "Phi_sum", 3)
    phi = _bcast_phi(phi, grad)
    phi_sum_ = _bcast_phi(phi_sum_, grad)
    B, H, W, T = _cube_dims(grad)
    if tuple(phi.shape) != (B, H, W, T) or tuple(phi_sum_.shape) != (B, H, W):
        raise DeqsciError("adjoint_solve: inconsistent shapes grad %s Phi %s Phi_sum %s" % (
            tuple(grad.shape), tuple(phi.shape), tuple(phi_sum_.shape)))
    need = lib().deqsci_adjoint_solve_workspace_bytes(B, H, W, T, int(m))
    if need == 0:
        raise DeqsciError("adjoint_solve: unsupported shape or history m=%d" % m)
    ws = _adjoint_ws.get(grad.device)
    if ws is None or ws.numel() < need:
        _adjoint_ws[grad.device] = None
        ws = _adjoint_ws[grad.device] = torch.empty(need, dtype=torch.uint8, device=grad.device)
    out = torch.empty_like(grad)
    opts = _lib.SolverOpts(int(m), float(lam), float(beta), int(max_iter), float(tol), 0.0, 1.0, 0, 0, 1e-5)
    res = _lib.SolverResult()
    with torch.cuda.device(grad.device):
        check(lib().deqsci_adjoint_solve(grad.data_ptr(), phi.data_ptr(), phi_sum_.data_ptr(), out.data_ptr(),
                                         ctypes.byref(opts), ws.data_ptr(), ws.numel(), ctypes.byref(res), B, H, W, T,
                                         _stream(grad)), "deqsci_adjoint_solve")
    return out, float(res.residual)
