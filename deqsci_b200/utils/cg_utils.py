"""SCI operator functions — drop-in for the reference's utils/cg_utils.py:85-90,124-129,228-229,
backed by the fused GAP kernels of libdeqsci (deqsci_b200/csrc/gap.cu)."""
import torch

from .. import ops


class _Forward(torch.autograd.Function):
    """sum_t x*Phi on the native kernel; the reference's expression is differentiable in both arguments,
    so is this one: d/dx = Phi * g[...,None] (the adjoint kernel), d/dPhi = x * g[...,None]."""

    @staticmethod
    def forward(ctx, x, Phi):
        ctx.save_for_backward(x, Phi)
        return ops.gap_forward(x, Phi)

    @staticmethod
    def backward(ctx, g):
        x, Phi = ctx.saved_tensors
        g = g.contiguous()
        gx = ops.gap_adjoint(g, Phi) if ctx.needs_input_grad[0] else None
        gphi = None
        if ctx.needs_input_grad[1]:
            gphi = ops.gap_adjoint(g, x)
            if Phi.shape[0] != x.shape[0]:          # broadcast [1,H,W,T] mask
                gphi = gphi.sum(0, keepdim=True)
        return gx, gphi


class _Adjoint(torch.autograd.Function):
    """y[...,None]*Phi on the native kernel; d/dy = sum_t g*Phi (the forward kernel), d/dPhi = g*y[...,None]."""

    @staticmethod
    def forward(ctx, y, Phi):
        ctx.save_for_backward(y, Phi)
        return ops.gap_adjoint(y, Phi)

    @staticmethod
    def backward(ctx, g):
        y, Phi = ctx.saved_tensors
        g = g.contiguous()
        gy = ops.gap_forward(g, Phi) if ctx.needs_input_grad[0] else None
        gphi = None
        if ctx.needs_input_grad[1]:
            gphi = g * y[..., None]
            if Phi.shape[0] != y.shape[0]:
                gphi = gphi.sum(0, keepdim=True)
        return gy, gphi


def _wants_graph(*ts):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


def A_torch_(x, Phi):
    """Forward model of snapshot compressive imaging: sum_t x*Phi.  [B,H,W,T] -> [B,H,W]
    (reference utils/cg_utils.py:85-90)."""
    if _wants_graph(x, Phi):
        return _Forward.apply(x, Phi)
    return ops.gap_forward(x, Phi)


def At_torch_(y, Phi):
    """Transpose of the forward model: y[...,None]*Phi.  [B,H,W] -> [B,H,W,T]
    (reference utils/cg_utils.py:124-129)."""
    if _wants_graph(y, Phi):
        return _Adjoint.apply(y, Phi)
    return ops.gap_adjoint(y, Phi)


def initial_point(y, Phi, Phi_sum, gt):
    """x0 = At(y, Phi); Phi_sum and gt are ignored, as in the reference (utils/cg_utils.py:228-229)."""
    return At_torch_(y, Phi)


def Phi_sum_(Phi):
    """sum_t Phi with zeros replaced by 1: the normaliser the reference's callers build at
    training/sci_equilibrium_training.py:61-62,162-163."""
    return ops.phi_sum(Phi)
