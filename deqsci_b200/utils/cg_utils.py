"""SCI operator functions — drop-in for the reference's utils/cg_utils.py:85-90,124-129,228-229,
backed by the fused GAP kernels of libdeqsci (deqsci_b200/csrc/gap.cu)."""
from .. import ops


def A_torch_(x, Phi):
    """Forward model of snapshot compressive imaging: sum_t x*Phi.  [B,H,W,T] -> [B,H,W]
    (reference utils/cg_utils.py:85-90)."""
    return ops.gap_forward(x, Phi)


def At_torch_(y, Phi):
    """Transpose of the forward model: y[...,None]*Phi.  [B,H,W] -> [B,H,W,T]
    (reference utils/cg_utils.py:124-129)."""
    return ops.gap_adjoint(y, Phi)


def initial_point(y, Phi, Phi_sum, gt):
    """x0 = At(y, Phi); Phi_sum and gt are ignored, as in the reference (utils/cg_utils.py:228-229)."""
    return At_torch_(y, Phi)


def Phi_sum_(Phi):
    """sum_t Phi with zeros replaced by 1: the normaliser the reference's callers build at
    training/sci_equilibrium_training.py:61-62,162-163."""
    return ops.phi_sum(Phi)
