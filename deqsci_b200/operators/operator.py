"""Operator interface of the reference (operators/operator.py:3-32) kept intact — `forward`,
`adjoint`, `gramian` — plus `SCIOperator`, the coded-mask snapshot operator expressed through that
interface on the fused CUDA kernels (the reference's SCI path passes bare functions instead,
utils/cg_utils.py:85-90,124-129)."""
import torch

from ..utils.cg_utils import A_torch_, At_torch_


class LinearOperator(torch.nn.Module):
    """Base class: subclasses provide forward(x) = A x and adjoint(y) = A^T y."""

    def __init__(self):
        super().__init__()

    def forward(self, x):
        return None

    def adjoint(self, x):
        return None

    def gramian(self, x):
        """A^T A x."""
        return self.adjoint(self.forward(x))


class SelfAdjointLinearOperator(LinearOperator):
    """A = A^T."""

    def adjoint(self, x):
        return self.forward(x)


class Identity(SelfAdjointLinearOperator):
    def forward(self, x):
        return x


class OperatorPlusNoise(torch.nn.Module):
    """y = A x + sigma * n, n ~ N(0, I)."""

    def __init__(self, operator, noise_sigma):
        super().__init__()
        self.internal_operator = operator
        self.noise_sigma = noise_sigma

    def forward(self, x):
        clean = self.internal_operator(x)
        return clean + self.noise_sigma * torch.randn_like(clean)


class SCIOperator(LinearOperator):
    """Snapshot compressive imaging with a fixed mask Phi [B or 1, H, W, T]:
    forward(x[B,H,W,T]) = sum_t x*Phi -> [B,H,W];  adjoint(y[B,H,W]) = y[...,None]*Phi."""

    def __init__(self, Phi):
        super().__init__()
        self.register_buffer("Phi", Phi)

    def forward(self, x):
        return A_torch_(x, self.Phi)

    def adjoint(self, y):
        return At_torch_(y, self.Phi)
