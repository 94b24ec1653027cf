"""Iterate map of DE-GAP — drop-in for the reference's EquilibriumProxGradSCI
(solvers/equilibrium_solvers_yaping.py:382-436):

    z' = z + At((y - A z) / Phi_sum)          GAP data-consistency step
    z+ = z' - D(z')                           learned denoiser predicts the noise

At inference the whole map is ONE C-ABI call (deqsci_iterate): the GAP step is fused into the first
conv kernel, the residual subtract and layout change into the last.  The sigma schedule of the
'ffdnet' tag (reset to 60/255 when y.mean() changes, else x0.971 per call, :409-413) is kept on
the host as an fp32 scalar, without the reference's per-call device sync."""
import weakref

import numpy as np
import torch
import torch.nn as nn

from .._lib import DeqsciError
from ..utils import cg_utils

_SIGMA0 = np.float32(60 / 255)
_DECAY = np.float32(0.971)


class EquilibriumProxGradSCI(nn.Module):
    def __init__(self, A, At, nonlinear_operator, eta, minval=-1, maxval=1):
        super().__init__()
        self.A = A
        self.At = At
        self.nonlinear_op = nonlinear_operator
        self.minval = minval          # stored, never applied — as in the reference (no clamp in forward)
        self.maxval = maxval
        self.eta = eta
        self.y = 0                    # mean of the measurement the sigma schedule belongs to
        self._n = 0                   # calls made since the last reset: the next call uses sigma_table[_n]
        self._y_ref = None            # weak reference to the measurement tensor whose mean is self.y ...
        self._y_version = None        # ... and its version counter when the mean was taken
        self._undo = None
        self.n_sigma_frames = 8

    # ---- sigma schedule (host side) ------------------------------------------------------------
    _table = [_SIGMA0]                # sigma_k = fp32(sigma_{k-1} * fp32(0.971)), shared by all instances

    @classmethod
    def sigma_at(cls, k):
        while len(cls._table) <= k:
            cls._table.append(np.float32(cls._table[-1] * _DECAY))
        return cls._table[k]

    @property
    def _sigma(self):
        """sigma used by the most recent call (60/255 before any call)."""
        return self.sigma_at(max(self._n - 1, 0))

    @property
    def noise_sigma(self):
        """Per-frame sigma vector, like the reference's attribute (:394,410,413)."""
        dev = next(self.nonlinear_op.parameters()).device
        return torch.full((self.n_sigma_frames,), float(self._sigma), dtype=torch.float32, device=dev)

    def _observe(self, y):
        """Reference :409-412: `if self.y != y.mean(): reset` — the measurement means are compared BY VALUE.
        The mean is recomputed for every tensor object not seen before; only the very same live tensor
        object with an unchanged version counter (what a solver loop passes on every call of one solve)
        skips the reduction, so a solve costs one device sync instead of one per call.  Identity is held by
        a weak reference, never by address: a new measurement allocated where a freed one lived is a new
        object and is measured again.  Returns True when the schedule was reset."""
        try:
            version = y._version
        except RuntimeError:          # inference tensors carry no version counter: always re-measure
            version = None
        seen = self._y_ref() if self._y_ref is not None else None
        if seen is y and version is not None and version == self._y_version:
            return False
        mean = float(y.mean())
        self._y_ref, self._y_version = weakref.ref(y), version
        if self.y != mean:
            self.y = mean
            self._n = 0
            return True
        return False

    def __getstate__(self):
        state = self.__dict__.copy()  # weak references do not pickle (torch.save of the whole module)
        state["_y_ref"] = state["_y_version"] = state["_undo"] = None
        return state

    def _advance_sigma(self, y):
        """sigma of this call: 60/255 right after a reset, else the previous one times 0.971 (:413).
        (The very first call of a measurement whose mean equals the stored one decays, as in the
        reference.)"""
        self._undo = (self.y, self._n, self._y_ref, self._y_version)      # state before this call (rollback_call)
        if self._observe(y):
            self._n = 1
            return self.sigma_at(0)
        if self._n == 0:          # no reset on a never-reset schedule: the reference decays its initial 60/255
            self._n = 1
        self._n += 1
        return self.sigma_at(self._n - 1)

    def rollback_call(self):
        """Undoes the schedule advance of the most recent forward() (a solver that queued one
        speculative iteration past convergence calls this; the map has no other per-call state in
        eval mode)."""
        if self.nonlinear_op.tag == 'ffdnet' and self._undo is not None:
            self.y, self._n, self._y_ref, self._y_version = self._undo
            self._undo = None

    def skip_call(self):
        """Advances the sigma schedule as one forward() call would, without computing anything
        (DEQFixedPoint uses it for the reference's wasted second post-solver call at inference)."""
        if self.nonlinear_op.tag == 'ffdnet':
            self._n = max(self._n, 1) + 1

    # ---- forward -----------------------------------------------------------------------------------
    def _native_ok(self, z):
        op = self.nonlinear_op
        return (z.is_cuda and hasattr(op, "native_plan") and op.tag in ('ffdnet', 'denoiser')
                and getattr(op, "uses_native", lambda t: False)(z))

    def forward(self, z, y, Phi, Phi_sum, out=None):
        bsz, w, h, c = z.shape
        tag = self.nonlinear_op.tag
        if self._native_ok(z):
            sigma = 0.0
            if tag == 'ffdnet':
                self.n_sigma_frames = bsz * c
                sigma = float(self._advance_sigma(y))
            plan = self.nonlinear_op.native_plan(z.device)
            if self.A is cg_utils.A_torch_ and self.At is cg_utils.At_torch_:
                return plan.iterate(z, y, Phi, Phi_sum, sigma, out=out)
            # custom operator callables: unfused data-consistency step, native denoiser
            fb = self.A(z, Phi)
            z = z + self.At((y - fb) / Phi_sum, Phi)
            return plan.denoise_residual(z, sigma, out=out)
        op = self.nonlinear_op
        if (tag in ('ffdnet', 'denoiser') and getattr(op, "native_train_ok", lambda t: False)(z)
                and self.A is cg_utils.A_torch_ and self.At is cg_utils.At_torch_):
            # train mode under no_grad (the DEQ forward solve while training): batch-statistics BatchNorm
            # on the native kernels, running statistics updated once per call like nn.BatchNorm2d
            sigma = 0.0
            if tag == 'ffdnet':
                self.n_sigma_frames = bsz * c
                sigma = float(self._advance_sigma(y))
            plan = op.native_plan(z.device, train=True)
            return plan.iterate_train(z, y, Phi, Phi_sum, sigma, op.bn_slots(), out=out)
        if not z.is_cuda and not (self.nonlinear_op.training and torch.is_grad_enabled()):
            raise DeqsciError("EquilibriumProxGradSCI inference on %s: deqsci_b200 has no CPU path" % z.device)
        from ..backward import native_backward_ok, native_iterate
        if torch.is_grad_enabled() and native_backward_ok(self, z, y, Phi, Phi_sum):
            # the graph-attached call as ONE native autograd node: tensor-core forward that keeps the activations,
            # weight gradients from csrc/backward.cu (no autograd tape through cuDNN)
            sigma = 0.0
            if tag == 'ffdnet':
                self.n_sigma_frames = bsz * c
                sigma = float(self._advance_sigma(y))
            return native_iterate(self, z, y, Phi, Phi_sum, sigma)
        return self._autograd_forward(z, y, Phi, Phi_sum)

    def _autograd_forward(self, z, y, Phi, Phi_sum):
        """Graph-attached evaluation for training (PyTorch autograd; not the inference hot path).
        cuDNN's TF32 convolutions (PyTorch's default on Ampere+) put 1e-3..1e-2 errors on the
        iterates and gradients, far outside the parity bar, so the library path is pinned to fp32 — for the
        process, because the backward kernels of this graph run later, outside any scope this call could
        open (`restore_library_math()` undoes it)."""
        if z.is_cuda:
            pin_fp32_library_math()
        bsz, w, h, c = z.shape
        tag = self.nonlinear_op.tag
        if self.A is cg_utils.A_torch_ and self.At is cg_utils.At_torch_:
            fb = torch.sum(z * Phi, dim=3)
            z = z + ((y - fb) / Phi_sum)[:, :, :, None] * Phi
        else:                         # user-supplied operator callables: the same ones the eval path uses
            z = z + self.At((y - self.A(z, Phi)) / Phi_sum, Phi)
        frames = z.permute(0, 3, 1, 2).contiguous().view(bsz * c, 1, w, h)
        if tag == 'ffdnet':
            self.n_sigma_frames = bsz * c
            sigma = float(self._advance_sigma(y))
            noise = self.nonlinear_op(frames, torch.full((bsz * c,), sigma, dtype=z.dtype, device=z.device))
            return z - noise.view(bsz, c, w, h).permute(0, 2, 3, 1)
        if tag == 'denoiser':
            noise = self.nonlinear_op(frames)
            return z - noise.view(bsz, c, w, h).permute(0, 2, 3, 1)
        if tag == 'conv2d':
            return self.nonlinear_op(frames).view(bsz, c, w, h).permute(0, 2, 3, 1)
        raise DeqsciError("nonlinear_op tag %r is not on the DE-GAP path built here" % (tag,))


class EquilibriumADMMSCI(nn.Module):
    """ADMM iterate of the SCI path — drop-in for the reference's EquilibriumADMMSCI
    (solvers/equilibrium_solvers_yaping.py:438-465):

        s  = z + u
        z' = s + At((y - A s) / (Phi_sum + 1e-8))      the GAP step on z+u (native kernel)
        x  = D(z' - u)                                 the denoiser REPLACES its input (no residual subtract)
        u' = u - (z' - x)                              returns (z', u')

    As in the reference the denoiser is called with ONE argument and must carry a `conv3d` attribute
    (False = a batch of independent frames [B*T,1,H,W]); FFDNet (two arguments) and a DnCNN without that
    attribute fail here exactly as they do there."""

    def __init__(self, A, At, nonlinear_operator, eta, minval=-1, maxval=1):
        super().__init__()
        self.A = A
        self.At = At
        self.nonlinear_op = nonlinear_operator
        self.minval = minval
        self.maxval = maxval

    def forward(self, z, u, y, Phi, Phi_sum):
        from .. import ops
        bsz, w, h, c = z.shape
        s = z + u
        if (s.is_cuda and not torch.is_grad_enabled() and self.A is cg_utils.A_torch_
                and self.At is cg_utils.At_torch_):
            zn = ops.gap_step(s, y, Phi, Phi_sum + 1e-8)
        else:
            zn = s + self.At((y - self.A(s, Phi)) / (Phi_sum + 1e-8), Phi)
        if not self.nonlinear_op.conv3d:
            x = self.nonlinear_op((zn - u).permute(0, 3, 1, 2).contiguous().view(bsz * c, 1, w, h))
            x = x.view(bsz, c, w, h).permute(0, 2, 3, 1)
        else:
            x = self.nonlinear_op((zn - u).permute(0, 3, 1, 2).unsqueeze(1).contiguous())
            x = x.squeeze(1).permute(0, 2, 3, 1)
        return zn, u - (zn - x)


_tf32_before = None


def pin_fp32_library_math():
    """fp32 (not TF32) cuDNN / cuBLAS math for the autograd-attached part of the training step; remembers the
    flags it found so restore_library_math() can put them back."""
    global _tf32_before
    if torch.backends.cudnn.allow_tf32 or torch.backends.cuda.matmul.allow_tf32:
        if _tf32_before is None:
            _tf32_before = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False


def restore_library_math():
    global _tf32_before
    if _tf32_before is not None:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = _tf32_before
        _tf32_before = None
