"""Device-side denoiser plan: owns a deqsci_denoiser handle (packed / split weights on one GPU)
and its scratch workspace.  Built by the nn.Module mirrors in deqsci_b200.networks from their
current parameters."""
import ctypes
import os
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import ConvLayer, DeqsciError, check, lib
from .ops import _bcast_phi, _req, _stream


def default_precision():
    """DEQSCI_PRECISION in {tc_split (default, parity mode), fp32, tc_single}."""
    name = os.environ.get("DEQSCI_PRECISION", "tc_split")
    if name not in _lib.PRECISIONS:
        raise DeqsciError("DEQSCI_PRECISION=%r not in %s" % (name, sorted(_lib.PRECISIONS)))
    return name


def graph_needed(module, *tensors):
    """True when an autograd graph built by this call could be back-propagated through: grad mode on and a
    parameter of `module` or one of `tensors` requires grad.  The native kernels return graph-less results,
    so they serve a call only when this is False (eval mode alone does not make a call inference: the
    reference builds the graph and registers its implicit-differentiation hook in eval mode too, eval only
    freezes the BatchNorm statistics)."""
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        return True
    return module is not None and any(p.requires_grad for p in module.parameters())


def _f32(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _destroy(handle):
    try:
        lib().deqsci_denoiser_destroy(handle)
    except Exception:
        pass


class SavedActivations:
    """What a forward call kept for the backward passes (device buffers + the ctypes view the C-ABI takes).
    release() hands the big plane buffers back to the plan's pool, so a training loop re-uses the same ~2 GB every
    step instead of going through the allocator (whose growth phase costs tens of milliseconds per step)."""
    _pool = None

    def release(self):
        if self._pool is not None:
            for t in list(self.acts) + [p for p in self.pre if p is not None]:
                self._pool.append(t)
        self.acts, self.pre, self._pool = [], [], None

    def struct(self):
        P = ctypes.c_void_p
        n = len(self.acts)
        self._acts_tab = (P * n)(*[a.data_ptr() for a in self.acts])
        self._pre_tab = (P * n)(*[p.data_ptr() if p is not None else None for p in self.pre])
        sf = _lib.SavedForward()
        sf.acts = ctypes.cast(self._acts_tab, ctypes.POINTER(P))
        sf.pre = ctypes.cast(self._pre_tab, ctypes.POINTER(P)) if any(p is not None for p in self.pre) else None
        sf.bn_record = self.bn_record.data_ptr() if self.bn_record is not None else None
        sf.zprime = self.zprime.data_ptr() if self.zprime is not None else None
        return sf


class NativeDenoiser:
    """kind: 'ffdnet' | 'dncnn'.  layers: list of dicts {weight [O,I,3,3], scale [O]|None,
    bias [O]|None, relu bool} (host arrays).  Lives on `device`."""

    def __init__(self, kind, layers, precision=None, device=None):
        self.kind = kind
        self.precision = precision or default_precision()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise DeqsciError("NativeDenoiser needs a CUDA device (got %s): no CPU path" % self.device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        keep = []
        arr = (ConvLayer * len(layers))()
        for i, L in enumerate(layers):
            w = _f32(L["weight"])
            keep.append(w)
            arr[i].cin, arr[i].cout, arr[i].relu = int(w.shape[1]), int(w.shape[0]), int(bool(L.get("relu", False)))
            arr[i].weight_host = w.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            for key, field in (("scale", "scale_host"), ("bias", "bias_host")):
                v = L.get(key)
                if v is not None:
                    v = _f32(v)
                    keep.append(v)
                    setattr(arr[i], field, v.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().deqsci_denoiser_create(_lib.NET_FFDNET if kind == "ffdnet" else _lib.NET_DNCNN,
                                               _lib.PRECISIONS[self.precision], len(layers), arr,
                                               ctypes.byref(handle)), "deqsci_denoiser_create")
        self._h = handle
        self._fin = weakref.finalize(self, _destroy, handle)
        self._ws = None
        self._rws = None
        self._bws = None
        self.num_layers = len(layers)
        self._layer_shapes = [(int(arr[i].cout), int(arr[i].cin)) for i in range(len(layers))]

    def update_weights(self, weights):
        """Refreshes the plan's conv weights from the live device tensors (deqsci_denoiser_update_weights):
        stream-ordered repack kernels, no host copy -- what a training loop needs after optimizer.step()."""
        if len(weights) != self.num_layers:
            raise DeqsciError("update_weights: %d tensors for %d conv layers" % (len(weights), self.num_layers))
        ptrs = (ctypes.c_void_p * self.num_layers)()
        for i, w in enumerate(weights):
            w = _req(w.detach(), "weight %d" % i, 4)
            self._check_dev(w)
            ptrs[i] = w.data_ptr()
        with torch.cuda.device(self.device):
            check(lib().deqsci_denoiser_update_weights(self._h, self.num_layers, ptrs,
                                                       torch.cuda.current_stream(self.device).cuda_stream),
                  "deqsci_denoiser_update_weights")

    def _workspace(self, B, H, W, T):
        need = lib().deqsci_denoiser_workspace_bytes(self._h, B, H, W, T)
        if need == 0:
            raise DeqsciError("unsupported cube shape B=%d H=%d W=%d T=%d: %s" % (
                B, H, W, T, lib().deqsci_last_error().decode()))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _check_dev(self, t):
        if t.device != self.device:
            raise DeqsciError("tensor on %s but denoiser plan lives on %s" % (t.device, self.device))

    def denoise_residual(self, zin, sigma=0.0, out=None):
        """out[B,H,W,T] = zin - D(zin), frames b*T+t, one sigma for the whole call."""
        zin = _req(zin, "z", 4)
        self._check_dev(zin)
        B, H, W, T = (int(s) for s in zin.shape)
        if out is None:
            out = torch.empty_like(zin)
        if zin.numel() == 0:
            return out
        ws = self._workspace(B, H, W, T)
        with torch.cuda.device(self.device):
            check(lib().deqsci_denoise_residual(self._h, zin.data_ptr(), float(sigma), out.data_ptr(), ws.data_ptr(),
                                                ws.numel(), B, H, W, T, _stream(zin)), "deqsci_denoise_residual")
        return out

    def iterate(self, z, y, phi, phi_sum, sigma=0.0, out=None):
        """out = denoise_residual(gap_step(z, y, phi, phi_sum)): one call of the iterate map f."""
        z, y = _req(z, "z", 4), _req(y, "y", 3)
        phi, phi_sum = _bcast_phi(_req(phi, "Phi", 4), z), _bcast_phi(_req(phi_sum, "Phi_sum", 3), z)
        self._check_dev(z)
        B, H, W, T = (int(s) for s in z.shape)
        if tuple(phi.shape) != (B, H, W, T) or tuple(y.shape) != (B, H, W) or tuple(phi_sum.shape) != (B, H, W):
            raise DeqsciError("iterate: inconsistent shapes z %s y %s Phi %s Phi_sum %s" % (
                tuple(z.shape), tuple(y.shape), tuple(phi.shape), tuple(phi_sum.shape)))
        if out is None:
            out = torch.empty_like(z)
        if z.numel() == 0:
            return out
        ws = self._workspace(B, H, W, T)
        with torch.cuda.device(self.device):
            check(lib().deqsci_iterate(self._h, z.data_ptr(), y.data_ptr(), phi.data_ptr(), phi_sum.data_ptr(),
                                       float(sigma), out.data_ptr(), ws.data_ptr(), ws.numel(), B, H, W, T,
                                       _stream(z)), "deqsci_iterate")
        return out

    # ---- backward of a conv / ReLU stack on the same kernels (tag 'denoiser', DE-GAP-CNN) ------------------
    def iterate_save(self, z, y, phi, phi_sum, sigma=0.0, bn_modules=None, want_zprime=False):
        """One call of the iterate map that also keeps what a backward pass needs (deqsci_iterate_save, or
        deqsci_iterate_train_save when bn_modules is given: batch-statistics BatchNorm, running statistics updated).
        Returns (out, saved): saved.acts[i] = output planes of conv layer i, fp16 [2, B*T, Hc, Wc, 64] (hi, lo);
        saved.pre[i] = the raw conv output of a layer followed by BatchNorm (train mode); saved.bn_record
        [num_layers, 256] = scale, shift, batch mean, 1/sqrt(var + eps); saved.zprime [B,T,H,W] = the input frames."""
        z, y = _req(z, "z", 4), _req(y, "y", 3)
        phi, phi_sum = _bcast_phi(_req(phi, "Phi", 4), z), _bcast_phi(_req(phi_sum, "Phi_sum", 3), z)
        self._check_dev(z)
        B, H, W, T = (int(s) for s in z.shape)
        out = torch.empty_like(z)
        ws = self._workspace(B, H, W, T)
        nbytes = lib().deqsci_denoiser_activation_bytes(self._h, B, H, W, T)
        n = self.num_layers - 1
        saved = SavedActivations()
        saved.shape = (B, H, W, T)
        saved.sigma = float(sigma)
        pool = self.__dict__.setdefault("_plane_pool", [])

        def take():
            for k, t in enumerate(pool):
                if t.numel() == nbytes:
                    return pool.pop(k)
            return torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        saved._pool = pool
        saved.acts = [take() for _ in range(n)]
        saved.pre = [None] * n
        saved.bn_record = None
        saved.zprime = torch.empty((B, T, H, W), dtype=torch.float32, device=self.device) if want_zprime else None
        if bn_modules is not None:
            saved.pre = [take() if bn_modules[i] is not None else None for i in range(n)]
            saved.bn_record = torch.zeros((self.num_layers, 256), dtype=torch.float32, device=self.device)
        sf = saved.struct()
        with torch.cuda.device(self.device):
            if bn_modules is None:
                check(lib().deqsci_iterate_save(self._h, z.data_ptr(), y.data_ptr(), phi.data_ptr(), phi_sum.data_ptr(),
                                                float(sigma), out.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(sf),
                                                B, H, W, T, _stream(z)), "deqsci_iterate_save")
            else:
                arr, momentum, eps = self._bn_table(bn_modules)
                check(lib().deqsci_iterate_train_save(self._h, z.data_ptr(), y.data_ptr(), phi.data_ptr(),
                                                      phi_sum.data_ptr(), float(sigma), out.data_ptr(), ws.data_ptr(),
                                                      ws.numel(), arr, momentum, eps, ctypes.byref(sf), B, H, W, T,
                                                      _stream(z)), "deqsci_iterate_train_save")
                for bn in bn_modules:
                    if bn is not None and bn.num_batches_tracked is not None:
                        bn.num_batches_tracked += 1
        return out, saved

    def backward_weights(self, adjoint, saved, grad, gammas, grad_scale=None):
        """Weight gradients of the call that produced `saved` (deqsci_backward_weights): returns
        (d_weight [list per conv layer], d_gamma, d_beta [lists, None where no BatchNorm follows]).  `adjoint`: the
        adjoint plan; gammas[i]: the BatchNorm weight tensor after conv layer i or None."""
        grad = _req(grad, "grad", 4)
        self._check_dev(grad)
        B, H, W, T = saved.shape
        if tuple(grad.shape) != (B, H, W, T):
            raise DeqsciError("backward_weights: grad %s does not match the saved call %s" % (tuple(grad.shape), saved.shape))
        if saved.zprime is None:
            raise DeqsciError("backward_weights: the forward call did not keep z' (want_zprime=True)")
        if grad_scale is None:
            gmax = float(grad.abs().max())
            grad_scale = 2.0 ** (-np.floor(np.log2(gmax))) if np.isfinite(gmax) and gmax > 0 else 1.0
            grad_scale = float(min(max(grad_scale, 2.0 ** -60), 2.0 ** 60))
        nl = self.num_layers
        shapes = self._layer_shapes
        dW = [torch.empty((co, ci, 3, 3), dtype=torch.float32, device=self.device) for (co, ci) in shapes]
        has_bn = [i < nl - 1 and saved.pre[i] is not None for i in range(nl)]
        dG = [torch.empty(64, dtype=torch.float32, device=self.device) if has_bn[i] else None for i in range(nl)]
        dBt = [torch.empty(64, dtype=torch.float32, device=self.device) if has_bn[i] else None for i in range(nl)]
        P = ctypes.c_void_p
        tab = lambda ts: (P * nl)(*[t.data_ptr() if t is not None else None for t in ts])
        gam = [gammas[i] if (has_bn[i] and gammas[i] is not None) else None for i in range(nl)]
        need = lib().deqsci_backward_workspace_bytes(self._h, B, H, W, T)
        if self._bws is None or self._bws.numel() < need:
            self._bws = None
            self._bws = torch.empty(need, dtype=torch.uint8, device=self.device)
        sf = saved.struct()
        with torch.cuda.device(self.device):
            check(lib().deqsci_backward_weights(self._h, adjoint._h, ctypes.byref(sf), tab(gam), grad.data_ptr(),
                                                float(grad_scale), float(saved.sigma), tab(dW), tab(dG), tab(dBt),
                                                self._bws.data_ptr(), self._bws.numel(), B, H, W, T, _stream(grad)),
                  "deqsci_backward_weights")
        return dW, dG, dBt

    def _mask_table(self, masks):
        if len(masks) != self.num_layers - 1:
            raise DeqsciError("%d mask planes for %d gated layers" % (len(masks), self.num_layers - 1))
        for m in masks:
            self._check_dev(m)
        return (ctypes.c_void_p * len(masks))(*[m.data_ptr() for m in masks])

    def denoise_residual_masked(self, v, masks, out=None):
        """On an ADJOINT plan: out = v - J_D^T v, the ReLUs of the stack replaced by the sign of the saved forward
        activations `masks` (masks[i] gates adjoint layer i; deqsci_denoise_residual_masked)."""
        v = _req(v, "v", 4)
        self._check_dev(v)
        B, H, W, T = (int(s) for s in v.shape)
        if out is None:
            out = torch.empty_like(v)
        ws = self._workspace(B, H, W, T)
        tab = self._mask_table(masks)
        with torch.cuda.device(self.device):
            check(lib().deqsci_denoise_residual_masked(self._h, v.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                       tab, B, H, W, T, _stream(v)), "deqsci_denoise_residual_masked")
        return out

    def adjoint_solve(self, grad, phi, phi_sum, masks, m=5, lam=1e-4, beta=1.0, max_iter=50, tol=1e-5, vjp_scale=None):
        """On an ADJOINT plan: the implicit-differentiation hook's backward solve for tag 'denoiser' in ONE C-ABI
        call (deqsci_adjoint_solve_denoiser): andersonexp on g -> gap_vjp(g - J_D^T g) + grad, started at grad
        (reference solvers/new_equilibrium_utils_yaping.py:274-277).  Returns (g, backward_res).
        vjp_scale (default: the power of two that brings max|grad| into [1, 2), one device sync): the conv stack is
        evaluated on vjp_scale * g and the result scaled back (fp16 operand planes; see include/deqsci.h)."""
        grad = _req(grad, "grad", 4)
        if vjp_scale is None:
            gmax = float(grad.abs().max())
            vjp_scale = 2.0 ** (-np.floor(np.log2(gmax))) if np.isfinite(gmax) and gmax > 0 else 1.0
            vjp_scale = float(min(max(vjp_scale, 2.0 ** -60), 2.0 ** 60))
        phi, phi_sum = _bcast_phi(_req(phi, "Phi", 4), grad), _bcast_phi(_req(phi_sum, "Phi_sum", 3), grad)
        self._check_dev(grad)
        B, H, W, T = (int(s) for s in grad.shape)
        need = lib().deqsci_reconstruct_workspace_bytes(self._h, B, H, W, T, int(m))
        if need == 0:
            raise DeqsciError("adjoint_solve: unsupported shape or history m=%d" % m)
        if self._rws is None or self._rws.numel() < need:
            self._rws = None
            self._rws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty_like(grad)
        opts = _lib.SolverOpts(int(m), float(lam), float(beta), int(max_iter), float(tol), 0.0, 1.0, 0, 0, 1e-5)
        res = _lib.SolverResult()
        tab = self._mask_table(masks)
        with torch.cuda.device(self.device):
            check(lib().deqsci_adjoint_solve_denoiser(self._h, tab, grad.data_ptr(), phi.data_ptr(), phi_sum.data_ptr(),
                                                      out.data_ptr(), ctypes.byref(opts), float(vjp_scale),
                                                      self._rws.data_ptr(), self._rws.numel(), ctypes.byref(res),
                                                      B, H, W, T, _stream(grad)),
                  "deqsci_adjoint_solve_denoiser")
        return out, float(res.residual)

    def _bn_table(self, bn_modules):
        if len(bn_modules) != self.num_layers:
            raise DeqsciError("%d BatchNorm slots for %d conv layers" % (len(bn_modules), self.num_layers))
        arr = (_lib.BNParams * self.num_layers)()
        momentum, eps = 0.1, 1e-5
        for i, bn in enumerate(bn_modules):
            if bn is None:
                continue
            if bn.momentum is None or not bn.track_running_stats:
                raise DeqsciError("native train-mode BatchNorm needs momentum and running statistics")
            momentum, eps = float(bn.momentum), float(bn.eps)
            arr[i].gamma = bn.weight.data_ptr() if bn.affine else None
            arr[i].beta = bn.bias.data_ptr() if bn.affine else None
            arr[i].running_mean = bn.running_mean.data_ptr()
            arr[i].running_var = bn.running_var.data_ptr()
        return arr, momentum, eps

    def iterate_train(self, z, y, phi, phi_sum, sigma, bn_modules, out=None):
        """One call of the iterate map with the denoiser in TRAIN mode (deqsci_iterate_train):
        batch-statistics BatchNorm, running statistics of `bn_modules` updated in place once (momentum,
        unbiased variance, num_batches_tracked += 1), like nn.BatchNorm2d.  bn_modules[i] is the
        BatchNorm2d that follows conv layer i of the plan, or None.  The plan must be a train plan
        (BatchNorm not folded)."""
        z, y = _req(z, "z", 4), _req(y, "y", 3)
        phi, phi_sum = _bcast_phi(_req(phi, "Phi", 4), z), _bcast_phi(_req(phi_sum, "Phi_sum", 3), z)
        self._check_dev(z)
        B, H, W, T = (int(s) for s in z.shape)
        arr, momentum, eps = self._bn_table(bn_modules)
        if out is None:
            out = torch.empty_like(z)
        ws = self._workspace(B, H, W, T)
        with torch.cuda.device(self.device):
            check(lib().deqsci_iterate_train(self._h, z.data_ptr(), y.data_ptr(), phi.data_ptr(), phi_sum.data_ptr(),
                                             float(sigma), out.data_ptr(), ws.data_ptr(), ws.numel(), arr, momentum,
                                             eps, B, H, W, T, _stream(z)), "deqsci_iterate_train")
        for bn in bn_modules:
            if bn is not None and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        return out

    def reconstruct(self, y, phi, phi_sum, x0=None, m=5, lam=1e-4, beta=1.0, max_iter=50, tol=1e-5,
                    sigma_start_call=0, final_call=True, sigma0=60 / 255, sigma_decay=0.971, bn_modules=None):
        """Whole DE-GAP reconstruction in ONE C-ABI call (deqsci_reconstruct): andersonexp on the
        iterate map + the final f call.  Returns (out [B,H,W,T], SolverResult).  bn_modules (train
        plans only): the solve runs with batch-statistics BatchNorm (deqsci_reconstruct_train) and
        updates the modules' running statistics / num_batches_tracked once per counted f call."""
        y = _req(y, "y", 3)
        phi, phi_sum = _bcast_phi(_req(phi, "Phi", 4), y), _bcast_phi(_req(phi_sum, "Phi_sum", 3), y)
        self._check_dev(y)
        B, H, W, T = (int(s) for s in phi.shape)
        if tuple(y.shape) != (B, H, W) or tuple(phi_sum.shape) != (B, H, W):
            raise DeqsciError("reconstruct: inconsistent shapes y %s Phi %s Phi_sum %s" % (
                tuple(y.shape), tuple(phi.shape), tuple(phi_sum.shape)))
        if x0 is not None:
            x0 = _req(x0, "x0", 4)
        out = torch.empty((B, H, W, T), dtype=torch.float32, device=self.device)
        need = lib().deqsci_reconstruct_workspace_bytes(self._h, B, H, W, T, int(m))
        if need == 0:
            raise DeqsciError("reconstruct: unsupported shape or history m=%d" % m)
        if self._rws is None or self._rws.numel() < need:
            self._rws = None
            self._rws = torch.empty(need, dtype=torch.uint8, device=self.device)
        opts = _lib.SolverOpts(int(m), float(lam), float(beta), int(max_iter), float(tol), float(sigma0),
                               float(sigma_decay), int(sigma_start_call), int(bool(final_call)), 1e-5)
        res = _lib.SolverResult()
        with torch.cuda.device(self.device):
            if bn_modules is None:
                check(lib().deqsci_reconstruct(self._h, y.data_ptr(), phi.data_ptr(), phi_sum.data_ptr(),
                                               x0.data_ptr() if x0 is not None else None, out.data_ptr(),
                                               ctypes.byref(opts), self._rws.data_ptr(), self._rws.numel(),
                                               ctypes.byref(res), B, H, W, T, _stream(y)), "deqsci_reconstruct")
            else:
                arr, momentum, eps = self._bn_table(bn_modules)
                check(lib().deqsci_reconstruct_train(self._h, y.data_ptr(), phi.data_ptr(), phi_sum.data_ptr(),
                                                     x0.data_ptr() if x0 is not None else None, out.data_ptr(),
                                                     ctypes.byref(opts), arr, momentum, eps, self._rws.data_ptr(),
                                                     self._rws.numel(), ctypes.byref(res), B, H, W, T, _stream(y)),
                      "deqsci_reconstruct_train")
                for bn in bn_modules:
                    if bn is not None and bn.num_batches_tracked is not None:
                        bn.num_batches_tracked += int(res.f_calls)
        return out, res

    def debug_hidden_layer(self, layer, act_in, NF, Hc, Wc):
        """Testing hook: act_in fp16 [2, NF, Hc, Wc, 64] (hi plane, lo plane) -> same shape."""
        assert act_in.dtype == torch.float16 and act_in.is_contiguous() and act_in.shape == (2, NF, Hc, Wc, 64)
        out = torch.empty_like(act_in)
        with torch.cuda.device(self.device):
            check(lib().deqsci_debug_hidden_layer(self._h, layer, act_in.data_ptr(), out.data_ptr(), NF, Hc, Wc,
                                                  _stream(act_in)), "deqsci_debug_hidden_layer")
        return out


class _PlanStore(dict):
    """Per-module plan cache; plans hold device handles, so copy.deepcopy / pickle / torch.save(module)
    carry an empty store instead."""

    def __deepcopy__(self, memo):
        return _PlanStore()

    def __reduce__(self):
        return (_PlanStore, ())


class NativePlanCache:
    """Mixin for the nn.Module mirrors: builds / caches a NativeDenoiser per (device, precision)
    and rebuilds it when any parameter or buffer changed (tensor version counters)."""

    def _plan_layers(self, train=False):  # -> (kind, [layer dicts]); train=True: BatchNorm NOT folded
        raise NotImplementedError

    def _plan_live_weights(self, train=False):
        """The conv weight tensors a plan can be refreshed from on the device (same order as the plan's
        layers), or None when the plan holds anything derived on the host (folded BatchNorm, spectral norm)."""
        return None

    def _plan_signature(self, train=False):
        # the train plan holds conv weights only: running statistics change on every call and must not
        # invalidate it
        tensors = [p for p in self.parameters() if p.dim() == 4] if train else \
            list(self.parameters()) + list(self.buffers())
        return tuple((id(t), t._version, t.data_ptr()) for t in tensors)

    # ---- the stack run backwards: layers reversed, weights transposed and flipped --------------------------
    def _plan_kind(self):
        raise NotImplementedError

    def _adjoint_conv_weights(self):
        """Conv weight tensors of the stack in forward order, or None when the stack is not a plain sequence of
        3x3 bias-free convolutions the adjoint plan can be derived from."""
        return None

    def native_adjoint_plan(self, device):
        """NativeDenoiser on W'[c][o][ky][kx] = W[o][c][2-ky][2-kx], layers in reverse order: its layers are the
        dgrads of the forward stack (backward solve of tag 'denoiser', weight-gradient pass).  FFDNet: the adjoint
        of the last layer takes the 4 pixel-unshuffled gradient channels behind a zero sigma channel, and the plan's
        own last layer (64 -> 4, never run) is zeros.  Refreshed on the device when the weights have changed."""
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        ws = self._adjoint_conv_weights()
        if ws is None:
            raise DeqsciError("no adjoint plan for this stack")
        kind = self._plan_kind()
        sig = tuple((id(w), w._version, w.data_ptr()) for w in ws)
        with torch.no_grad():
            adj = [w.detach().permute(1, 0, 2, 3).flip(2, 3).contiguous() for w in reversed(ws)]
            if kind == "ffdnet":
                adj[0] = torch.cat([torch.zeros_like(adj[0][:, :1]), adj[0]], 1).contiguous()
                adj[-1] = torch.zeros((4, 64, 3, 3), dtype=torch.float32, device=adj[0].device)
        hit = self.__dict__.get("_native_adjoint")
        if hit is not None and hit[1].device == device and hit[1].num_layers == len(adj):
            if hit[0] != sig:
                hit[1].update_weights(adj)
                self.__dict__["_native_adjoint"] = (sig, hit[1])
            return self.__dict__["_native_adjoint"][1]
        layers = [{"weight": a.float().cpu(), "scale": None, "bias": None, "relu": i < len(adj) - 1}
                  for i, a in enumerate(adj)]
        plan = NativeDenoiser(kind, layers, getattr(self, "precision", None), device)
        self.__dict__["_native_adjoint"] = (sig, plan)
        return plan

    def native_plan(self, device, precision=None, train=False):
        precision = precision or getattr(self, "precision", None) or default_precision()
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        cache = self.__dict__.setdefault("_native_cache", _PlanStore())
        key = (str(device), precision, bool(train))
        sig = self._plan_signature(train)
        hit = cache.get(key)
        if hit is None or hit[0] != sig:
            live = self._plan_live_weights(train)
            ident = None if live is None else tuple((id(t), t.data_ptr(), tuple(t.shape)) for t in live)
            if hit is not None and ident is not None and hit[2] == ident and all(t.device == device for t in live):
                hit[1].update_weights(live)       # same tensors, new values (an optimizer step): repack on the device
                cache[key] = (sig, hit[1], ident)
            else:
                kind, layers = self._plan_layers(train) if train else self._plan_layers()
                cache[key] = (sig, NativeDenoiser(kind, layers, precision, device), ident)
        return cache[key][1]
