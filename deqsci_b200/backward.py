"""The graph-attached iterate-map call on the native kernels.

The reference evaluates `z = f(z*)` once with the autograd tape (solvers/new_equilibrium_utils_yaping.py:268) and
loss.backward() runs cuDNN dgrad / wgrad / BatchNorm-backward kernels through it.  `NativeIterate` is that call as ONE
autograd node: forward = deqsci_iterate_save / deqsci_iterate_train_save (the tensor-core stack, keeping the
activations), backward = deqsci_backward_weights (csrc/backward.cu + the adjoint plan's dgrads).  Gradients are
produced for the denoiser's parameters only: the iterate itself is the detached solver output, exactly as in the
reference, and FFDNet detaches its input anyway (networks/ffdnet/models.py:103-104)."""
import os

import torch
import torch.nn as nn


def parameter_layout(seq):
    """[(parameter, role, conv layer index)] in nn.Module.parameters() order for a Conv2d / BatchNorm2d / ReLU
    Sequential: role 'w' = conv weight, 'g' / 'b' = weight / bias of the BatchNorm that follows conv layer i."""
    out, conv = [], -1
    for mod in seq:
        if isinstance(mod, nn.Conv2d):
            conv += 1
            out.append((mod.weight, 'w', conv))
        elif isinstance(mod, nn.BatchNorm2d):
            if mod.affine:
                out.append((mod.weight, 'g', conv))
                out.append((mod.bias, 'b', conv))
        elif not isinstance(mod, nn.ReLU):
            return None
    return out


def stack_of(op):
    seq = getattr(getattr(op, "intermediate_dncnn", None), "itermediate_dncnn", None)
    return seq if seq is not None else getattr(op, "dncnn", None)


def native_backward_ok(solver, z, y, Phi, Phi_sum):
    """The native autograd node serves the call when: CUDA fp32, nothing but the denoiser's parameters needs a
    gradient, the stack is plain Conv2d / BatchNorm2d / ReLU on the tensor-core kernels (conv images wider than 64
    pixels, precision tc_split), and BatchNorm (if any) is in train mode (eval-mode BatchNorm is an affine the
    backward kernels do not cover).  DEQSCI_NATIVE_BACKWARD=0 switches it off (autograd / cuDNN instead)."""
    from .native import default_precision
    from .utils import cg_utils
    if os.environ.get("DEQSCI_NATIVE_BACKWARD", "1") == "0":
        return False
    op = solver.nonlinear_op
    if getattr(op, "tag", None) not in ("ffdnet", "denoiser") or not hasattr(op, "native_adjoint_plan"):
        return False
    if not (z.is_cuda and z.dtype == torch.float32 and z.dim() == 4):
        return False
    if any(t.requires_grad for t in (z, y, Phi, Phi_sum)):
        return False
    if solver.A is not cg_utils.A_torch_ or solver.At is not cg_utils.At_torch_:
        return False
    seq = stack_of(op)
    if seq is None or parameter_layout(seq) is None:
        return False
    if any(hasattr(m, "weight_orig") or hasattr(m, "plan_weight") for m in seq):      # spectral-norm hooks
        return False
    convs = [m for m in seq if isinstance(m, nn.Conv2d)]
    if len(convs) < 3 or any(c.bias is not None or tuple(c.weight.shape[2:]) != (3, 3) for c in convs):
        return False
    has_bn = any(isinstance(m, nn.BatchNorm2d) for m in seq)
    if has_bn and not (op.training and all(m.track_running_stats and m.momentum is not None
                                           for m in seq if isinstance(m, nn.BatchNorm2d))):
        return False
    if (getattr(op, "precision", None) or default_precision()) != "tc_split":
        return False
    if os.environ.get("DEQSCI_TC_PAIR", "1") == "0" or os.environ.get("DEQSCI_TC_FIRST", "1") == "0" \
            or os.environ.get("DEQSCI_TC_LAST", "1") == "0":
        return False
    H, W = int(z.shape[1]), int(z.shape[2])
    sc = 2 if op.tag == "ffdnet" else 1
    if op.tag == "ffdnet" and (getattr(op, "num_input_channels", 1) != 1 or H % 2 or W % 2):
        return False
    if op.tag == "denoiser" and getattr(op, "channels", 1) != 1:
        return False
    return W // sc > 64


class NativeIterate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, solver, z, y, Phi, Phi_sum, sigma, *params):
        op = solver.nonlinear_op
        seq = stack_of(op)
        train_bn = any(isinstance(m, nn.BatchNorm2d) for m in seq)
        plan = op.native_plan(z.device, train=train_bn)
        out, saved = plan.iterate_save(z.detach(), y, Phi, Phi_sum, sigma, bn_modules=op.bn_slots() if train_bn else None,
                                       want_zprime=True)
        ctx.op, ctx.plan, ctx.saved = op, plan, saved
        return out

    @staticmethod
    def backward(ctx, grad):
        if ctx.saved is None:
            raise RuntimeError("NativeIterate: the saved activations were released by the first backward pass "
                               "(retain_graph / double backward is not supported; DEQSCI_NATIVE_BACKWARD=0 uses autograd)")
        op, seq = ctx.op, stack_of(ctx.op)
        layout = parameter_layout(seq)
        n_conv = ctx.plan.num_layers
        gammas = [None] * n_conv
        for p_, role, i in layout:
            if role == 'g':
                gammas[i] = p_.detach()
        adj = op.native_adjoint_plan(grad.device)
        dW, dG, dB = ctx.plan.backward_weights(adj, ctx.saved, grad.contiguous(), gammas)
        ctx.saved.release()                                # the activations are dead after one backward pass
        ctx.saved = None
        grads = []
        for p_, role, i in layout:
            g = {'w': dW, 'g': dG, 'b': dB}[role][i]
            grads.append(g.view_as(p_) if (g is not None and p_.requires_grad) else None)
        return (None,) * 6 + tuple(grads)


def native_iterate(solver, z, y, Phi, Phi_sum, sigma):
    params = [p_ for p_, _, _ in parameter_layout(stack_of(solver.nonlinear_op))]
    return NativeIterate.apply(solver, z, y, Phi, Phi_sum, float(sigma), *params)
